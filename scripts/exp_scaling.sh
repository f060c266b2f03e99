# multi-GPU check (run under gpurun --gpus N): slab tests, then strong scaling of the metric config over N GPUs
set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_slab.py -q 2>&1 | tail -5
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 150 $TR --nproc-per-node $n --master-port $((29560+n)) bench.py --gpus $n --steps 2000 --warmup 20 > gpurun_out/s_n${n}.json 2> gpurun_out/s_n${n}.err
    tail -2 gpurun_out/s_n${n}.err
  fi
done
for f in gpurun_out/s_n*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "dual ms", round(r["ms_per_launch"],4), "primal ms", round(r["primal_pass"]["ms_per_launch"],4), "e2e", round(d["e2e"]["value"],1), d.get("halo_mode"))
except Exception as e:
    print("ERR", e)
PY
done
