# 2-GPU check of the one-pass ring kernel on slabs (run under gpurun --gpus N): slab parity tests, then the
# metric config split over N GPUs with the one-pass kernel and (A/B) with the two-pass kernels
set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_slab.py -q -x 2>&1 | tail -25) > gpurun_out/sr_pytest.log
tail -25 gpurun_out/sr_pytest.log
timeout 200 $TR --nproc-per-node $N --master-port 29571 bench.py --gpus $N --steps 2000 --warmup 20 > gpurun_out/sr_n${N}.json 2> gpurun_out/sr_n${N}.err
tail -3 gpurun_out/sr_n${N}.err
PB_SLAB_TILE=0 timeout 200 $TR --nproc-per-node $N --master-port 29572 bench.py --gpus $N --steps 2000 --warmup 20 > gpurun_out/sr_n${N}_twopass.json 2> gpurun_out/sr_n${N}_twopass.err
tail -3 gpurun_out/sr_n${N}_twopass.err
for f in gpurun_out/sr_n${N}.json gpurun_out/sr_n${N}_twopass.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[1], d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "kernel ms", round(r["ms_per_launch"],4), "one_pass", d.get("one_pass_iterations"), "e2e", round(d["e2e"]["value"],1), d.get("halo_mode"))
except Exception as e:
    print("ERR", sys.argv[1], e)
PY
done
