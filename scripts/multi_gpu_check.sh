# N-GPU evidence (run under gpurun --gpus N): slab parity at N ranks (interior ranks talk to both neighbours),
# the metric config split over N GPUs, BASELINE config 3 (lifting 2048^2 x 32) split over N GPUs
set -x
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
PB_HALO=p2p timeout 200 $TR --nproc-per-node $N --master-port 29581 tests/slab_worker.py rof_tiles65,rof_tiles63,rof_big,rof_goldstein,lifting,tvl1 > gpurun_out/mg_slab_n${N}.log 2>&1
grep SLAB_REPORT gpurun_out/mg_slab_n${N}.log | python -c "
import sys, json
rep = json.loads(sys.stdin.readline()[len('SLAB_REPORT '):])
print('world', rep['world'], 'p2p', rep['p2p'])
for k, c in rep['cases'].items():
    print(k, 'max err', max(c['err'].values()), 'one_pass', c['one_pass'], 'of', c['iters'])
"
timeout 200 $TR --nproc-per-node $N --master-port 29582 bench.py --gpus $N --steps 2000 --warmup 20 > gpurun_out/mg_rof_n${N}.json 2> gpurun_out/mg_rof_n${N}.err
tail -2 gpurun_out/mg_rof_n${N}.err
timeout 300 $TR --nproc-per-node $N --master-port 29583 scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/mg_lift_n${N}.json 2> gpurun_out/mg_lift_n${N}.err
tail -2 gpurun_out/mg_lift_n${N}.err
python - <<PY
import json
for f in ["gpurun_out/mg_rof_n${N}.json", "gpurun_out/mg_lift_n${N}.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), d.get("halo_mode"),
              "e2e", d.get("e2e", {}).get("value"), "ttr", d.get("time_to_residual_1e-4", {}).get("seconds"))
    except Exception as e:
        print("ERR", f, e)
PY
