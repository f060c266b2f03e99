# Round 2, twenty-eighth call (8 GPUs): lifting at N = 4, 8 after the epigraph-projection rewrite; 4-rank sharded ADMM / slab tests
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
timeout 100 $TR --nproc-per-node $N --master-port 2958$N scripts/bench_lifting.py --steps 40 --warmup 5 > gpurun_out/r2c28_lift_n$N.json 2> gpurun_out/r2c28_lift_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c28_lift_n$N.json").read().strip().splitlines()[-1])
print("lifting n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4))
PY
done
timeout 300 python -m pytest tests/test_gpu_admm_sharded.py -m gpu -q > gpurun_out/r2c28_pytest.log 2>&1
tail -3 gpurun_out/r2c28_pytest.log | cut -c1-300
