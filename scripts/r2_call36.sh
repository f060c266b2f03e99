# Round 2, thirty-sixth call (2 GPUs): cooperative staging on column slabs: slab tests, lifting at N = 2
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q > gpurun_out/r2c36_pytest.log 2>&1
tail -5 gpurun_out/r2c36_pytest.log | cut -c1-300
timeout 120 $TR --nproc-per-node 2 --master-port 29583 scripts/bench_lifting.py --steps 40 --warmup 5 > gpurun_out/r2c36_lift_n2.json 2> gpurun_out/r2c36_lift_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c36_lift_n2.json").read().strip().splitlines()[-1])
print("lifting n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4))
PY
