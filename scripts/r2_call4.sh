# Round 2, fourth call (2 GPUs): flag-in-data halo protocol -- slab parity + bench on 2 GPUs
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q -x > gpurun_out/r2c4_pytest.log 2>&1
tail -5 gpurun_out/r2c4_pytest.log
timeout 300 $TR --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --steps 2000 --warmup 50 --no-workloads --no-cpu-baseline > gpurun_out/r2c4_bench_n2.json 2> gpurun_out/r2c4_bench_n2.err
tail -2 gpurun_out/r2c4_bench_n2.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c4_bench_n2.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "launches", d["gpu_launches"], "tile ms", r["ms_per_launch"],
          "check ms", r["residual_refresh_iterations"]["ms_per_launch"], "hash", d["iterate_hash"], "e2e", round(d["e2e"]["value"], 1), "ttr", d["time_to_residual_1e-4"]["seconds"], d["time_to_residual_1e-4"]["iterations"])
except Exception as e:
    print("ERR", e)
PY
