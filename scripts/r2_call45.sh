# Round 2, forty-fifth call (1 GPU): the driver's short invocation (K = 20, W = 3) on the final build
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2c45_bench_k20.json 2> gpurun_out/r2c45_bench_k20.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c45_bench_k20.json").read().strip().splitlines()[-1])
print(d["steps"], d["warmup"], round(d["value"], 1), round(d["e2e"]["value"], 1), d["gpu_launches"], round(d["roofline"]["frac"], 3), {k: round(v["value"], 1) for k, v in d["workloads"].items()})
PY
