# Round 2, forty-first call (1 GPU): whole GPU suite and the default bench on the final build
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r2c41_pytest.log
tail -6 gpurun_out/r2c41_pytest.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2c41_bench.json 2> gpurun_out/r2c41_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c41_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["value"], 1) for k, v in d["workloads"].items()}, d["reference_cuda"].get("speedup_iterations"), d["reference_cuda"].get("parity_ok"))
PY
