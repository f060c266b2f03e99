# 1-GPU check of the host-side fast paths: scaling segments, staged host copies, prefault; PB_TRACE breakdown
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_pdhg.py tests/test_gpu_ops.py tests/test_gpu_admm.py tests/test_gpu_tile.py -q -x 2>&1 | tail -15) > gpurun_out/e_pytest.log
tail -6 gpurun_out/e_pytest.log
PB_TRACE=1 timeout 200 python scripts/e2e_breakdown.py 200 > gpurun_out/e_e2e_trace.json 2> gpurun_out/e_e2e_trace.err
cat gpurun_out/e_e2e_trace.err | tail -40
timeout 200 python scripts/e2e_breakdown.py 2000 > gpurun_out/e_e2e.json 2> gpurun_out/e_e2e.err
cat gpurun_out/e_e2e.json
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/e_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", json.dumps(d["e2e"]))
PY
