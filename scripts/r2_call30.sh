# Round 2, thirtieth call (1 GPU): default bench after the reference_cuda timing change; eigen_nxn 12 x 12 against the reference
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2c30_bench.json 2> gpurun_out/r2c30_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c30_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], json.dumps(d["reference_cuda"])[:1200])
PY
timeout 300 python -m pytest tests/test_zz_gpu_next_rows.py tests/test_gpu_ops.py -m gpu -q -k "eigen_nxn or spectral" > gpurun_out/r2c30_pytest.log 2>&1
tail -3 gpurun_out/r2c30_pytest.log | cut -c1-300
