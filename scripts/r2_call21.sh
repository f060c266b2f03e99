# Round 2, twenty-first call (1 GPU): kron(I, K) output through per-warp shared-memory tiles
set -x
mkdir -p gpurun_out
timeout 120 python scripts/check_kron_tc.py > gpurun_out/r2c21_check.log 2>&1
echo "rc $?"; tail -12 gpurun_out/r2c21_check.log | cut -c1-200
timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c21_linops.json 2> gpurun_out/r2c21_linops.err
tail -2 gpurun_out/r2c21_linops.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c21_linops.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    if "dense" in k:
        print(f"{k:60s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
timeout 600 python -m pytest tests/test_zz_gpu_next_rows.py -m gpu -q -k "kron" > gpurun_out/r2c21_pytest.log 2>&1
tail -3 gpurun_out/r2c21_pytest.log | cut -c1-300
