# Round 2, eleventh call (1 GPU): whole GPU suite after the operator / Kronecker rewrite; operator timings
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c11_pytest.log 2>&1
tail -6 gpurun_out/r2c11_pytest.log | cut -c1-300
timeout 600 python scripts/bench_linops.py --reps 20 --only dense,kron > gpurun_out/r2c11_linops.json 2> gpurun_out/r2c11_linops.err
tail -2 gpurun_out/r2c11_linops.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c11_linops.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    print(f"{k:60s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
