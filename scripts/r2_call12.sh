# Round 2, twelfth call (1 GPU): mass / comass norms vs oracle and live reference, Kronecker kernels (warp tiles), fixtures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_next_rows.py -m gpu -q -k "spectral or kron" > gpurun_out/r2c12_pytest.log 2>&1
tail -8 gpurun_out/r2c12_pytest.log | cut -c1-300
timeout 600 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c12_linops.json 2> gpurun_out/r2c12_linops.err
tail -2 gpurun_out/r2c12_linops.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c12_linops.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    print(f"{k:60s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
timeout 300 python tests/golden/make_golden.py gpurun_out/golden_new3 new_prox > gpurun_out/r2c12_golden.log 2>&1
ls gpurun_out/golden_new3 | wc -l; tail -3 gpurun_out/r2c12_golden.log
