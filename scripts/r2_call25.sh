# Round 2, twenty-fifth call (1 GPU): CSR SpMV with 2 x kUnroll gathers in flight per thread
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_reference_parity.py tests/test_gpu_admm.py -m gpu -q -k "linop or sparse or admm or lasso" > gpurun_out/r2c25_pytest.log 2>&1
tail -4 gpurun_out/r2c25_pytest.log | cut -c1-300
for v in 0 1 2; do
PB_SPMV_VARIANT=$v timeout 300 python scripts/bench_linops.py --reps 20 --only sparse > gpurun_out/r2c25_linops_v$v.json 2> gpurun_out/r2c25_linops_v$v.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c25_linops_v$v.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    print(f"variant $v {k:50s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
done
timeout 300 python scripts/bench_admm.py > gpurun_out/r2c25_admm.json 2> gpurun_out/r2c25_admm.err
tail -1 gpurun_out/r2c25_admm.json | cut -c1-500
