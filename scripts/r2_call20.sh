# Round 2, twentieth call (1 GPU): Kronecker tests (fp32 and tcgen05 paths) and the operator bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_next_rows.py -m gpu -q -k "kron" > gpurun_out/r2c20_pytest.log 2>&1
tail -6 gpurun_out/r2c20_pytest.log | cut -c1-300
timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c20_linops.json 2> gpurun_out/r2c20_linops.err
tail -2 gpurun_out/r2c20_linops.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c20_linops.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    print(f"{k:60s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kron_tc -s 24 -c 1 -o gpurun_out/r2c20_kron_tc_dense python scripts/bench_linops.py --reps 1 --only kron > gpurun_out/r2c20_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kron_tc -s 32 -c 1 -o gpurun_out/r2c20_kron_tc_idfirst python scripts/bench_linops.py --reps 1 --only kron > gpurun_out/r2c20_ncu2.log 2>&1
