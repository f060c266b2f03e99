# Round 2, fifteenth call (1 GPU): tcgen05 Kronecker path, K-major for both products, prefetch depth sweep
set -x
mkdir -p gpurun_out
timeout 120 python scripts/check_kron_tc.py > gpurun_out/r2c15_check.log 2>&1
echo "rc $?"; tail -12 gpurun_out/r2c15_check.log | cut -c1-200
for depth in 1 2 3; do
PB_KRON_TC_DEPTH=$depth timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c15_linops_d$depth.json 2> gpurun_out/r2c15_linops_d$depth.err
tail -2 gpurun_out/r2c15_linops_d$depth.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c15_linops_d$depth.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    if "dense" in k:
        print(f"depth $depth {k:50s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
done
