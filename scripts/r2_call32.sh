# Round 2, thirty-second call (1 GPU): kron(K, I) results through per-warp TMA tensor stores
set -x
mkdir -p gpurun_out
timeout 120 python scripts/check_kron_tc.py > gpurun_out/r2c32_check.log 2>&1
echo "rc $?"; tail -12 gpurun_out/r2c32_check.log | cut -c1-200
for t in 1 0; do
PB_KRON_TC_TMA_OUT=$t timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c32_linops_$t.json 2> gpurun_out/r2c32_linops_$t.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c32_linops_$t.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    if "dense_kron_id" in k:
        print(f"tma_out $t {k:50s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
done
