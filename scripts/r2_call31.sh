# Round 2, thirty-first call (1 GPU): tcgen05 Kronecker kernel, drain lag 1 / 2 x register prefetch depth 1 / 2
set -x
mkdir -p gpurun_out
timeout 120 python scripts/check_kron_tc.py > gpurun_out/r2c31_check.log 2>&1
echo "rc $?"; tail -12 gpurun_out/r2c31_check.log | cut -c1-200
for cfg in "2 1" "2 2" "1 1" "1 2"; do
set -- $cfg
PB_KRON_TC_LAG=$1 PB_KRON_TC_DEPTH=$2 timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c31_linops_$1_$2.json 2> gpurun_out/r2c31_linops_$1_$2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c31_linops_$1_$2.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    if "dense" in k:
        print(f"lag $1 depth $2 {k:50s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
done
