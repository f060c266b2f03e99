# Round 2, nineteenth call (1 GPU): tcgen05 Kronecker kernel with staged kron(I, K) output and lean index math
set -x
mkdir -p gpurun_out
timeout 120 python scripts/check_kron_tc.py > gpurun_out/r2c19_check.log 2>&1
echo "rc $?"; tail -12 gpurun_out/r2c19_check.log | cut -c1-200
for cfg in "2 0" "1 0" "2 2" "2 15"; do
set -- $cfg
PB_KRON_TC_DEPTH=$1 PB_KRON_TC_DEBUG=$2 timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c19_linops_$1_$2.json 2> gpurun_out/r2c19_linops_$1_$2.err
tail -2 gpurun_out/r2c19_linops_$1_$2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c19_linops_$1_$2.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    if "dense" in k:
        print(f"depth $1 debug $2 {k:50s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
done
