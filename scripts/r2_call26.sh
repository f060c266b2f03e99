# Round 2, twenty-sixth call (1 GPU): SpMV variants (rows per group step, lanes per row)
set -x
mkdir -p gpurun_out
for v in 0 1 2 3; do
PB_SPMV_VARIANT=$v timeout 300 python scripts/bench_linops.py --reps 20 --only sparse > gpurun_out/r2c26_linops_v$v.json 2> gpurun_out/r2c26_linops_v$v.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c26_linops_v$v.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    print(f"variant $v {k:50s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
done
