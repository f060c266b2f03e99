# Round 2, sixteenth call (1 GPU): where does the tcgen05 Kronecker kernel spend its time
set -x
mkdir -p gpurun_out
for dbg in 0 1 2 4 8 3 15; do
PB_KRON_TC_DEBUG=$dbg timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c16_dbg$dbg.json 2> gpurun_out/r2c16_dbg$dbg.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c16_dbg$dbg.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    if "64x64_d1048576:forward" in k or "16x32_d1048576:forward" in k and "dense" in k:
        print(f"debug $dbg {k:50s} {v['ms']*1e3:9.1f} us")
PY
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kron_tc -c 6 -o gpurun_out/r2c16_kron_tc python scripts/bench_linops.py --reps 1 --only kron > gpurun_out/r2c16_ncu.log 2>&1
tail -3 gpurun_out/r2c16_ncu.log
