# one-pass iteration kernel experiments (run under gpurun, one GPU): persistent TMA ring (2), one-shot TMA (1), plain (0)
set -x
mkdir -p gpurun_out
for m in 2; do
  PB_TILE_MODE=$m timeout 120 python -m pytest tests/test_gpu_tile.py -q 2>&1 | tail -8
done
for m in 2; do
  PB_TILE_MODE=$m timeout 200 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/g_mode$m.json 2> gpurun_out/g_mode$m.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grad2d_iteration -s 5 -c 1 -f -o gpurun_out/r01_tile_ring python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/g_ncu.log 2>&1
for f in gpurun_out/g_mode*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[1], "value", round(d["value"],1), "tile ms", round(r["ms_per_launch"],4), "frac", round(r["frac"],3), "e2e", round(d["e2e"]["value"],1), r["kernel"][:40])
PY
done
tail -3 gpurun_out/g_mode2.err
