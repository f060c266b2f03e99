# Round 2, thirty-eighth call (1 GPU): launch list of the lifting config with cooperative staging
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2c38_lifting_launches.csv python scripts/bench_lifting.py --steps 12 --warmup 2 > gpurun_out/r2c38_launch.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2c38_lifting_launches.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[k][:76]].append(float(r[v].replace(",", "")))
    except ValueError: pass
for name, xs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name:78s} n={len(xs):3d} mean={sum(xs)/len(xs)/1e3:9.1f} us")
PY
