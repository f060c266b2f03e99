# Round 2, twenty-second call (1 GPU): epigraph projection with cbrtf / t sqrt(t) / exact reciprocals: parity + lifting rate
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_reference_parity.py tests/test_gpu_pdhg.py -m gpu -q -k "epi_quad or lifting or prox_matches" > gpurun_out/r2c22_pytest.log 2>&1
tail -6 gpurun_out/r2c22_pytest.log | cut -c1-300
timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/r2c22_lifting.json 2> gpurun_out/r2c22_lifting.err
tail -1 gpurun_out/r2c22_lifting.json | cut -c1-600
timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 100 > gpurun_out/r2c22_lifting_w100.json 2> gpurun_out/r2c22_lifting_w100.err
tail -1 gpurun_out/r2c22_lifting_w100.json | cut -c1-600
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2c22_lifting_launches.csv python scripts/bench_lifting.py --steps 12 --warmup 2 > gpurun_out/r2c22_launch.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2c22_lifting_launches.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[k][:70]].append(float(r[v].replace(",", "")))
    except ValueError: pass
for name, xs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name:72s} n={len(xs):3d} mean={sum(xs)/len(xs)/1e3:9.1f} us")
PY
