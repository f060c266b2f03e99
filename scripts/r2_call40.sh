# Round 2, fortieth call (1 GPU): cooperative staging in the refresh variants: bitwise check, rate, parity, launch list
set -x
mkdir -p gpurun_out
timeout 300 python scripts/check_staged_coop.py > gpurun_out/r2c40_check.log 2>&1
echo "rc $?"; tail -8 gpurun_out/r2c40_check.log | cut -c1-300
timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/r2c40_lifting.json 2> gpurun_out/r2c40_lifting.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c40_lifting.json").read().strip().splitlines()[-1])
print("lifting", round(d["value"], 1), "iter/s", round(d["ms_per_step"], 3), "ms")
PY
timeout 600 python -m pytest tests/test_reference_parity.py tests/test_gpu_pdhg.py -m gpu -q -k "lifting" > gpurun_out/r2c40_pytest.log 2>&1
tail -3 gpurun_out/r2c40_pytest.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2c40_lifting_launches.csv python scripts/bench_lifting.py --steps 12 --warmup 2 > gpurun_out/r2c40_launch.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2c40_lifting_launches.csv")) if len(r) > 5]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[k][:76]].append(float(r[v].replace(",", "")))
    except ValueError: pass
for name, xs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{name:78s} n={len(xs):3d} mean={sum(xs)/len(xs)/1e3:9.1f} us")
PY
