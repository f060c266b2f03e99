# Round 2, seventh call (1 GPU): operator-apply evidence -- timings, ncu launch list and full captures for the CSR
# SpMV / dense GEMV / Kronecker kernels; ADMM launch list; ring kernel traffic capture; ind_sum vs live reference
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zz_gpu_next_rows.py -m gpu -q -k "ind_sum_indexed" 2>&1 | tail -3
timeout 600 python scripts/bench_linops.py --reps 20 > gpurun_out/r2c7_linops.json 2> gpurun_out/r2c7_linops.err
tail -2 gpurun_out/r2c7_linops.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c7_linops.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    print(f"{k:60s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"csr_spmv|dense_gemv|kron" -c 24 -f -o gpurun_out/r2c7_linops python scripts/bench_linops.py --reps 1 > /dev/null 2>&1
ncu -i gpurun_out/r2c7_linops.ncu-rep --page raw --csv > gpurun_out/r2c7_linops_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2c7_linops.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2c7_admm_launches.csv python scripts/bench_admm.py --iters 4 > gpurun_out/r2c7_admm_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2c7_admm_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki][:100]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):4d} x {sum(v)/len(v)/1e3:9.1f} us  {100*sum(v)/tot:5.1f} %  {k}")
PY
timeout 300 python scripts/bench_admm.py --iters 20 > gpurun_out/r2c7_admm.json 2> gpurun_out/r2c7_admm.err
cat gpurun_out/r2c7_admm.json | cut -c1-1500
# ring kernel DRAM traffic (roofline.traffic): one full capture of a steady-state launch
timeout 300 ncu --set full --clock-control none -k regex:"grad2d_iteration_ring_kernel" -s 30 -c 2 -f -o gpurun_out/r2c7_ring python bench.py --steps 60 --warmup 5 --no-workloads --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/r2c7_ring.ncu-rep --page raw --csv > gpurun_out/r2c7_ring_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2c7_ring.ncu-rep
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2c7_ring_full_raw.csv")))
h = rows[0]
def col(name):
    return next(i for i, c in enumerate(h) if c == name)
for r in rows[2:]:
    try:
        print(r[col("Kernel Name")][:60], "dur", r[col("gpu__time_duration.sum")], "rd", r[col("dram__bytes_read.sum")], "wr", r[col("dram__bytes_write.sum")])
    except Exception as e:
        print("ERR", e)
print(rows[1][col("dram__bytes_read.sum")], rows[1][col("gpu__time_duration.sum")])
PY
