# Round 2, sixth call (2 GPUs): PDL + halo prefetch -- slab parity, timeline and rate at N = 1 and 2; lifting launch list
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_tile.py -m gpu -q -x > gpurun_out/r2c6_pytest.log 2>&1
tail -4 gpurun_out/r2c6_pytest.log
for pdl in 1 0; do
export PB_RING_PDL=$pdl
PB_RING_TRACE=1 timeout 120 python scripts/ring_trace.py 2>&1 | grep RING_TRACE | tee gpurun_out/r2c6_trace_n1_pdl$pdl.log
PB_RING_TRACE=1 timeout 200 $TR --nproc-per-node 2 --master-port 29631 scripts/ring_trace.py 2>&1 | grep RING_TRACE | tee gpurun_out/r2c6_trace_n2_pdl$pdl.log
PB_RING_TRACE=1 PB_TRACE_NX=1024 timeout 200 $TR --nproc-per-node 2 --master-port 29632 scripts/ring_trace.py 2>&1 | grep RING_TRACE | tee gpurun_out/r2c6_trace_n2_1024_pdl$pdl.log
done
unset PB_RING_PDL
timeout 300 $TR --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --steps 2000 --warmup 50 --no-workloads --no-cpu-baseline > gpurun_out/r2c6_bench_n2.json 2> gpurun_out/r2c6_bench_n2.err
timeout 300 python bench.py --steps 2000 --warmup 50 --no-workloads --no-cpu-baseline > gpurun_out/r2c6_bench_n1.json 2> gpurun_out/r2c6_bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/r2c6_bench_k20.json 2> gpurun_out/r2c6_bench_k20.err
python - <<PY
import json
for f in ("n2", "n1", "k20"):
    try:
        d = json.loads(open(f"gpurun_out/r2c6_bench_{f}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "tile ms", r["ms_per_launch"], "frac", round(r["frac"], 3),
              "check ms", r["residual_refresh_iterations"]["ms_per_launch"], "hash", d["iterate_hash"], "e2e", round(d["e2e"]["value"], 1), d["e2e"]["seconds"],
              "ttr", d["time_to_residual_1e-4"]["seconds"], d["time_to_residual_1e-4"]["iterations"])
    except Exception as e:
        print("ERR", f, e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2c6_lifting_launches.csv python scripts/bench_lifting.py --steps 12 --warmup 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2c6_lifting_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki][:110]].append(float(r[vi].replace(",", "")))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):4d} x {sum(v)/len(v)/1e3:9.1f} us  {k}")
PY
