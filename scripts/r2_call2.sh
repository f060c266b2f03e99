# Round 2, second call (1 GPU): whole GPU suite after the RingFinish / MULTI changes, multi-iteration check, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_pytest.log 2>&1
tail -6 gpurun_out/r2c2_pytest.log
timeout 600 python scripts/check_ring_multi.py > gpurun_out/r2c2_multi_n1.log 2>&1
tail -6 gpurun_out/r2c2_multi_n1.log
timeout 900 python bench.py --steps 2000 --warmup 50 > gpurun_out/r2c2_bench_n1.json 2> gpurun_out/r2c2_bench_n1.err
tail -3 gpurun_out/r2c2_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c2_bench_n1.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "ms/step", d["ms_per_step"], "launches", d["gpu_launches"])
print("roofline", r["frac"], r["ms_per_launch"], "check", r["residual_refresh_iterations"]["ms_per_launch"], r["residual_refresh_iterations"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e"]["seconds"])
print("ref_cuda", d.get("reference_cuda"))
print("cpu", d.get("cpu_baseline"))
for k, v in d["workloads"].items():
    print(k, v.get("value"), v.get("roofline", {}).get("frac"), v.get("error"))
print("ttr", d["time_to_residual_1e-4"]["seconds"], d["time_to_residual_1e-4"]["iterations"], "hash", d["iterate_hash"])
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/r2c2_bench_k20.json 2> gpurun_out/r2c2_bench_k20.err
python -c "
import json
d = json.loads(open('gpurun_out/r2c2_bench_k20.json').read().strip().splitlines()[-1])
print('K=20 e2e', d['e2e'])
"
