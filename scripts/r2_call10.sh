# Round 2, tenth call (8 GPUs): scaling of the metric config and of configs 2-4 at N = 8, 4, 2; slab parity at 8 ranks;
# row-sharded ADMM at 4 ranks
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
timeout 400 $TR --nproc-per-node $n --master-port 2962$n bench.py --gpus $n --steps 2000 --warmup 50 --no-cpu-baseline > gpurun_out/r2c10_bench_n$n.json 2> gpurun_out/r2c10_bench_n$n.err
tail -2 gpurun_out/r2c10_bench_n$n.err | cut -c1-300
done
timeout 300 python -m pytest tests/test_gpu_slab.py tests/test_gpu_admm_sharded.py -m gpu -q > gpurun_out/r2c10_pytest.log 2>&1
tail -4 gpurun_out/r2c10_pytest.log
python - <<'PY'
import json
for n in (8, 4, 2):
    try:
        d = json.loads(open(f"gpurun_out/r2c10_bench_n{n}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "tile ms", round(r["ms_per_launch"], 5),
              "check ms", round(r["residual_refresh_iterations"]["ms_per_launch"], 5), "hash", d["iterate_hash"], "e2e", round(d["e2e"]["value"], 1),
              "ttr", round(d["time_to_residual_1e-4"]["seconds"], 4), d["time_to_residual_1e-4"]["iterations"], d["halo_mode"])
        for k, v in d["workloads"].items():
            print("   ", k, v.get("value"), v.get("roofline", {}).get("frac"), v.get("error"))
    except Exception as e:
        print("ERR", n, e)
PY
