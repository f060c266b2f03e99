# Round 2, third call (2 GPUs): remaining tests, MULTI variants on 1 GPU, slab parity + MULTI + bench on 2 GPUs
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_reference_parity.py tests/test_zz_gpu_next_rows.py tests/test_gpu_slab.py -m gpu -q > gpurun_out/r2c3_pytest.log 2>&1
tail -8 gpurun_out/r2c3_pytest.log
PB_CHECK_DEBUG_VARIANTS=1 timeout 600 python scripts/check_ring_multi.py > gpurun_out/r2c3_multi_n1.log 2>&1
tail -14 gpurun_out/r2c3_multi_n1.log
timeout 600 $TR --nproc-per-node 2 --master-port 29611 scripts/check_ring_multi.py > gpurun_out/r2c3_multi_n2.log 2>&1
tail -14 gpurun_out/r2c3_multi_n2.log
for mode in default coarse single; do
  case $mode in
    default) export PB_RING_ITERS=16 PB_RING_COARSE=0;;
    coarse) export PB_RING_ITERS=16 PB_RING_COARSE=1;;
    single) export PB_RING_ITERS=1 PB_RING_COARSE=0;;
  esac
  timeout 600 $TR --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --steps 2000 --warmup 50 --no-workloads --no-cpu-baseline > gpurun_out/r2c3_bench_n2_$mode.json 2> gpurun_out/r2c3_bench_n2_$mode.err
  tail -2 gpurun_out/r2c3_bench_n2_$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c3_bench_n2_$mode.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$mode", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "launches", d["gpu_launches"], "tile ms", r["ms_per_launch"],
          "check ms", r["residual_refresh_iterations"]["ms_per_launch"], "hash", d["iterate_hash"], "e2e", round(d["e2e"]["value"], 1), "ttr", d["time_to_residual_1e-4"]["seconds"])
except Exception as e:
    print("ERR $mode", e)
PY
done
