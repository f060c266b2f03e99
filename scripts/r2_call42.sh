# Round 2, forty-second call (2 GPUs): the driver's N = 2 bench command on the final build
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 > gpurun_out/r2c42_bench_n2.json 2> gpurun_out/r2c42_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c42_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d.get("iterate_hash"), {k: round(v["value"], 1) for k, v in d.get("workloads", {}).items()})
PY
tail -3 gpurun_out/r2c42_bench_n2.err | cut -c1-300
