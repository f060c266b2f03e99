# 8-GPU evidence, shortest form: BASELINE config 3 (lifting) and the metric config on 8 GPUs
N=8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 80 $TR --nproc-per-node $N --master-port 29583 scripts/bench_lifting.py --steps 40 --warmup 5 > gpurun_out/mg_lift_n8.json 2> gpurun_out/mg_lift_n8.err
timeout 80 $TR --nproc-per-node $N --master-port 29582 bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/mg_rof_n8.json 2> gpurun_out/mg_rof_n8.err
python - <<PY
import json
for f in ["gpurun_out/mg_lift_n8.json", "gpurun_out/mg_rof_n8.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), d.get("halo_mode"),
              "e2e", d.get("e2e", {}).get("value"), "ttr", d.get("time_to_residual_1e-4", {}).get("seconds"))
    except Exception as e:
        print("ERR", f, e)
PY
