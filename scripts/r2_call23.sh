# Round 2, twenty-third call (1 GPU): identity-row pass beside the gradient-row pass (second stream), CTAs per SM sweep
set -x
mkdir -p gpurun_out
for n in 0 1 2 3 4 6; do
PB_OVERLAP_IDENTITY=$n timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/r2c23_lifting_$n.json 2> gpurun_out/r2c23_lifting_$n.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c23_lifting_$n.json").read().strip().splitlines()[-1])
print("overlap $n", round(d["value"], 1), "iter/s", round(d["ms_per_step"], 3), "ms", d.get("hash"))
PY
done
PB_OVERLAP_IDENTITY=2 timeout 600 python -m pytest tests/test_reference_parity.py tests/test_gpu_pdhg.py -m gpu -q -k "lifting" > gpurun_out/r2c23_pytest.log 2>&1
tail -3 gpurun_out/r2c23_pytest.log | cut -c1-300
