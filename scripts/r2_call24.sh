# Round 2, twenty-fourth call (1 GPU): identity pass beside the gradient pass, larger CTA counts
set -x
mkdir -p gpurun_out
for n in 8 10 12 16; do
PB_OVERLAP_IDENTITY=$n timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/r2c24_lifting_$n.json 2> gpurun_out/r2c24_lifting_$n.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c24_lifting_$n.json").read().strip().splitlines()[-1])
print("overlap $n", round(d["value"], 1), "iter/s", round(d["ms_per_step"], 3), "ms")
PY
done
