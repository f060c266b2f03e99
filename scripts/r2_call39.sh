# Round 2, thirty-ninth call (1 GPU): identity-row pass with four pairs per thread: bitwise check, rate, parity
set -x
mkdir -p gpurun_out
timeout 300 python scripts/check_staged_coop.py > gpurun_out/r2c39_check.log 2>&1
echo "rc $?"; tail -8 gpurun_out/r2c39_check.log | cut -c1-300
for c in 1 0; do
PB_IDENT_VEC4=$c timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/r2c39_lifting_$c.json 2> gpurun_out/r2c39_lifting_$c.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c39_lifting_$c.json").read().strip().splitlines()[-1])
print("vec4 $c", round(d["value"], 1), "iter/s", round(d["ms_per_step"], 3), "ms")
PY
done
timeout 600 python -m pytest tests/test_reference_parity.py tests/test_gpu_pdhg.py -m gpu -q -k "lifting" > gpurun_out/r2c39_pytest.log 2>&1
tail -3 gpurun_out/r2c39_pytest.log | cut -c1-300
