# evidence run for profiles/ (one GPU, under gpurun): full GPU test suite, default bench, reference arm, ncu launch
# lists of the bench and of the lifting config, one full capture of each lifting pass
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/l_pytest.log
tail -4 gpurun_out/l_pytest.log
timeout 300 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/l_ref.json 2> gpurun_out/l_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/l_ncu_bench.log 2>&1
timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/lift_n1.json 2> gpurun_out/lift_n1.err
timeout 300 ncu --set full --clock-control none -k regex:"staged_kernel|prox_pass_kernel" -s 9 -c 3 -f -o gpurun_out/r01_lifting_staged_full python scripts/bench_lifting.py --steps 12 --warmup 2 > gpurun_out/lp_full.log 2>&1
ncu -i gpurun_out/r01_lifting_staged_full.ncu-rep --page raw --csv > gpurun_out/r01_lifting_staged_full_raw.csv 2> gpurun_out/lp_export.err
rm -f gpurun_out/r01_lifting_staged_full.ncu-rep
python - <<'PY'
import json
d=json.loads(open("gpurun_out/l_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", json.dumps(d["e2e"]), "ttr", d["time_to_residual_1e-4"]["seconds"], d["time_to_residual_1e-4"]["iterations"])
print(open("gpurun_out/lift_n1.json").read()[:400])
PY
