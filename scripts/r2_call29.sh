# Round 2, twenty-ninth call (1 GPU): whole GPU suite, new fixtures, default bench + reference arm, bench launch list
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r2c29_pytest.log
tail -6 gpurun_out/r2c29_pytest.log | cut -c1-300
timeout 300 python tests/golden/make_golden.py gpurun_out/golden_new5 new_prox > gpurun_out/r2c29_golden.log 2>&1
ls gpurun_out/golden_new5 | wc -l
timeout 600 python bench.py > gpurun_out/r2c29_bench.json 2> gpurun_out/r2c29_bench.err
tail -c 3000 gpurun_out/r2c29_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2c29_ref.json 2> gpurun_out/r2c29_ref.err
tail -c 800 gpurun_out/r2c29_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 100 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/r2c29_ncu_bench.log 2>&1
tail -2 gpurun_out/r2c29_ncu_bench.log | cut -c1-300
