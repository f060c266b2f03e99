# evidence run for profiles/ (one GPU, under gpurun): full GPU test suite, default bench, ncu launch lists
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/l_pytest.log
timeout 300 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/l_ref.json 2> gpurun_out/l_ref.err
timeout 120 python scripts/e2e_breakdown.py 200 > gpurun_out/l_e2e.json 2> gpurun_out/l_e2e.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/l_ncu_bench.log 2>&1
timeout 400 python scripts/bench_admm.py > gpurun_out/l_admm.json 2> gpurun_out/l_admm.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/r01_admm_launches.csv python scripts/bench_admm.py --m 1048576 --n 262144 --dense 2048 --iters 8 > gpurun_out/l_ncu_admm.log 2>&1
tail -4 gpurun_out/l_pytest.log; cat gpurun_out/l_e2e.json; cat gpurun_out/l_admm.json; tail -2 gpurun_out/l_admm.err
