# 1-GPU: lifting-type problems through the staged kernels (parity vs oracle / live reference / single-vs-slab),
# then BASELINE config 3 at N=1
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_pdhg.py tests/test_gpu_slab.py tests/test_reference_parity.py -q -x -k "lifting or scaling or baseline_configs or world1 or execution_modes" 2>&1 | tail -8) > gpurun_out/f_pytest.log
tail -5 gpurun_out/f_pytest.log
timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/lift_n1.json 2> gpurun_out/lift_n1.err
tail -3 gpurun_out/lift_n1.err; cat gpurun_out/lift_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r01_lifting_launches_staged.csv python scripts/bench_lifting.py --steps 12 --warmup 2 > gpurun_out/lp_launch.log 2>&1
