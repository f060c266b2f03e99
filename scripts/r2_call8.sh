# Round 2, eighth call (2 GPUs): row-sharded ADMM parity on 2 ranks; slab + ADMM suites after the cross-sum refactor
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_admm_sharded.py tests/test_gpu_slab.py tests/test_gpu_admm.py -m gpu -q > gpurun_out/r2c8_pytest.log 2>&1
tail -12 gpurun_out/r2c8_pytest.log
