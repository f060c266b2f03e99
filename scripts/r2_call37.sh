# Round 2, thirty-seventh call (2 GPUs): slab tests with the cooperative-staging lifting cases
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q > gpurun_out/r2c37_pytest.log 2>&1
tail -5 gpurun_out/r2c37_pytest.log | cut -c1-300
