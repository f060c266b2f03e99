# Round 2, forty-fourth call (4 GPUs): slab cases (cooperative lifting staging included) on 4 ranks against one GPU
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_slab.py -m gpu -q -k all_gpus > gpurun_out/r2c44_pytest.log 2>&1
tail -4 gpurun_out/r2c44_pytest.log | cut -c1-300
