# Round 2, thirty-fourth call (1 GPU): RingMulti bitwise test, Kronecker tests with the TMA tensor stores
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ring_multi.py tests/test_zz_gpu_next_rows.py -m gpu -q -k "multi or kron" > gpurun_out/r2c34_pytest.log 2>&1
tail -5 gpurun_out/r2c34_pytest.log | cut -c1-400
