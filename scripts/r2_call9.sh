# Round 2, ninth call (1 GPU): new proxes against oracle / live reference; golden fixtures for them
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_next_rows.py -m gpu -q > gpurun_out/r2c9_pytest.log 2>&1
tail -15 gpurun_out/r2c9_pytest.log | cut -c1-300
timeout 300 python tests/golden/make_golden.py gpurun_out/golden_new2 new_prox > gpurun_out/r2c9_golden.log 2>&1
ls gpurun_out/golden_new2 | wc -l; tail -3 gpurun_out/r2c9_golden.log
