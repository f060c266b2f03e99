# Round 2, thirteenth call (1 GPU): ind_range vs closed form / oracle / live reference, fixture
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_next_rows.py -m gpu -q -k "ind_range" > gpurun_out/r2c13_pytest.log 2>&1
tail -8 gpurun_out/r2c13_pytest.log | cut -c1-300
timeout 300 python tests/golden/make_golden.py gpurun_out/golden_new4 new_prox > gpurun_out/r2c13_golden.log 2>&1
ls gpurun_out/golden_new4 | wc -l; tail -3 gpurun_out/r2c13_golden.log
