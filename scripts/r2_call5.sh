# Round 2, fifth call (2 GPUs): ring-kernel launch timeline at N = 1 (full and half problem) and N = 2; chunked staged
# lifting kernels: parity tests + rate
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export PB_RING_TRACE=1
timeout 120 python scripts/ring_trace.py 2>&1 | grep RING_TRACE | tee gpurun_out/r2c5_trace_n1_4096.log
PB_TRACE_NX=2048 timeout 120 python scripts/ring_trace.py 2>&1 | grep RING_TRACE | tee gpurun_out/r2c5_trace_n1_2048.log
PB_TRACE_NX=512 timeout 120 python scripts/ring_trace.py 2>&1 | grep RING_TRACE | tee gpurun_out/r2c5_trace_n1_512.log
timeout 200 $TR --nproc-per-node 2 --master-port 29631 scripts/ring_trace.py 2>&1 | grep RING_TRACE | tee gpurun_out/r2c5_trace_n2.log
PB_TRACE_NX=1024 timeout 200 $TR --nproc-per-node 2 --master-port 29632 scripts/ring_trace.py 2>&1 | grep RING_TRACE | tee gpurun_out/r2c5_trace_n2_1024.log
unset PB_RING_TRACE
timeout 600 python -m pytest tests/test_gpu_pdhg.py tests/test_reference_parity.py tests/test_zz_gpu_next_rows.py -m gpu -q -k "lifting or ind_sum_indexed or dual_problem" > gpurun_out/r2c5_pytest.log 2>&1
tail -5 gpurun_out/r2c5_pytest.log
timeout 300 python scripts/bench_lifting.py --steps 60 --warmup 5 > gpurun_out/r2c5_lifting.json 2> gpurun_out/r2c5_lifting.err
python -c "
import json
d = json.loads(open('gpurun_out/r2c5_lifting.json').read().strip().splitlines()[-1])
print('lifting', d['value'], d['ms_per_step'], d['roofline']['frac'])
"
