# Round 2, eighteenth call (1 GPU): source-level profile of the warp-specialised tcgen05 Kronecker kernel
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kron_tc -s 24 -c 1 -o gpurun_out/r2c18_kron_tc_dense python scripts/bench_linops.py --reps 1 --only kron > gpurun_out/r2c18_ncu1.log 2>&1
tail -2 gpurun_out/r2c18_ncu1.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kron_tc -s 32 -c 1 -o gpurun_out/r2c18_kron_tc_idfirst python scripts/bench_linops.py --reps 1 --only kron > gpurun_out/r2c18_ncu2.log 2>&1
tail -2 gpurun_out/r2c18_ncu2.log | cut -c1-200
