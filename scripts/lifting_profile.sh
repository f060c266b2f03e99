# 1-GPU ncu evidence for BASELINE config 3 (lifting 2048^2 x 32): launch list + one full capture of each pass
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r01_lifting_launches.csv python scripts/bench_lifting.py --steps 12 --warmup 2 > gpurun_out/lp_launch.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"grad_primal_kernel|grad_dual_norm2_kernel|prox_pass_kernel" -s 9 -c 3 -f -o gpurun_out/r01_lifting_full python scripts/bench_lifting.py --steps 12 --warmup 2 > gpurun_out/lp_full.log 2>&1
ncu -i gpurun_out/r01_lifting_full.ncu-rep --page raw --csv > gpurun_out/r01_lifting_full_raw.csv 2> gpurun_out/lp_export.err
rm -f gpurun_out/r01_lifting_full.ncu-rep
tail -3 gpurun_out/lp_launch.log gpurun_out/lp_full.log
