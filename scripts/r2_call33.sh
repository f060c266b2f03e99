# Round 2, thirty-third call (1 GPU): ncu --set full of the two staged lifting kernels (plain iterations)
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"staged_kernel" -s 8 -c 2 -f -o gpurun_out/r2c33_lifting_staged python scripts/bench_lifting.py --steps 12 --warmup 2 > gpurun_out/r2c33_ncu.log 2>&1
tail -2 gpurun_out/r2c33_ncu.log | cut -c1-200
