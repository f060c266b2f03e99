# Round 2, first 1-GPU call (under gpurun): everything that was prepared without a GPU at the end of round 1.
#  1. bring-up of the multi-iteration ring launches (PB_RING_ITERS), bitwise vs single launches + rate
#  2. BASELINE.md line A: unmodified reference CUDA solver vs this library through the same C++ driver
#  3. source-level ncu captures of the slow lifting kernels (residual-refresh variants of the staged passes and
#     the identity-row epigraph pass): top source lines by warp-stall samples
set -x
mkdir -p gpurun_out
timeout 600 python scripts/check_ring_multi.py > gpurun_out/r2_ring_multi.log 2>&1
tail -8 gpurun_out/r2_ring_multi.log
timeout 900 python scripts/bench_reference_cuda.py --iters 300 > gpurun_out/r2_reference_cuda.json 2> gpurun_out/r2_reference_cuda.err
cat gpurun_out/r2_reference_cuda.json; tail -2 gpurun_out/r2_reference_cuda.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"staged_kernel|prox_pass_kernel" -s 27 -c 6 -f \
  -o gpurun_out/r2_lifting_check python scripts/bench_lifting.py --steps 22 --warmup 2 > gpurun_out/r2_lifting_ncu.log 2>&1
ncu -i gpurun_out/r2_lifting_check.ncu-rep --page raw --csv > gpurun_out/r2_lifting_check_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_lifting_check.ncu-rep --page source --csv > gpurun_out/r2_lifting_check_source.csv 2>/dev/null
python - <<'PY'
# keep the source page small: the 40 lines with the most stall samples per kernel
import csv, collections
try:
    rows = list(csv.reader(open("gpurun_out/r2_lifting_check_source.csv")))
    hdr = next(r for r in rows if "Source" in r)
    si, wi = hdr.index("Source"), next(i for i, h in enumerate(hdr) if "Warp Stall Sampling (All" in h)
    best = sorted((r for r in rows if len(r) > wi and r[wi].replace(",", "").isdigit()),
                  key=lambda r: -int(r[wi].replace(",", "")))[:120]
    with open("gpurun_out/r2_lifting_check_source_top.txt", "w") as f:
        for r in best:
            f.write(f"{r[wi]:>10}  {r[si][:160]}\n")
except Exception as e:
    print("source page summary failed:", e)
PY
rm -f gpurun_out/r2_lifting_check.ncu-rep gpurun_out/r2_lifting_check_source.csv
#  4. golden fixtures of the 8(f) row-2 proxes (ProxTransform, ind_sum, ind_halfspace, ind_soc) from the live reference;
#     copy gpurun_out/golden_new/*.npz into tests/golden/ afterwards
timeout 300 python tests/golden/make_golden.py gpurun_out/golden_new new_prox > gpurun_out/r2_golden.log 2>&1
ls gpurun_out/golden_new | wc -l
