# Round 2, fourteenth call (1 GPU): tcgen05 Kronecker path bring-up
set -x
mkdir -p gpurun_out
for sw in 0 1; do
  PB_KRON_TC_SWAP=$sw timeout 120 python scripts/check_kron_tc.py > gpurun_out/r2c14_check_swap$sw.log 2>&1
  echo "swap $sw rc $?"; tail -12 gpurun_out/r2c14_check_swap$sw.log | cut -c1-200
done
timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c14_linops.json 2> gpurun_out/r2c14_linops.err
tail -2 gpurun_out/r2c14_linops.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c14_linops.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    if "dense" in k:
        print(f"{k:60s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
PB_KRON_TC=0 timeout 300 python scripts/bench_linops.py --reps 20 --only kron > gpurun_out/r2c14_linops_fp32.json 2> gpurun_out/r2c14_linops_fp32.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c14_linops_fp32.json").read().strip().splitlines()[-1])
for k, v in d["ops"].items():
    if "dense" in k:
        print(f"fp32 {k:60s} {v['ms']*1e3:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_hbm_peak']:.3f}")
PY
