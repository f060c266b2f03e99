# Round 2, forty-third call (1 GPU): smoke() on the final build
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c43_smoke.log 2>&1
tail -3 gpurun_out/r2c43_smoke.log | cut -c1-300
