"""Diagnostic (not a test): how far do float trajectories of the adaptive step-size schemes drift?
Compares reference-f32, reference-f64, this library and the CPU oracle pairwise."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import prost_b200 as pb
import ref_driver
from pdhg_util import rel_err, run_cuda, run_oracle
from prost_b200 import synthetic as syn

F64 = ref_driver.REF_DRIVER + "_f64"
TOL4 = dict(tol_rel_primal=1e-4, tol_rel_dual=1e-4, tol_abs_primal=1e-4, tol_abs_dual=1e-4)
CASES = {
    "rof_alg1": (lambda: syn.rof(48, 37), dict(stepsize="alg1", residual_iter=3)),
    "rof_alg2": (lambda: syn.rof(40, 36), dict(stepsize="alg2", residual_iter=3, alg2_gamma=0.5)),
    "rof_goldstein": (lambda: syn.rof(48, 37), dict(stepsize="goldstein", residual_iter=3)),
    "rof_boyd": (lambda: syn.rof(40, 36), dict(stepsize="boyd", residual_iter=3)),
    "tvl1_color": (lambda: syn.tvl1(64, 48, nc=3), dict(stepsize="boyd", residual_iter=10)),
}
ctx = pb.Context(0)
for name, (fn, opts) in CASES.items():
    desc = fn()
    for iters in (10, 30, 60, 100, 150, 200):
        r32 = ref_driver.run_solve(desc, iters, tol=TOL4, **opts)
        r64 = ref_driver.run_solve(desc, iters, tol=TOL4, binary=F64, **opts)
        mine = run_cuda(ctx, desc, iters, fuse=1, tol=TOL4, **opts)
        orc = run_oracle(desc, iters, tol=TOL4, **opts)
        print(f"{name:14s} it={iters:4d}  x: ref32-ref64 {rel_err(r32['x'], r64['x']):.2e}  mine-ref64 {rel_err(mine['x'], r64['x']):.2e}  "
              f"mine-ref32 {rel_err(mine['x'], r32['x']):.2e}  orc-ref64 {rel_err(orc['x'], r64['x']):.2e} | "
              f"res_p ref32 {r32['res']['primal_residual']:.6g} ref64 {r64['res']['primal_residual']:.6g} mine {mine['res']['primal_residual']:.6g} "
              f"steps mine {mine['steps']}", flush=True)
