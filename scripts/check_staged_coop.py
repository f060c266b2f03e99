#!/usr/bin/env python
"""Cooperative staging of the lifting passes (PB_STAGED_COOP, pb_stencil_staged.cuh) and the four-pairs-per-thread
identity-row pass (PB_IDENT_VEC4, prox_pass_pairs4_kernel) against the per-thread / per-pair forms: the arithmetic
is identical, so x, y, z, w and the residuals must match bit for bit.  The switch is read once per
process, so the script re-executes itself with both settings (tests/test_gpu_staged_coop.py runs it)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = [("l32_5x128", (5, 128, 32), "boyd", 10, 57), ("l8_6x64", (6, 64, 8), "alg1", 7, 64),
         ("l16_3x192", (3, 192, 16), "goldstein", 4, 41), ("l4_9x64", (9, 64, 4), "boyd", 5, 33)]


def worker(tag, out_dir):
    import numpy as np
    import prost_b200 as pb
    from prost_b200 import synthetic as syn
    ctx = pb.Context(0)
    for name, (nx, ny, L), step, res_iter, iters in CASES:
        prob = pb.create_problem(ctx, syn.lifting(nx, ny, L))
        popts = pb.pdhg_options(scale_steps_operator=0, stepsize=step, residual_iter=res_iter)
        sopts = pb.solver_options(verbose=0, max_iters=iters, num_cback_calls=0, tol_rel_primal=0, tol_rel_dual=0,
                                  tol_abs_primal=0, tol_abs_dual=0)
        be = pb.BackendPDHG(ctx, prob, popts, sopts)
        prob.Initialize()
        be.Initialize()
        be.PerformIteration(iters)
        x, z, y, w = be.current_solution()
        np.savez(os.path.join(out_dir, f"{name}_{tag}.npz"), x=x, y=y, z=z, w=w,
                 res=np.array(list(be.residuals().values())))
        del be, prob


def main():
    if os.environ.get("PB_COOP_CHECK_DIR"):
        return worker(os.environ["PB_COOP_CHECK_TAG"], os.environ["PB_COOP_CHECK_DIR"])
    import tempfile
    import numpy as np
    tmp = tempfile.mkdtemp(prefix="staged_coop_")
    for tag in ("0", "1"):
        env = dict(os.environ, PB_STAGED_COOP=tag, PB_IDENT_VEC4=tag, PB_COOP_CHECK_DIR=tmp, PB_COOP_CHECK_TAG=tag)
        p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=600)
        if p.returncode != 0:
            print(tag, p.stdout[-2000:], p.stderr[-2000:])
            sys.exit(1)
    ok = True
    for name, *_ in CASES:
        a, b = np.load(os.path.join(tmp, f"{name}_0.npz")), np.load(os.path.join(tmp, f"{name}_1.npz"))
        same = all(np.array_equal(a[k], b[k]) for k in ("x", "y", "z", "w", "res"))
        ok &= same
        print(f"{name}: {'bit-identical' if same else 'DIFFERENT'}", flush=True)
        if not same:
            for k in ("x", "y"):
                d = np.abs(a[k] - b[k])
                print("   ", k, "max diff", d.max(), "first index", int(np.argmax(d > 0)), "count", int((d > 0).sum()))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
