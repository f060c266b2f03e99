"""Generates the golden vectors in this directory by running the UNMODIFIED reference
(oracle/_ref/prost_ref_driver = tum-vision/prost compiled for sm_100, driven through its public
C++ API) on a GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy *.npz here

Every fixture stores the exact inputs together with the reference's outputs, so the CPU suite
(tests/test_oracle_golden.py) can pin the oracle without a GPU and without /root/reference.
Inputs are seeded; problem descriptions come from tests/cases.py (small=True) and
prost_b200/synthetic.py."""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import admm_cases     # noqa: E402
import cases          # noqa: E402
import ref_driver     # noqa: E402
from prost_b200 import synthetic as syn   # noqa: E402

PDHG_SMALL = {
    "rof_alg1": (lambda: syn.rof(24, 19), 60, dict(stepsize="alg1", residual_iter=3)),
    "rof_alg2": (lambda: syn.rof(20, 16), 40, dict(stepsize="alg2", residual_iter=3, alg2_gamma=0.5)),
    "rof_goldstein": (lambda: syn.rof(24, 19), 80, dict(stepsize="goldstein", residual_iter=3)),
    "rof_boyd": (lambda: syn.rof(20, 16), 200, dict(stepsize="boyd", residual_iter=3)),
    "tvl1_color": (lambda: syn.tvl1(16, 12, nc=3), 120, dict(stepsize="boyd", residual_iter=10)),
    "tv3d": (lambda: syn.tv3d(8, 8, 6), 120, dict(stepsize="boyd", residual_iter=10)),
    "lifting_L8": (lambda: syn.lifting(8, 6, 8), 120, dict(stepsize="boyd", residual_iter=10)),
    "rof_warm": (lambda: syn.rof(12, 10), 3, dict(stepsize="alg1", residual_iter=1)),
}
TOL4 = dict(tol_rel_primal=1e-4, tol_rel_dual=1e-4, tol_abs_primal=1e-4, tol_abs_dual=1e-4)


def main(out, only=""):
    os.makedirs(out, exist_ok=True)
    assert ref_driver.available(), "oracle/_ref/prost_ref_driver missing"
    if only == "admm":
        return main_admm(out)
    if only == "new_prox":        # only the prox cases that have no fixture yet (small files, quick)
        for name, (desc, n) in cases.all_prox_cases(small=True).items():
            if os.path.exists(os.path.join(HERE, f"prox_{name}.npz")):
                continue
            r = np.random.default_rng(zlib.crc32(name.encode()))
            arg = (2 * r.standard_normal(n)).astype(np.float32)
            td = r.uniform(0.5, 1.5, n).astype(np.float32)
            res = ref_driver.run_prox(desc, arg, td, 0.7)
            np.savez_compressed(os.path.join(out, f"prox_{name}.npz"), arg=arg, tau_diag=td, tau=np.float32(0.7),
                                res=res)
        return
    for name, blocks in cases.linop_cases(small=True).items():
        r = np.random.default_rng(zlib.crc32(name.encode()))
        from oracle_binding import OracleProblem
        m, n = OracleProblem(blocks=blocks).linop_size()
        if m == 0 or n == 0:
            continue
        x, y = r.random(n).astype(np.float32), r.random(m).astype(np.float32)
        f = ref_driver.run_linop(blocks, x, False)
        a = ref_driver.run_linop(blocks, y, True)
        np.savez_compressed(os.path.join(out, f"linop_{name}.npz"), x=x, y=y, fwd=f["res"], adj=a["res"],
                            rowsum=f["rowsum"], colsum=f["colsum"])
    for name, (desc, n) in cases.all_prox_cases(small=True).items():
        r = np.random.default_rng(zlib.crc32(name.encode()))
        arg = (2 * r.standard_normal(n)).astype(np.float32)
        td = r.uniform(0.5, 1.5, n).astype(np.float32)
        res = ref_driver.run_prox(desc, arg, td, 0.7)
        np.savez_compressed(os.path.join(out, f"prox_{name}.npz"), arg=arg, tau_diag=td, tau=np.float32(0.7), res=res)
    for name, (fn, iters, opts) in PDHG_SMALL.items():
        desc = fn()
        x0 = y0 = None
        if name == "rof_warm":
            r = np.random.default_rng(0)
            x0 = r.random(desc["ncols"]).astype(np.float32)
            y0 = (0.3 * r.standard_normal(desc["nrows"])).astype(np.float32)
        w = ref_driver.run_solve(desc, iters, x0=x0, y0=y0, tol=TOL4, **opts)
        keys = sorted(w["res"])
        extra = dict(x0=x0, y0=y0) if x0 is not None else {}
        np.savez_compressed(os.path.join(out, f"pdhg_{name}.npz"), x=w["x"], y=w["y"], z=w["z"], w=w["w"],
                            res=np.array([w["res"][k] for k in keys]), res_keys=np.array(keys),
                            iterations=np.int64(w["info"]["iterations"]), **extra)
    main_admm(out)
    print("wrote", len(os.listdir(out)), "fixtures to", out)


def main_admm(out):
    for name, (fn, iters, opts, tol) in admm_cases.small().items():
        desc = fn()
        w = ref_driver.run_solve(desc, iters, tol=tol, admm=opts)
        keys = sorted(w["res"])
        np.savez_compressed(os.path.join(out, f"admm_{name}.npz"), x=w["x"], y=w["y"], z=w["z"], w=w["w"],
                            res=np.array([w["res"][k] for k in keys]), res_keys=np.array(keys),
                            iterations=np.int64(w["info"]["iterations"]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE, sys.argv[2] if len(sys.argv) > 2 else "")
