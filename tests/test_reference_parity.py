"""GPU suite: parity with the REFERENCE'S OWN CUDA implementation (tum-vision/prost compiled
unmodified for sm_100 into oracle/_ref, driven through its public C++ API by
oracle/driver/prost_driver.cu) on identical synthetic inputs and iteration counts.

North-star bars: per-element iterates within 1e-5 relative, final residuals / objective within
1e-4 relative, index work bit-exact.  The same runs also pin the CPU oracle against the live
reference (the committed fixtures under tests/golden/ come from tests/golden/make_golden.py)."""
import os
import zlib

import numpy as np
import pytest

import admm_cases
import cases
import prost_b200 as pb
import ref_driver
from oracle_binding import OracleProblem, oracle_prox_eval
from pdhg_util import (assert_admm_parity, rel_err, rof_energy, run_cuda, run_cuda_admm, run_oracle,
                       run_oracle_admm)
from prost_b200 import synthetic as syn

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref/prost_ref_driver not built")]

LINOPS = cases.linop_cases(small=False)
PROXES = cases.prox_cases(small=False)

def close_frac(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.abs(a - b) > tol * np.maximum(1.0, np.abs(b))).mean()) if a.size else 0.0


@pytest.mark.parametrize("name", sorted(LINOPS))
def test_linop_vs_reference(ctx, name):
    blocks = LINOPS[name]
    op = pb.create_linop(ctx, blocks)
    r = np.random.default_rng(3)
    x, y = r.random(op.ncols).astype(np.float32), r.random(op.nrows).astype(np.float32)
    ref_f = ref_driver.run_linop(blocks, x, False)
    ref_a = ref_driver.run_linop(blocks, y, True)
    lib = not all(b[0] in ("gradient2d", "gradient3d", "diags", "zero") for b in blocks)
    tol = 1e-4 if lib else 1e-6            # cuSPARSE / cuBLAS sum in a different order
    assert close_frac(op.Eval(x), ref_f["res"], tol) == 0
    if name != "diags_wide":               # reference adjoint launch is sized by nrows (block_diags.cu:211)
        assert close_frac(op.EvalAdjoint(y), ref_a["res"], tol) == 0
    assert np.array_equal(op.row_sums(1.0), ref_f["rowsum"])
    assert np.array_equal(op.col_sums(1.0), ref_f["colsum"])
    # pin the CPU oracle against the same reference outputs
    orc = OracleProblem(blocks=blocks)      # (contraction-free CPU arithmetic: looser on long diagonal sums)
    assert close_frac(orc.linop(x, False), ref_f["res"], max(tol, 1e-5)) == 0


@pytest.mark.parametrize("name", sorted(PROXES))
def test_prox_vs_reference(ctx, name):
    desc, n = PROXES[name]
    r = np.random.default_rng(zlib.crc32(name.encode()))
    arg = (2 * r.standard_normal(n)).astype(np.float32)
    tau_diag = r.uniform(0.5, 1.5, n).astype(np.float32)
    tau = 0.7
    want = ref_driver.run_prox(desc, arg, tau_diag, tau)
    got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau)
    lo, hi = desc[1], desc[1] + desc[2]
    jumpy = any(k in name for k in ("l0", "truncquad", "trunclin", "lq"))
    frac = close_frac(got[lo:hi], want[lo:hi], 1e-5)
    assert frac <= (2e-3 if jumpy else 0.0), (name, frac)
    if "permute" in name and "moreau" not in name and "norm2" not in name:
        assert np.array_equal(got[lo:hi], want[lo:hi])
    orc = oracle_prox_eval(desc, arg, tau_diag, tau)
    frac_o = close_frac(orc[lo:hi], want[lo:hi], 2e-5)
    assert frac_o <= (2e-3 if jumpy else 0.0), ("oracle", name, frac_o)


def assert_ref_parity(got, want, label, iter_tol=1e-5, res_tol=1e-4):
    for k in ("x", "y"):
        e = rel_err(got[k], want[k])
        assert e <= iter_tol, f"{label}: iterate {k} rel err {e:.3e}"
    for k in ("z", "w"):
        e = rel_err(got[k], want[k])
        assert e <= 20 * iter_tol, f"{label}: {k} rel err {e:.3e}"
    for k in ("primal_residual", "dual_residual", "primal_var_norm", "dual_var_norm", "eps_primal", "eps_dual"):
        a, b = got["res"][k], want["res"][k]
        assert abs(a - b) <= res_tol * max(abs(b), 1e-6) + 1e-7, f"{label}: {k} {a} vs {b}"


PDHG_CASES = {
    "rof_alg1": (lambda: syn.rof(48, 37), 150, dict(stepsize="alg1", residual_iter=3)),
    "rof_alg2": (lambda: syn.rof(40, 36), 150, dict(stepsize="alg2", residual_iter=3, alg2_gamma=0.5)),
    "rof_goldstein": (lambda: syn.rof(48, 37), 150, dict(stepsize="goldstein", residual_iter=3)),
    "rof_boyd": (lambda: syn.rof(40, 36), 150, dict(stepsize="boyd", residual_iter=3)),
    "tvl1_color": (lambda: syn.tvl1(64, 48, nc=3), 200, dict(stepsize="boyd", residual_iter=10)),
    "tv3d": (lambda: syn.tv3d(20, 24, 16), 200, dict(stepsize="boyd", residual_iter=10)),
    "lifting_L8": (lambda: syn.lifting(24, 20, 8), 200, dict(stepsize="boyd", residual_iter=10)),
    "lifting_L32": (lambda: syn.lifting(16, 12, 32), 100, dict(stepsize="boyd", residual_iter=10)),
    # column height a multiple of 64: the staged passes copy their operands cooperatively (pb_stencil_staged.cuh);
    # 128 rows = two CTAs per column (the +-1 row neighbours cross the CTA boundary), 5 / 6 columns incl. both edges
    "lifting_coop_L32": (lambda: syn.lifting(5, 128, 32), 60, dict(stepsize="boyd", residual_iter=10)),
    "lifting_coop_L8": (lambda: syn.lifting(6, 64, 8), 120, dict(stepsize="alg1", residual_iter=7)),
}
TOL4 = dict(tol_rel_primal=1e-4, tol_rel_dual=1e-4, tol_abs_primal=1e-4, tol_abs_dual=1e-4)


@pytest.mark.parametrize("name", sorted(PDHG_CASES))
@pytest.mark.parametrize("fuse", [1, 2, 0])
def test_pdhg_vs_reference(ctx, name, fuse):
    desc_fn, iters, opts = PDHG_CASES[name]
    desc = desc_fn()
    want = ref_driver.run_solve(desc, iters, tol=TOL4, **opts)
    # Solver::Solve on both sides: the loop stops at the first iteration whose residuals are below
    # the tolerances, and must do so at the same iteration
    got = run_cuda(ctx, desc, iters, fuse=fuse, tol=TOL4, use_solver=True, **opts)
    assert got["iterations"] == int(want["info"]["iterations"]), (got["iterations"], want["info"]["iterations"])
    assert_ref_parity(got, want, f"{name} fuse={fuse}")
    if fuse == 1:
        # the same reference run pins the CPU oracle (looser under Alg2, see test_gpu_pdhg.py)
        orc = run_oracle(desc, got["iterations"], tol=TOL4, **opts)
        loose = opts["stepsize"] == "alg2"
        assert_ref_parity(orc, want, f"oracle {name}", iter_tol=5e-5 if loose else 1e-5,
                          res_tol=5e-3 if loose else 1e-4)


@pytest.mark.parametrize("name", ["rof_alg1", "rof_boyd", "tvl1_color", "lifting_L8"])
def test_pdhg_dual_problem_vs_reference(ctx, name):
    """Solver::Options::solve_dual_problem (solver.cu:80-84, 199-203; dual_linearoperator.cu:38-80): PDHG on the
    dualised problem (prox_g <-> prox_fstar, K -> -K^T, x0 <-> y0, swapped scalings), solutions swapped back by
    cur_primal_sol() & co.  Same iteration count and iterates as the live reference; warm start included."""
    desc_fn, iters, opts = PDHG_CASES[name]
    desc = desc_fn()
    r = np.random.default_rng(5)
    x0 = r.random(desc["ncols"]).astype(np.float32)
    y0 = (0.1 * r.standard_normal(desc["nrows"])).astype(np.float32)
    want = ref_driver.run_solve(desc, iters, tol=TOL4, x0=x0, y0=y0, solve_dual_problem=True, **opts)
    for fuse in (1, 0):
        got = run_cuda(ctx, desc, iters, fuse=fuse, tol=TOL4, x0=x0, y0=y0, solve_dual_problem=True, **opts)
        assert got["iterations"] == int(want["info"]["iterations"]), (got["iterations"], want["info"]["iterations"])
        assert_ref_parity(got, want, f"dual problem {name} fuse={fuse}")
    assert not got["fused"]          # a dualised problem takes the unfused schedule (K -> -K^T at run time)


def test_c1_rof_512_1000_iterations_vs_reference(ctx):
    """BASELINE config 1 against the reference's CUDA solver: ROF 512x512, 1000 PDHG iterations."""
    desc = syn.rof(512, 512)
    opts = dict(stepsize="alg1", residual_iter=10)
    want = ref_driver.run_solve(desc, 1000, **opts)
    got = run_cuda(ctx, desc, 1000, **opts)
    assert got["fused"]
    assert_ref_parity(got, want, "C1")
    e_got, e_want = rof_energy(desc, got["x"]), rof_energy(desc, want["x"])
    assert abs(e_got - e_want) <= 1e-4 * abs(e_want)


def test_warm_start_vs_reference(ctx):
    desc = syn.rof(32, 24)
    r = np.random.default_rng(0)
    x0 = r.random(desc["ncols"]).astype(np.float32)
    y0 = (0.3 * r.standard_normal(desc["nrows"])).astype(np.float32)
    for iters in (1, 2, 25):
        want = ref_driver.run_solve(desc, iters, x0=x0, y0=y0, stepsize="alg1", residual_iter=1)
        for fuse in (1, 2, 0):
            got = run_cuda(ctx, desc, iters, fuse=fuse, x0=x0, y0=y0, stepsize="alg1", residual_iter=1)
            assert_ref_parity(got, want, f"warm start iters={iters} fuse={fuse}")


ADMM_CASES = admm_cases.medium()


@pytest.mark.parametrize("name", sorted(ADMM_CASES))
def test_admm_vs_reference(ctx, name):
    """BackendADMM against the reference's BackendADMM<float> + cgls::Solve (cuSPARSE SpMV, cuBLAS
    gemv / axpy inside) through Solver::Solve on both sides; the same run pins the CPU oracle."""
    fn, iters, opts, tol = ADMM_CASES[name]
    desc = fn()
    want = ref_driver.run_solve(desc, iters, tol=tol, admm=opts)
    want["steps"] = None
    got = run_cuda_admm(ctx, desc, iters, tol=tol, use_solver=True, **opts)
    assert got["iterations"] == int(want["info"]["iterations"]), (got["iterations"], want["info"]["iterations"])
    orc = run_oracle_admm(desc, got["iterations"], tol=tol, **opts)
    for label, res in (("cuda", got), ("oracle", orc)):
        res = dict(res, steps=(1.0,))
        assert_admm_parity(res, dict(want, steps=(1.0,)), label=f"{label} {name}")


def test_admm_cxx_api_drop_in():
    """The driver written against prost's public C++ API with BackendADMM, linked against
    include/prost/*.hpp + libprost_b200.so, vs the same program linked against the reference."""
    if not ref_driver.available(ref_driver.OUR_DRIVER):
        pytest.skip("prost_b200_driver not built")
    fn, iters, opts, tol = ADMM_CASES["lasso_sparse_dense"]
    desc = fn()
    want = ref_driver.run_solve(desc, iters, tol=tol, admm=opts)
    got = ref_driver.run_solve(desc, iters, tol=tol, admm=opts, binary=ref_driver.OUR_DRIVER)
    assert int(got["info"]["iterations"]) == int(want["info"]["iterations"])
    assert_admm_parity(dict(got, steps=(1.0,)), dict(want, steps=(1.0,)), label="c++ drop-in admm")


@pytest.mark.skipif(not ref_driver.available(ref_driver.OUR_DRIVER), reason="prost_b200_driver not built")
@pytest.mark.parametrize("name", ["rof_boyd", "tvl1_color", "lifting_L8"])
def test_cxx_api_drop_in(name):
    """The SAME C++ program (oracle/driver/prost_driver.cu, prost's public API only) linked once
    against the reference and once against include/prost/*.hpp + libprost_b200.so."""
    desc_fn, iters, opts = PDHG_CASES[name]
    desc = desc_fn()
    want = ref_driver.run_solve(desc, iters, tol=TOL4, **opts)
    got = ref_driver.run_solve(desc, iters, tol=TOL4, binary=ref_driver.OUR_DRIVER, **opts)
    assert int(got["info"]["iterations"]) == int(want["info"]["iterations"])
    assert_ref_parity(got, want, f"c++ drop-in {name}")
    blocks = cases.linop_cases(small=True)["lifting_K"]
    r = np.random.default_rng(1)
    from oracle_binding import OracleProblem as _OP
    m, n = _OP(blocks=blocks).linop_size()
    x = r.random(n).astype(np.float32)
    a = ref_driver.run_linop(blocks, x, False)
    b = ref_driver.run_linop(blocks, x, False, binary=ref_driver.OUR_DRIVER)
    assert np.array_equal(a["res"], b["res"]) and np.array_equal(a["rowsum"], b["rowsum"])
    pdesc, pn = cases.prox_cases(small=True)["simplex_d5_il0"]
    arg = (2 * r.standard_normal(pn)).astype(np.float32)
    pa = ref_driver.run_prox(pdesc, arg, np.ones(pn), 1.0)
    pbv = ref_driver.run_prox(pdesc, arg, np.ones(pn), 1.0, binary=ref_driver.OUR_DRIVER)
    assert close_frac(pbv, pa, 1e-6) == 0
