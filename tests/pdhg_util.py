"""Helpers shared by the PDHG parity tests: run one description through the CUDA backend and
through the oracle with identical options."""
import numpy as np

import prost_b200 as pb
from oracle_binding import OracleProblem, OraclePDHG

TOL = dict(tol_rel_primal=0.0, tol_rel_dual=0.0, tol_abs_primal=0.0, tol_abs_dual=0.0)


def run_cuda(ctx, desc, iters, fuse=True, x0=None, y0=None, tol=None, use_solver=False, solve_dual_problem=False,
             **opts):
    """use_solver=False: exactly `iters` x PerformIteration.  use_solver=True: Solver::Solve with
    max_iters=iters, which stops early on convergence like the reference's loop (solver.cu:141-196)."""
    prob = pb.create_problem(ctx, desc)
    popts = pb.pdhg_options(scale_steps_operator=0, fuse=int(fuse), **opts)
    sopts = pb.solver_options(verbose=0, max_iters=iters, num_cback_calls=0,
                              solve_dual_problem=int(bool(solve_dual_problem)), **(tol or TOL))
    be = pb.BackendPDHG(ctx, prob, popts, sopts)
    if use_solver or solve_dual_problem:
        solver = pb.Solver(prob, be)
        solver.SetOptions(sopts, x0=x0, y0=y0)
        solver.Initialize()
        solver.Solve()
        x, z, y, w = (solver.cur_primal_sol, solver.cur_primal_constr_sol, solver.cur_dual_sol,
                      solver.cur_dual_constr_sol)
        done = solver.iterations
    else:
        prob.Initialize()
        be.Initialize(x0, y0)
        be.PerformIteration(iters)
        x, z, y, w = be.current_solution()
        done = iters
    return dict(x=x, z=z, y=y, w=w, res=be.residuals(), steps=be.stepsizes(), fused=be.is_fused, backend=be,
                problem=prob, iterations=done)


def run_oracle(desc, iters, x0=None, y0=None, tol=None, **opts):
    prob = OracleProblem(desc)
    o = OraclePDHG(prob, **opts, **(tol or TOL))
    o.initialize(x0, y0)
    o.iterate(iters)
    x, z, y, w = o.solution()
    return dict(x=x, z=z, y=y, w=w, res=o.residuals(), steps=o.stepsizes())


def rel_err(a, b):
    """max |a - b| / max(|b|_inf, tiny): the north star's per-element relative agreement."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def assert_parity(got, want, iter_tol=1e-5, res_tol=1e-4, label=""):
    for k in ("x", "y"):
        e = rel_err(got[k], want[k])
        assert e <= iter_tol, f"{label}: iterate {k} rel err {e:.3e}"
    for k in ("z", "w"):
        e = rel_err(got[k], want[k])
        assert e <= 20 * iter_tol, f"{label}: {k} rel err {e:.3e}"
    for k in ("primal_residual", "dual_residual", "primal_var_norm", "dual_var_norm", "eps_primal", "eps_dual"):
        a, b = got["res"][k], want["res"][k]
        assert abs(a - b) <= res_tol * max(abs(b), 1e-6) + 1e-7, f"{label}: {k} {a} vs {b}"
    for a, b in zip(got["steps"], want["steps"]):
        assert abs(a - b) <= 1e-6 * max(abs(b), 1e-12), f"{label}: step sizes {got['steps']} vs {want['steps']}"


def rof_energy(desc, x, lam=10.0):
    """Primal ROF objective (lam/2)|u - f|^2 + sum |grad u|_2 (example_rof_pdgap.m)."""
    nx, ny = desc["blocks"][0][3][0], desc["blocks"][0][3][1]
    f = desc["data"]["f"].astype(np.float64).reshape(nx, ny)
    u = x.astype(np.float64).reshape(nx, ny)
    gx = np.zeros_like(u); gy = np.zeros_like(u)
    gx[:-1, :] = u[1:, :] - u[:-1, :]
    gy[:, :-1] = u[:, 1:] - u[:, :-1]
    return 0.5 * lam * ((u - f) ** 2).sum() + np.sqrt(gx * gx + gy * gy).sum()


# ---- ADMM twins ------------------------------------------------------------------------------

def run_cuda_admm(ctx, desc, iters, tol=None, use_solver=False, **opts):
    """BackendADMM through the C ABI: `iters` x PerformIteration, or Solver::Solve(max_iters=iters)."""
    prob = pb.create_problem(ctx, desc)
    aopts = pb.admm_options(**opts)
    sopts = pb.solver_options(verbose=0, max_iters=iters, num_cback_calls=0, **(tol or TOL))
    be = pb.BackendADMM(ctx, prob, aopts, sopts)
    if use_solver:
        solver = pb.Solver(prob, be)
        solver.SetOptions(sopts)
        solver.Initialize()
        solver.Solve()
        x, z, y, w = (solver.cur_primal_sol, solver.cur_primal_constr_sol, solver.cur_dual_sol,
                      solver.cur_dual_constr_sol)
        done = solver.iterations
    else:
        prob.Initialize()
        be.Initialize()
        be.PerformIteration(iters)
        x, z, y, w = be.current_solution()
        done = iters
    return dict(x=x, z=z, y=y, w=w, res=be.residuals(), steps=be.stepsizes(), backend=be, problem=prob,
                iterations=done)


def run_oracle_admm(desc, iters, tol=None, **opts):
    from oracle_binding import OracleADMM
    prob = OracleProblem(desc)
    o = OracleADMM(prob, **opts, **(tol or TOL))
    o.initialize()
    o.iterate(iters)
    x, z, y, w = o.solution()
    return dict(x=x, z=z, y=y, w=w, res=o.residuals(), steps=o.stepsizes())


def assert_admm_parity(got, want, iter_tol=1e-5, res_tol=1e-4, label="", res_floor=1e-6):
    """Iterates x, z (and the derived duals y, w) per element relative to the largest entry; residual
    norms within res_tol, with an absolute floor where they have converged to rounding noise."""
    for k in ("x", "z"):
        e = rel_err(got[k], want[k])
        assert e <= iter_tol, f"{label}: iterate {k} rel err {e:.3e}"
    for k in ("y", "w"):
        e = rel_err(got[k], want[k])
        assert e <= 20 * iter_tol, f"{label}: {k} rel err {e:.3e}"
    for k in ("primal_var_norm", "dual_var_norm", "eps_primal", "eps_dual"):
        a, b = got["res"][k], want["res"][k]
        assert abs(a - b) <= res_tol * max(abs(b), 1e-6) + 1e-7, f"{label}: {k} {a} vs {b}"
    for k, scale in (("primal_residual", "primal_var_norm"), ("dual_residual", "dual_var_norm")):
        a, b = got["res"][k], want["res"][k]
        floor = res_floor * max(want["res"][scale], 1.0)
        assert abs(a - b) <= res_tol * abs(b) + floor, f"{label}: {k} {a} vs {b}"
    assert abs(got["steps"][0] - want["steps"][0]) <= 1e-6 * abs(want["steps"][0]), f"{label}: rho {got['steps']} vs {want['steps']}"
