"""ADMM parity cases shared by the GPU tests, the live-reference tests and the golden generator.
Each entry: name -> (description factory, outer iterations, BackendADMM options, solver tolerances)."""
import numpy as np

from prost_b200 import synthetic as syn

TOL0 = dict(tol_rel_primal=0.0, tol_rel_dual=0.0, tol_abs_primal=0.0, tol_abs_dual=0.0)
TOL4 = dict(tol_rel_primal=1e-4, tol_rel_dual=1e-4, tol_abs_primal=1e-4, tol_abs_dual=1e-4)


def lasso_alpha(m, n, nnz):
    """LASSO with Pock-Chambolle alpha = 1 scaling: per-element Sigma and T (GemvPrecondK with
    non-trivial diagonal preconditioners, backend_admm.cu:198-272)."""
    d = syn.lasso(m, n, nnz_per_row=nnz)
    d["scaling"] = ("alpha", 1.0)
    return d


def rof_admm(nx, ny):
    """ROF through ADMM: prox_f comes from prox_fstar via Moreau (backend_admm.cu:329-343)."""
    return syn.rof(nx, ny)


# NOTE on residual_iter > 1 with non-zero tolerances: the reference's BackendADMM leaves the residual
# members uninitialised until the first refresh (backend.hpp:82-92, backend_admm.cu:345-351), and
# Solver::Solve compares them every iteration (solver.cu:141-150): on fresh (zero) heap memory the
# reference "converges" after one iteration.  prost_b200 starts from FLT_MAX instead, so such cases
# are only compared with zero tolerances.
def small():
    return {
        "lasso_sparse": (lambda: syn.lasso(600, 200, nnz_per_row=6), 25, dict(), TOL0),
        "lasso_sparse_dense": (lambda: syn.lasso(500, 160, nnz_per_row=8, dense=64), 25, dict(), TOL0),
        "lasso_alpha_scaling": (lambda: lasso_alpha(400, 150, 5), 25, dict(), TOL0),
        "lasso_adaptive_rho": (lambda: syn.lasso(300, 120, nnz_per_row=6), 60, dict(residual_iter=1, rho0=4.0), TOL4),
        "lasso_residual_iter3_cg3": (lambda: syn.lasso(300, 100, nnz_per_row=4), 20,
                                     dict(residual_iter=3, cg_max_iter=3, alpha=1.0), TOL0),
        "rof_admm": (lambda: rof_admm(20, 16), 30, dict(), TOL0),
    }


def medium():
    return {
        "lasso_sparse": (lambda: syn.lasso(20000, 6000, nnz_per_row=12), 30, dict(), TOL0),
        "lasso_sparse_dense": (lambda: syn.lasso(12000, 3000, nnz_per_row=12, dense=256), 30, dict(), TOL0),
        "lasso_alpha_scaling": (lambda: lasso_alpha(8000, 3000, 8), 30, dict(), TOL0),
        "lasso_adaptive_rho": (lambda: syn.lasso(6000, 2000, nnz_per_row=8), 80, dict(residual_iter=1, rho0=4.0), TOL4),
        "lasso_residual_iter3_cg3": (lambda: syn.lasso(5000, 1500, nnz_per_row=6), 30,
                                     dict(residual_iter=3, cg_max_iter=3, alpha=1.0), TOL0),
        "rof_admm": (lambda: rof_admm(64, 48), 40, dict(), TOL0),
    }
