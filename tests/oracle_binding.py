"""ctypes binding of the CPU oracle (oracle/prost_oracle.cpp) -- TEST INFRASTRUCTURE.

Builds oracle problems from the same descriptions as prost_b200.factory so a test can run one
description through the CUDA library and through the oracle and compare.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libprost_oracle.so")

FUNCTIONS_1D = ["zero", "abs", "square", "ind_leq0", "ind_geq0", "ind_eq0", "ind_box01", "max_pos0",
                "l0", "huber", "lq", "lq_plus_eps", "truncquad", "trunclin"]


def _load():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return C.CDLL(LIB)


lib = _load()
fp = C.POINTER(C.c_float)
sz = C.c_size_t
lib.orc_problem_new.restype = C.c_void_p
lib.orc_pdhg_new.restype = C.c_void_p
lib.orc_last_error.restype = C.c_char_p
lib.orc_nrows.restype = sz
lib.orc_ncols.restype = sz
lib.orc_problem_free.argtypes = [C.c_void_p]
lib.orc_add_gradient.argtypes = [C.c_void_p, C.c_int, sz, sz, sz, sz, sz, C.c_int]
lib.orc_add_diags.argtypes = [C.c_void_p, sz, sz, sz, sz, sz, C.POINTER(C.c_int64), fp]
lib.orc_add_sparse_csc.argtypes = [C.c_void_p, sz, sz, C.c_int, C.c_int, C.c_int, fp,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
lib.orc_add_dense.argtypes = [C.c_void_p, sz, sz, sz, sz, fp]
lib.orc_add_kron.argtypes = [C.c_void_p, sz, sz, sz, sz, sz, C.c_int, fp]
lib.orc_add_zero.argtypes = [C.c_void_p, sz, sz, sz, sz]
lib.orc_prox_elem.argtypes = [C.c_void_p, C.c_int, sz, sz, sz, C.c_int, C.c_int, C.c_int, C.POINTER(fp),
                              C.POINTER(sz)]
lib.orc_prox_simplex.argtypes = [C.c_void_p, sz, sz, sz, C.c_int, C.c_int]
lib.orc_prox_ind_sum.argtypes = [C.c_void_p, sz, sz, sz, C.c_int, C.c_int]
lib.orc_prox_ind_sum_indexed.argtypes = [C.c_void_p, sz, sz, sz, sz, C.POINTER(C.c_ulonglong), C.c_float, sz, sz,
                                         C.POINTER(C.c_ulonglong), C.c_float]
lib.orc_prox_spectral.argtypes = [C.c_void_p, C.c_int, sz, sz, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(fp),
                                  C.POINTER(sz)]
lib.orc_prox_ind_range.argtypes = [C.c_void_p, sz, sz, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), fp]
lib.orc_prox_ind_epi_conjquad_1d.argtypes = [C.c_void_p, sz, sz, C.c_int, C.c_int, C.POINTER(fp), C.POINTER(sz)]
lib.orc_prox_ind_halfspace.argtypes = [C.c_void_p, sz, sz, sz, C.c_int, C.c_int, fp, sz, fp, sz]
lib.orc_prox_ind_soc.argtypes = [C.c_void_p, sz, sz, sz, C.c_int, C.c_int]
lib.orc_prox_epi_quad.argtypes = [C.c_void_p, sz, sz, sz, C.c_int, C.c_int, fp, sz, fp, sz, fp, sz]
lib.orc_prox_moreau.argtypes = [C.c_void_p, C.c_int]
lib.orc_prox_permute.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), sz]
lib.orc_prox_transform.argtypes = [C.c_void_p, C.c_int, C.POINTER(fp), C.POINTER(sz)]
lib.orc_prox_zero.argtypes = [C.c_void_p, sz, sz]
lib.orc_set_prox.argtypes = [C.c_void_p, C.c_int, C.c_int]
lib.orc_set_dims.argtypes = [C.c_void_p, sz, sz]
lib.orc_set_scaling_alpha.argtypes = [C.c_void_p, C.c_float]
lib.orc_set_scaling_identity.argtypes = [C.c_void_p]
lib.orc_set_scaling_custom.argtypes = [C.c_void_p, fp, sz, fp, sz]
lib.orc_initialize.argtypes = [C.c_void_p]
lib.orc_nrows.argtypes = [C.c_void_p]
lib.orc_ncols.argtypes = [C.c_void_p]
lib.orc_linop_size.argtypes = [C.c_void_p, C.POINTER(sz), C.POINTER(sz)]
lib.orc_linop_eval.argtypes = [C.c_void_p, fp, fp, C.c_int]
lib.orc_row_sums.argtypes = [C.c_void_p, C.c_float, fp, sz]
lib.orc_col_sums.argtypes = [C.c_void_p, C.c_float, fp, sz]
lib.orc_get_scaling.argtypes = [C.c_void_p, fp, fp]
lib.orc_prox_eval.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_float, C.c_int]
lib.orc_pdhg_new.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int] + [C.c_float] * 6 + [C.c_int] + \
    [C.c_float] * 4
lib.orc_pdhg_free.argtypes = [C.c_void_p]
lib.orc_pdhg_init.argtypes = [C.c_void_p, fp, sz, fp, sz]
lib.orc_pdhg_iterate.argtypes = [C.c_void_p, C.c_int]
lib.orc_pdhg_residuals.argtypes = [C.c_void_p, fp]
lib.orc_pdhg_stepsizes.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
lib.orc_pdhg_solution.argtypes = [C.c_void_p, fp, fp, fp, fp]
lib.orc_set_num_threads.argtypes = [C.c_int]
lib.orc_admm_new.restype = C.c_void_p
lib.orc_admm_new.argtypes = [C.c_void_p] + [C.c_double] * 5 + [C.c_int, C.c_int] + [C.c_float] * 3 + [C.c_float] * 4
lib.orc_admm_free.argtypes = [C.c_void_p]
lib.orc_admm_init.argtypes = [C.c_void_p]
lib.orc_admm_iterate.argtypes = [C.c_void_p, C.c_int]
lib.orc_admm_residuals.argtypes = [C.c_void_p, fp]
lib.orc_admm_stepsizes.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
lib.orc_admm_solution.argtypes = [C.c_void_p, fp, fp, fp, fp]


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).ravel())


def _p(a):
    return a.ctypes.data_as(fp)


class OracleProblem:
    """Oracle-side twin of prost_b200.factory.create_problem / create_linop / create_prox."""

    def __init__(self, desc=None, blocks=None):
        self.h = C.c_void_p(lib.orc_problem_new())
        self._keep = []
        for b in (blocks or []):
            self.add_block(b)
        if desc is not None:
            for b in desc.get("blocks", []):
                self.add_block(b)
            for which, key in enumerate(("prox_g", "prox_f", "prox_gstar", "prox_fstar")):
                for d in desc.get(key, []):
                    lib.orc_set_prox(self.h, which, self.add_prox(d))
            if "nrows" in desc:
                lib.orc_set_dims(self.h, desc["nrows"], desc["ncols"])
            sc = desc.get("scaling", ("alpha", 1.0))
            if sc[0] == "alpha":
                lib.orc_set_scaling_alpha(self.h, sc[1])
            elif sc[0] == "identity":
                lib.orc_set_scaling_identity(self.h)
            else:
                l, r = _f32(sc[1]), _f32(sc[2])
                lib.orc_set_scaling_custom(self.h, _p(l), l.size, _p(r), r.size)
            if lib.orc_initialize(self.h) != 0:
                raise RuntimeError(lib.orc_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_problem_free(self.h)
            self.h = None

    def add_block(self, desc):
        name, row, col, data = desc
        if name in ("gradient2d", "gradient3d"):
            nx, ny, L, lf = data
            lib.orc_add_gradient(self.h, int(name == "gradient3d"), row, col, nx, ny, L, int(lf))
        elif name == "diags":
            nrows, ncols, factors, offsets = data
            o = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
            f = _f32(factors)
            lib.orc_add_diags(self.h, row, col, nrows, ncols, o.size, o.ctypes.data_as(C.POINTER(C.c_int64)), _p(f))
        elif name == "sparse":
            import scipy.sparse as sp
            A = sp.csc_matrix(data[0])
            A.sort_indices()
            v = _f32(A.data)
            ptr = np.ascontiguousarray(A.indptr.astype(np.int32))
            ind = np.ascontiguousarray(A.indices.astype(np.int32))
            lib.orc_add_sparse_csc(self.h, row, col, A.shape[0], A.shape[1], A.nnz, _p(v),
                                   ptr.ctypes.data_as(C.POINTER(C.c_int32)),
                                   ind.ctypes.data_as(C.POINTER(C.c_int32)))
        elif name == "dense":
            A = np.asarray(data[0], dtype=np.float32)
            d = np.ascontiguousarray(A.T).ravel()
            lib.orc_add_dense(self.h, row, col, A.shape[0], A.shape[1], _p(d))
        elif name in ("dense_kron_id", "id_kron_dense", "sparse_kron_id", "id_kron_sparse"):
            # oracle only so far: data = [K, diaglength]
            K = data[0].toarray() if hasattr(data[0], "toarray") else np.asarray(data[0])
            K = np.asarray(K, dtype=np.float32)
            d = np.ascontiguousarray(K.T).ravel()
            lib.orc_add_kron(self.h, row, col, K.shape[0], K.shape[1], int(data[1]), int(name.startswith("id_")), _p(d))
        elif name == "zero":
            lib.orc_add_zero(self.h, row, col, data[0], data[1])
        else:
            raise ValueError(name)

    def add_prox(self, desc):
        name, idx, size, diagsteps, data = desc
        if name.startswith("elem_operation:1d:") or name.startswith("elem_operation:norm2:"):
            count, dim, il, coeffs = data
            arrs = [_f32(c) for c in coeffs]
            ptrs = (fp * 7)(*[_p(a) for a in arrs])
            lens = (sz * 7)(*[a.size for a in arrs])
            fn = FUNCTIONS_1D.index(name.split(":")[2])
            return lib.orc_prox_elem(self.h, int(":norm2:" in name), idx, count, dim, int(il), int(diagsteps),
                                     fn, ptrs, lens)
        if name == "elem_operation:ind_simplex":
            count, dim, il = data[:3]
            return lib.orc_prox_simplex(self.h, idx, count, dim, int(il), int(diagsteps))
        if name == "ind_halfspace":            # oracle only so far
            count, dim, il, (a, b) = data
            a, b = _f32(a), _f32(b)
            return lib.orc_prox_ind_halfspace(self.h, idx, count, dim, int(il), int(diagsteps), _p(a), a.size,
                                              _p(b), b.size)
        if name == "ind_sum":                  # ProxIndSum: [dim, inds, sum(, dim2, inds2, sum2)]
            u64p = C.POINTER(C.c_ulonglong)
            a = np.ascontiguousarray(np.asarray(data[1], dtype=np.uint64).ravel())
            two = len(data) == 6
            b = np.ascontiguousarray(np.asarray(data[4] if two else [], dtype=np.uint64).ravel())
            d1, d2 = int(data[0]), int(data[3]) if two else 0
            return lib.orc_prox_ind_sum_indexed(self.h, idx, size, a.size // d1, d1, a.ctypes.data_as(u64p),
                                                float(data[2]), (b.size // d2) if two else 0, d2,
                                                b.ctypes.data_as(u64p) if two else None,
                                                float(data[5]) if two else 0.0)
        for kind_id, kind in enumerate(("singular_nx2", "eigen_2x2", "eigen_3x3", "eigen_nxn")):
            prefix = f"elem_operation:{kind}:"
            if name.startswith(prefix):
                count, dim, il, coeffs = data
                fun = name[len(prefix):]
                fn2d = {"ind_l1_ball": 1, "moreau:ind_l1_ball": 2}.get(fun, 0)
                fn1d = 0 if fn2d else FUNCTIONS_1D.index(fun[len("sum_1d:"):] if fun.startswith("sum_1d:") else fun)
                arrs = [_f32(c) for c in coeffs]
                ptrs = (fp * 7)(*[_p(a) for a in arrs])
                lens = (sz * 7)(*[a.size for a in arrs])
                return lib.orc_prox_spectral(self.h, kind_id, idx, count, dim, int(il), int(diagsteps), fn1d, fn2d,
                                             ptrs, lens)
        if name in ("elem_operation:mass4", "elem_operation:ind_comass4_ball", "elem_operation:mass5",
                    "elem_operation:ind_comass5_ball"):
            count, dim, il = data[:3]
            kind_id = {"mass4": 4, "ind_comass4_ball": 5, "mass5": 6, "ind_comass5_ball": 7}[name.split(":")[1]]
            cost = data[3][0] if len(data) > 3 else [1.0]
            arrs = [_f32(np.atleast_1d(c)) for c in (cost, [0.0], [1.0], [0.0], [0.0], [0.0], [0.0])]
            ptrs = (fp * 7)(*[_p(a) for a in arrs])
            lens = (sz * 7)(*[a.size for a in arrs])
            return lib.orc_prox_spectral(self.h, kind_id, idx, count, dim, int(il), int(diagsteps), 0, 0, ptrs, lens)
        if name == "ind_range":
            import scipy.sparse as sp
            A = sp.csc_matrix(data[0])
            A.sort_indices()
            AA = np.asarray(data[1] if len(data) > 1 and data[1] is not None else (A.T @ A).toarray(), np.float32)
            val = _f32(A.data)
            ptr = np.ascontiguousarray(A.indptr.astype(np.int32))
            ix = np.ascontiguousarray(A.indices.astype(np.int32))
            aa = np.ascontiguousarray(AA.T).ravel()
            ip = C.POINTER(C.c_int)
            return lib.orc_prox_ind_range(self.h, idx, size, int(diagsteps), A.shape[0], A.shape[1], A.nnz, _p(val),
                                          ptr.ctypes.data_as(ip), ix.ctypes.data_as(ip), _p(aa))
        if name == "ind_epi_conjquad_1d":
            count, il, coeffs = data
            arrs = [_f32(np.atleast_1d(c)) for c in coeffs]
            ptrs = (fp * 5)(*[_p(a) for a in arrs])
            lens = (sz * 5)(*[a.size for a in arrs])
            return lib.orc_prox_ind_epi_conjquad_1d(self.h, idx, count, int(il), int(diagsteps), ptrs, lens)
        if name == "ind_soc":
            count, dim, il = data[:3]
            return lib.orc_prox_ind_soc(self.h, idx, count, dim, int(il), int(diagsteps))
        if name == "elem_operation:ind_sum":
            count, dim, il = data[:3]
            return lib.orc_prox_ind_sum(self.h, idx, count, dim, int(il), int(diagsteps))
        if name == "ind_epi_quad":
            count, dim, il, (a, b, c) = data
            a, b, c = _f32(a), _f32(b), _f32(c)
            return lib.orc_prox_epi_quad(self.h, idx, count, dim, int(il), int(diagsteps), _p(a), a.size,
                                         _p(b), b.size, _p(c), c.size)
        if name == "moreau":
            return lib.orc_prox_moreau(self.h, self.add_prox(data[0]))
        if name == "permute":
            perm = np.ascontiguousarray(np.asarray(data[1], dtype=np.int32))
            return lib.orc_prox_permute(self.h, self.add_prox(data[0]), perm.ctypes.data_as(C.POINTER(C.c_int)),
                                        perm.size)
        if name == "transform":
            arrs = [_f32(np.atleast_1d(c)) for c in data[:5]]
            ptrs = (fp * 5)(*[_p(a) for a in arrs])
            lens = (sz * 5)(*[a.size for a in arrs])
            return lib.orc_prox_transform(self.h, self.add_prox(data[5]), ptrs, lens)
        if name == "zero":
            return lib.orc_prox_zero(self.h, idx, size)
        raise ValueError(name)

    # -- operator ------------------------------------------------------------------------------
    def linop_size(self):
        r, c = sz(), sz()
        lib.orc_linop_size(self.h, C.byref(r), C.byref(c))
        return r.value, c.value

    def linop(self, rhs, transpose=False):
        m, n = self.linop_size()
        if lib.orc_nrows(self.h) == 0:
            lib.orc_set_dims(self.h, m, n)
        m, n = lib.orc_nrows(self.h), lib.orc_ncols(self.h)
        rhs = _f32(rhs)
        out = np.zeros(n if transpose else m, dtype=np.float32)
        lib.orc_linop_eval(self.h, _p(out), _p(rhs), int(transpose))
        return out

    def row_sums(self, alpha):
        m, _ = self.linop_size()
        out = np.empty(m, np.float32)
        lib.orc_row_sums(self.h, alpha, _p(out), m)
        return out

    def col_sums(self, alpha):
        _, n = self.linop_size()
        out = np.empty(n, np.float32)
        lib.orc_col_sums(self.h, alpha, _p(out), n)
        return out

    def scaling(self):
        l = np.empty(lib.orc_nrows(self.h), np.float32)
        r = np.empty(lib.orc_ncols(self.h), np.float32)
        lib.orc_get_scaling(self.h, _p(l), _p(r))
        return l, r

    # -- prox ----------------------------------------------------------------------------------
    def prox_eval(self, prox_id, arg, tau_diag, tau, invert=False):
        arg, td = _f32(arg), _f32(tau_diag)
        res = np.zeros_like(arg)
        lib.orc_prox_eval(self.h, prox_id, _p(res), _p(arg), _p(td), tau, int(invert))
        return res

    nrows = property(lambda s: lib.orc_nrows(s.h))
    ncols = property(lambda s: lib.orc_ncols(s.h))


def oracle_prox_eval(desc, arg, tau_diag, tau, invert=False):
    p = OracleProblem()
    pid = p.add_prox(desc)
    return p.prox_eval(pid, arg, tau_diag, tau, invert)


_STEPS = {"alg1": 1, "alg2": 2, "goldstein": 3, "boyd": 4}


class OraclePDHG:
    """Oracle twin of BackendPDHG with the same option names."""

    def __init__(self, prob, tau0=1.0, sigma0=1.0, residual_iter=1, alg2_gamma=0.0, arg_alpha0=0.5,
                 arg_nu=0.95, arg_delta=1.5, arb_delta=1.05, arb_tau=0.8, stepsize="boyd",
                 tol_rel_primal=1e-4, tol_rel_dual=1e-4, tol_abs_primal=1e-4, tol_abs_dual=1e-4):
        self.prob = prob
        self.h = C.c_void_p(lib.orc_pdhg_new(prob.h, tau0, sigma0, residual_iter, alg2_gamma, arg_alpha0,
                                             arg_nu, arg_delta, arb_delta, arb_tau, _STEPS[stepsize],
                                             tol_rel_primal, tol_rel_dual, tol_abs_primal, tol_abs_dual))

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_pdhg_free(self.h)
            self.h = None

    def initialize(self, x0=None, y0=None):
        x0a = _f32(x0) if x0 is not None else None
        y0a = _f32(y0) if y0 is not None else None
        rc = lib.orc_pdhg_init(self.h, _p(x0a) if x0a is not None else None, x0a.size if x0a is not None else 0,
                               _p(y0a) if y0a is not None else None, y0a.size if y0a is not None else 0)
        if rc != 0:
            raise RuntimeError(lib.orc_last_error().decode())

    def iterate(self, n=1):
        lib.orc_pdhg_iterate(self.h, n)

    def residuals(self):
        out = (C.c_float * 6)()
        lib.orc_pdhg_residuals(self.h, out)
        keys = ["primal_residual", "dual_residual", "primal_var_norm", "dual_var_norm", "eps_primal", "eps_dual"]
        return dict(zip(keys, [float(v) for v in out]))

    def stepsizes(self):
        out = (C.c_double * 3)()
        lib.orc_pdhg_stepsizes(self.h, out)
        return float(out[0]), float(out[1]), float(out[2])

    def solution(self):
        n, m = self.prob.ncols, self.prob.nrows
        x, w = np.empty(n, np.float32), np.empty(n, np.float32)
        y, z = np.empty(m, np.float32), np.empty(m, np.float32)
        lib.orc_pdhg_solution(self.h, _p(x), _p(z), _p(y), _p(w))
        return x, z, y, w


class OracleADMM:
    """Oracle twin of BackendADMM with the option names of +backend/admm.m."""

    def __init__(self, prob, rho0=1.0, alpha=1.7, cg_tol_pow=1.3, cg_tol_min=1e-5, cg_tol_max=1e-8, cg_max_iter=10,
                 residual_iter=1, arb_delta=1.05, arb_tau=0.8, arb_gamma=1.01,
                 tol_rel_primal=1e-4, tol_rel_dual=1e-4, tol_abs_primal=1e-4, tol_abs_dual=1e-4):
        self.prob = prob
        self.h = C.c_void_p(lib.orc_admm_new(prob.h, rho0, alpha, cg_tol_pow, cg_tol_min, cg_tol_max, cg_max_iter,
                                             residual_iter, arb_delta, arb_tau, arb_gamma,
                                             tol_rel_primal, tol_rel_dual, tol_abs_primal, tol_abs_dual))

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_admm_free(self.h)
            self.h = None

    def initialize(self):
        if lib.orc_admm_init(self.h) != 0:
            raise RuntimeError(lib.orc_last_error().decode())

    def iterate(self, n=1):
        lib.orc_admm_iterate(self.h, n)

    def residuals(self):
        out = (C.c_float * 6)()
        lib.orc_admm_residuals(self.h, out)
        keys = ["primal_residual", "dual_residual", "primal_var_norm", "dual_var_norm", "eps_primal", "eps_dual"]
        return dict(zip(keys, [float(v) for v in out]))

    def stepsizes(self):
        """(rho, delta, total CG steps so far)"""
        out = (C.c_double * 3)()
        lib.orc_admm_stepsizes(self.h, out)
        return float(out[0]), float(out[1]), float(out[2])

    def solution(self):
        n, m = self.prob.ncols, self.prob.nrows
        x, w = np.empty(n, np.float32), np.empty(n, np.float32)
        y, z = np.empty(m, np.float32), np.empty(m, np.float32)
        lib.orc_admm_solution(self.h, _p(x), _p(z), _p(y), _p(w))
        return x, z, y, w


def num_threads():
    return lib.orc_num_threads()
