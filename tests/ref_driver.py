"""Runs the UNMODIFIED reference (oracle/_ref/prost_ref_driver: tum-vision/prost compiled for
sm_100 + oracle/driver/prost_driver.cu) on a problem description -- TEST INFRASTRUCTURE.

The same description dicts as prost_b200.factory / oracle_binding are serialised to the driver's
text format plus raw array files."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "prost_ref_driver")
OUR_DRIVER = os.path.join(ROOT, "prost_b200", "lib", "prost_b200_driver")

_STEPS = {"alg1": 1, "alg2": 2, "goldstein": 3, "boyd": 4}


def available(binary=REF_DRIVER):
    return os.path.exists(binary) and os.access(binary, os.X_OK)


class _Writer:
    def __init__(self, d):
        self.d = d
        self.n = 0
        self.lines = []

    def arr(self, a, dtype):
        a = np.ascontiguousarray(np.asarray(a, dtype=dtype).ravel())
        name = f"a{self.n}.bin"
        self.n += 1
        a.tofile(os.path.join(self.d, name))
        return name, a.size

    def coeff(self, c):
        c = np.atleast_1d(np.asarray(c, dtype=np.float32)).ravel()
        if c.size == 1:
            return f"s:{float(c[0])!r}"
        name, n = self.arr(c, np.float32)
        return f"f:{name}:{n}"

    def block(self, desc):
        name, row, col, data = desc
        if name in ("gradient2d", "gradient3d"):
            nx, ny, L, lf = data
            return f"block {name} {row} {col} {nx} {ny} {L} {int(lf)}"
        if name == "diags":
            nr, nc, fac, ofs = data
            fo, n = self.arr(ofs, np.int64)
            ff, _ = self.arr(fac, np.float32)
            return f"block diags {row} {col} {nr} {nc} {n} {fo} {ff}"
        if name == "sparse":
            import scipy.sparse as sp
            A = sp.csc_matrix(data[0])
            A.sort_indices()
            fv, _ = self.arr(A.data, np.float32)
            fp, _ = self.arr(A.indptr, np.int32)
            fi, _ = self.arr(A.indices, np.int32)
            return f"block sparse {row} {col} {A.shape[0]} {A.shape[1]} {A.nnz} {fv} {fp} {fi}"
        if name == "dense":
            A = np.asarray(data[0], dtype=np.float32)
            fd, _ = self.arr(np.ascontiguousarray(A.T), np.float32)
            return f"block dense {row} {col} {A.shape[0]} {A.shape[1]} {fd}"
        if name in ("sparse_kron_id", "id_kron_sparse"):
            import scipy.sparse as sp
            A = sp.csc_matrix(data[0])
            A.sort_indices()
            fv, _ = self.arr(A.data, np.float32)
            fp, _ = self.arr(A.indptr, np.int32)
            fi, _ = self.arr(A.indices, np.int32)
            return f"block {name} {row} {col} {int(data[1])} {A.shape[0]} {A.shape[1]} {A.nnz} {fv} {fp} {fi}"
        if name in ("dense_kron_id", "id_kron_dense"):
            K = np.asarray(data[0], dtype=np.float32)
            fd, _ = self.arr(np.ascontiguousarray(K.T), np.float32)
            return f"block {name} {row} {col} {K.shape[0]} {K.shape[1]} {int(data[1])} {fd}"
        if name == "zero":
            return f"block zero {row} {col} {data[0]} {data[1]}"
        raise ValueError(name)

    def prox(self, desc):
        name, idx, size, ds, data = desc
        if name.startswith("elem_operation:1d:") or name.startswith("elem_operation:norm2:"):
            count, dim, il, coeffs = data
            kind = "norm2" if ":norm2:" in name else "elem1d"
            cs = " ".join(self.coeff(c) for c in coeffs)
            return f"{kind} {name.split(':')[2]} {idx} {count} {dim} {int(il)} {int(ds)} {cs}"
        for kind in ("singular_nx2", "eigen_2x2", "eigen_3x3", "eigen_nxn"):
            prefix = f"elem_operation:{kind}:"
            if name.startswith(prefix):
                count, dim, il, coeffs = data
                cs = " ".join(self.coeff(c) for c in coeffs)
                return f"spectral {kind} {name[len(prefix):]} {idx} {count} {dim} {int(il)} {int(ds)} {cs}"
        if name in ("elem_operation:mass4", "elem_operation:ind_comass4_ball", "elem_operation:mass5",
                    "elem_operation:ind_comass5_ball"):
            count, dim, il = data[:3]
            cost = self.coeff(data[3][0]) if len(data) > 3 else "s:1.0"
            return f"massnorm {name.split(':')[1]} {idx} {count} {dim} {int(il)} {int(ds)} {cost}"
        if name == "elem_operation:ind_simplex":
            count, dim, il = data[:3]
            return f"simplex {idx} {count} {dim} {int(il)} {int(ds)}"
        if name == "ind_halfspace":
            count, dim, il, (a, b) = data
            return f"halfspace {idx} {count} {dim} {int(il)} {int(ds)} {self.coeff(a)} {self.coeff(b)}"
        if name == "ind_soc":
            count, dim, il = data[:3]
            return f"soc {idx} {count} {dim} {int(il)} {int(ds)} {float(data[3]) if len(data) > 3 else 1.0!r}"
        if name == "ind_range":
            import scipy.sparse as sp
            A = sp.csc_matrix(data[0])
            A.sort_indices()
            AA = np.asarray(data[1] if len(data) > 1 and data[1] is not None else (A.T @ A).toarray(), np.float32)
            fv, _ = self.arr(A.data, np.float32)
            fp, _ = self.arr(A.indptr, np.int32)
            fi, _ = self.arr(A.indices, np.int32)
            fa, _ = self.arr(np.ascontiguousarray(AA.T), np.float32)
            return f"indrange {idx} {size} {int(ds)} {A.shape[0]} {A.shape[1]} {A.nnz} {fv} {fp} {fi} {fa}"
        if name == "ind_sum":
            parts = [f"indsumidx {idx} {size} {len(data) // 3}"]
            for k in range(0, len(data), 3):
                fi, n = self.arr(data[k + 1], np.uint64)
                parts.append(f"{int(data[k])} {fi} {n} {float(data[k + 2])!r}")
            return " ".join(parts)
        if name == "elem_operation:ind_sum":
            count, dim, il = data[:3]
            return f"indsum {idx} {count} {dim} {int(il)} {int(ds)}"
        if name == "ind_epi_quad":
            count, dim, il, (a, b, c) = data
            return f"epiquad {idx} {count} {dim} {int(il)} {int(ds)} {self.coeff(a)} {self.coeff(b)} {self.coeff(c)}"
        if name == "moreau":
            return "moreau " + self.prox(data[0])
        if name == "permute":
            fp, n = self.arr(data[1], np.int32)
            return f"permute {fp} {n} " + self.prox(data[0])
        if name == "transform":
            return "transform " + " ".join(self.coeff(c) for c in data[:5]) + " " + self.prox(data[5])
        if name == "zero":
            return f"zero {idx} {size}"
        raise ValueError(name)


def _run(binary, d, lines, timeout=600):
    spec = os.path.join(d, "problem.txt")
    with open(spec, "w") as f:
        f.write("\n".join(lines) + "\n")
    r = subprocess.run([binary, spec], capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"{os.path.basename(binary)} failed ({r.returncode}): {r.stderr[-2000:]}")
    info = {}
    with open(os.path.join(d, "out_info.txt")) as f:
        for ln in f:
            k, v = ln.split()
            info[k] = float(v)
    return info


def run_linop(blocks, rhs, transpose, binary=REF_DRIVER):
    with tempfile.TemporaryDirectory() as d:
        w = _Writer(d)
        lines = [w.block(b) for b in blocks]
        fin, _ = w.arr(rhs, np.float32)
        lines += [f"action linop {fin} {int(transpose)} -", f"out {d}/out"]
        info = _run(binary, d, lines)
        return dict(res=np.fromfile(f"{d}/out_res.f32", np.float32),
                    rowsum=np.fromfile(f"{d}/out_rowsum.f32", np.float32),
                    colsum=np.fromfile(f"{d}/out_colsum.f32", np.float32), info=info)


def run_prox(desc, arg, tau_diag, tau, binary=REF_DRIVER):
    with tempfile.TemporaryDirectory() as d:
        w = _Writer(d)
        lines = ["prox eval " + w.prox(desc)]
        fa, _ = w.arr(arg, np.float32)
        ft, _ = w.arr(tau_diag, np.float32)
        lines += [f"action prox {fa} {ft} {float(tau)!r}", f"out {d}/out"]
        _run(binary, d, lines)
        return np.fromfile(f"{d}/out_res.f32", np.float32)


def run_solve(desc, iters, x0=None, y0=None, tol=None, binary=REF_DRIVER, tau0=1.0, sigma0=1.0, residual_iter=1,
              alg2_gamma=0.0, arg_alpha0=0.5, arg_nu=0.95, arg_delta=1.5, arb_delta=1.05, arb_tau=0.8,
              stepsize="boyd", scale_steps_operator=0, num_cback_calls=0, timeout=1200, admm=None,
              solve_dual_problem=False):
    """admm: None for BackendPDHG, or a dict of BackendADMM options (rho0, alpha, cg_tol_pow, cg_tol_min,
    cg_tol_max, cg_max_iter, residual_iter, arb_delta, arb_tau, arb_gamma; defaults of +backend/admm.m)."""
    tol = tol or dict(tol_rel_primal=0.0, tol_rel_dual=0.0, tol_abs_primal=0.0, tol_abs_dual=0.0)
    with tempfile.TemporaryDirectory() as d:
        w = _Writer(d)
        lines = [f"dims {desc['nrows']} {desc['ncols']}"]
        lines += [w.block(b) for b in desc["blocks"]]
        for key, tag in (("prox_g", "g"), ("prox_f", "f"), ("prox_gstar", "gstar"), ("prox_fstar", "fstar")):
            for p in desc.get(key, []):
                lines.append(f"prox {tag} " + w.prox(p))
        sc = desc.get("scaling", ("alpha", 1.0))
        if sc[0] == "alpha":
            lines.append(f"scaling alpha {float(sc[1])!r}")
        elif sc[0] == "identity":
            lines.append("scaling identity")
        else:
            fl, _ = w.arr(sc[1], np.float32)
            fr, _ = w.arr(sc[2], np.float32)
            lines.append(f"scaling custom {fl} {fr}")
        lines.append(f"pdhg {tau0!r} {sigma0!r} {residual_iter} {int(scale_steps_operator)} {alg2_gamma!r} "
                     f"{arg_alpha0!r} {arg_nu!r} {arg_delta!r} {arb_delta!r} {arb_tau!r} {_STEPS[stepsize]}")
        if admm is not None:
            a = dict(rho0=1.0, alpha=1.7, cg_tol_pow=1.3, cg_tol_min=1e-5, cg_tol_max=1e-8, cg_max_iter=10,
                     residual_iter=1, arb_delta=1.05, arb_tau=0.8, arb_gamma=1.01)
            a.update(admm)
            lines.append(f"admm {a['rho0']!r} {a['alpha']!r} {a['cg_tol_pow']!r} {a['cg_tol_min']!r} "
                         f"{a['cg_tol_max']!r} {a['cg_max_iter']} {a['residual_iter']} {a['arb_delta']!r} "
                         f"{a['arb_tau']!r} {a['arb_gamma']!r}")
        fx = w.arr(x0, np.float32)[0] if x0 is not None else "-"
        fy = w.arr(y0, np.float32)[0] if y0 is not None else "-"
        lines.append(f"solver {tol['tol_rel_primal']!r} {tol['tol_rel_dual']!r} {tol['tol_abs_primal']!r} "
                     f"{tol['tol_abs_dual']!r} {iters} {num_cback_calls} {fx} {fy} {int(bool(solve_dual_problem))}")
        lines += ["action solve - - -", f"out {d}/out"]
        info = _run(binary, d, lines, timeout=timeout)
        out = {k: np.fromfile(f"{d}/out_{k}.f32", np.float32) for k in ("x", "y", "z", "w")}
        out["res"] = {k: info[k] for k in ("primal_residual", "dual_residual", "primal_var_norm",
                                           "dual_var_norm", "eps_primal", "eps_dual")}
        out["info"] = info
        return out
