"""GPU suite, SURVEY.md 8(f) "next" rows (the widening steps after the hot path), all through the C ABI against
the oracle, the closed forms of the reference's own MATLAB tests and the live reference build:
  row 2: ProxTransform, elem_operation:ind_sum, ProxIndHalfspace, ProxIndSOC
  row 3: Kronecker blocks dense_kron_id, id_kron_dense, sparse_kron_id, id_kron_sparse
Runs last (file name): these were written after round 1's GPU budget was spent, so the round-end suite is their
first run."""
import zlib

import numpy as np
import pytest

import cases
import prost_b200 as pb
import ref_driver
from oracle_binding import oracle_prox_eval
from pdhg_util import assert_parity, run_cuda, run_oracle

pytestmark = pytest.mark.gpu

CASES = cases.prox_transform_cases()


def _inputs(name, n):
    r = np.random.default_rng(zlib.crc32(name.encode()))
    return (2 * r.standard_normal(n)).astype(np.float32), r.uniform(0.5, 1.5, n).astype(np.float32), 0.7


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("invert", [False, True])
def test_transform_matches_oracle_and_closed_form(ctx, name, invert):
    desc, n, direct = CASES[name]
    arg, tau_diag, tau = _inputs(name, n)
    got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau, invert)
    want = oracle_prox_eval(desc, arg, tau_diag, tau, invert)
    lo, hi = desc[1], desc[1] + desc[2]
    assert np.abs(got[lo:hi] - want[lo:hi]).max() <= 2e-5 * max(1.0, float(np.abs(want[lo:hi]).max())), name
    if direct is not None:      # test_prox_transform.m: same function written with direct coefficients
        ref = pb.create_prox(ctx, direct).Eval(arg, tau_diag, tau, invert)
        assert np.abs(got[lo:hi] - ref[lo:hi]).max() < 2e-4 * max(1.0, float(np.abs(ref[lo:hi]).max())), name


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref (compiled reference) not built")
@pytest.mark.parametrize("name", sorted(CASES))
def test_transform_vs_reference(ctx, name):
    desc, n, _ = CASES[name]
    arg, tau_diag, tau = _inputs(name, n)
    want = ref_driver.run_prox(desc, arg, tau_diag, tau)
    got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau)
    lo, hi = desc[1], desc[1] + desc[2]
    assert np.abs(got[lo:hi] - want[lo:hi]).max() <= 1e-5 * max(1.0, float(np.abs(want[lo:hi]).max())), name


def test_transform_errors(ctx):
    inner = pb.ProxElemOperation1D(ctx, "abs", 0, 10, 1, False, True, cases.coeffs())
    with pytest.raises(pb.ProstError) as e:
        pb.ProxTransform(ctx, inner, a=np.array([1, 0, 1, 1, 1, 1, 1, 1, 1, 1], np.float32))
    assert "isn't allowed to contain zero" in str(e.value)
    with pytest.raises(pb.ProstError):
        pb.ProxTransform(ctx, inner, b=np.zeros(3, np.float32))      # neither 1 nor size elements


def test_rof_written_with_transform_solves_like_direct(ctx):
    """+function/transform.m: sum_1d('square', 1, f, lmb) == transform(sum_1d('square'), 1, f, lmb).  The
    transformed prox is not a fusable leaf, so PDHG takes the unfused schedule; iterates agree with the oracle
    running the same description and with the fused direct formulation."""
    from prost_b200 import synthetic as syn
    nx, ny = 24, 20
    N = nx * ny
    desc = syn.rof(nx, ny)
    f = desc["data"]["f"]
    tdesc = dict(desc)
    tdesc["prox_g"] = [("transform", 0, N, True, [[1.0], f, [10.0], [0.0], [0.0],
                                                   ("elem_operation:1d:square", 0, N, True, [N, 1, False, cases.coeffs()])])]
    got = run_cuda(ctx, tdesc, 60, stepsize="alg1", residual_iter=5)
    assert not got["fused"]
    want = run_oracle(tdesc, 60, stepsize="alg1", residual_iter=5)
    assert_parity(got, want, label="ROF via transform")
    direct = run_cuda(ctx, desc, 60, stepsize="alg1", residual_iter=5)
    assert np.abs(got["x"] - direct["x"]).max() < 1e-4


# ---- elem_operation:ind_sum (second item) -----------------------------------------------------
SUM_CASES = cases.prox_ind_sum_cases()


@pytest.mark.parametrize("name", sorted(SUM_CASES))
def test_ind_sum_matches_oracle_and_reference(ctx, name):
    desc, n = SUM_CASES[name]
    arg, tau_diag, tau = _inputs(name, n)
    got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau)
    want = oracle_prox_eval(desc, arg, tau_diag, tau)
    lo, hi = desc[1], desc[1] + desc[2]
    assert np.abs(got[lo:hi] - want[lo:hi]).max() <= 1e-5, name
    if desc[0] == "elem_operation:ind_sum":          # test_prox_sum_ind_sum.m: groups sum to one
        count, dim, il = desc[4]
        G = got[lo:hi].reshape(count, dim) if il else got[lo:hi].reshape(dim, count).T
        assert np.abs(G.sum(axis=1) - 1).max() < 1e-5 * max(1, dim / 8), name
    if ref_driver.available():
        ref = ref_driver.run_prox(desc, arg, tau_diag, tau)
        assert np.abs(got[lo:hi] - ref[lo:hi]).max() <= 1e-5, name


# ---- ind_sum: ProxIndSum over index lists (last item) -------------------------------------
IDX_CASES = cases.prox_ind_sum_indexed_cases()


@pytest.mark.parametrize("name", sorted(IDX_CASES))
@pytest.mark.parametrize("invert", [False, True])
def test_ind_sum_indexed_matches_oracle_and_reference(ctx, name, invert):
    desc, n = IDX_CASES[name]
    arg, tau_diag, tau = _inputs(name, n)
    got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau, invert)
    want = oracle_prox_eval(desc, arg, tau_diag, tau, invert)
    lo, hi = desc[1], desc[1] + desc[2]
    # same float expressions in the same order: identical bits
    assert np.array_equal(got[lo:hi], want[lo:hi]), name
    data = desc[4]
    last = 3 if len(data) == 6 else 0              # the groups of the list that ran last sum to its constant
    dim, inds, total = int(data[last]), np.asarray(data[last + 1], np.int64), float(data[last + 2])
    if name != "ind_sum_idx_second_list_truncated":
        sums = got[lo:hi][inds].reshape(-1, dim).sum(axis=1)
        assert np.abs(sums - total).max() < 1e-4 * max(1.0, dim / 8), name
    listed = np.concatenate([np.asarray(data[k], np.int64).ravel() for k in range(1, len(data), 3)])
    untouched = np.setdiff1d(np.arange(hi - lo), listed)
    assert np.array_equal(got[lo:hi][untouched], arg[lo:hi][untouched])
    if ref_driver.available() and not invert:
        ref = ref_driver.run_prox(desc, arg, tau_diag, tau)
        assert np.array_equal(got[lo:hi], ref[lo:hi]), name


def test_ind_sum_indexed_errors(ctx):
    with pytest.raises(pb.ProstError) as e:
        pb.ProxIndSum(ctx, 0, 10, 2, np.array([0, 1, 2, 10], np.uint64), 1.0)
    assert "outside the prox range" in str(e.value)


# ---- spectral element operations ---------------------------------------------------------
SPEC_CASES = cases.prox_spectral_cases()


@pytest.mark.parametrize("name", sorted(SPEC_CASES))
def test_spectral_prox_matches_oracle_and_reference(ctx, name):
    """singular_nx2 / eigen_2x2 / eigen_3x3 / eigen_nxn: the CUDA kernels (Sylvester's formula, Jacobi rotations)
    against the oracle and the live reference (dlaev2-style 2 x 2, Kopp's 3 x 3, EISPACK n x n): same spectral
    function, different eigen-solvers -> agreement to rounding, not bit for bit."""
    desc, n = SPEC_CASES[name]
    arg, tau_diag, tau = _inputs(name, n)
    arg = (1.5 * arg).astype(np.float32)
    for invert in (False, True):
        got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau, invert)
        want = oracle_prox_eval(desc, arg, tau_diag, tau, invert)
        lo, hi = desc[1], desc[1] + desc[2]
        scale = max(1.0, float(np.abs(want[lo:hi]).max()))
        assert np.abs(got[lo:hi] - want[lo:hi]).max() <= 2e-5 * scale, (name, invert)
    if ref_driver.available():
        got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau)
        ref = ref_driver.run_prox(desc, arg, tau_diag, tau)
        assert np.abs(got[lo:hi] - ref[lo:hi]).max() <= 2e-5 * scale, name


KRON_TC_CASES = cases.linop_kron_tensor_core_cases()


@pytest.mark.parametrize("name", sorted(KRON_TC_CASES))
def test_kron_tensor_core_path_keeps_fp32_parity(ctx, name):
    """pb_kron_tc.cu: tcgen05 kind::tf32 with the 3 x TF32 split.  Same 1e-5 (max norm) as the fp32 kernels, against a
    float64 product, the oracle and the live reference; forward, adjoint, overwrite and accumulate."""
    from oracle_binding import OracleProblem
    blocks = KRON_TC_CASES[name]
    op = pb.create_linop(ctx, blocks)
    m, n = op.nrows, op.ncols
    r = np.random.default_rng(zlib.crc32(name.encode()))
    x, y = r.standard_normal(n).astype(np.float32), r.standard_normal(m).astype(np.float32)
    fwd, adj = op.Eval(x), op.EvalAdjoint(y)
    wf, wa = cases.kron_apply_f64(blocks, x, False), cases.kron_apply_f64(blocks, y, True)
    assert np.abs(fwd - wf).max() <= 1e-5 * max(1.0, float(np.abs(wf).max())), (name, np.abs(fwd - wf).max())
    assert np.abs(adj - wa).max() <= 1e-5 * max(1.0, float(np.abs(wa).max())), (name, np.abs(adj - wa).max())
    orc = OracleProblem(blocks=blocks)
    assert np.abs(fwd - orc.linop(x, False)).max() <= 1e-5 * max(1.0, float(np.abs(wf).max())), name
    assert np.abs(adj - orc.linop(y, True)).max() <= 1e-5 * max(1.0, float(np.abs(wa).max())), name
    if ref_driver.available():
        ref_f = ref_driver.run_linop(blocks, x, False)["res"]
        ref_a = ref_driver.run_linop(blocks, y, True)["res"]
        assert np.abs(fwd - ref_f).max() <= 1e-5 * max(1.0, float(np.abs(ref_f).max())), name
        assert np.abs(adj - ref_a).max() <= 1e-5 * max(1.0, float(np.abs(ref_a).max())), name


# ---- ind_range: projection onto the range of a sparse matrix ------------------------------
RANGE_CASES = cases.prox_ind_range_cases()


@pytest.mark.parametrize("name", sorted(RANGE_CASES))
def test_ind_range_matches_oracle_closed_form_and_reference(ctx, name):
    """x = A (A^T A)^{-1} A^T x0: three parallel kernels around a host-side inverse here, csrmv + potrs + csrmv in the
    reference; test_prox_ind_range.m's closed form at its 1e-4 (norm), oracle and live reference likewise."""
    desc, n = RANGE_CASES[name]
    arg, tau_diag, tau = _inputs(name, n)
    got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau)
    lo, hi = desc[1], desc[1] + desc[2]
    A = desc[4][0].toarray().astype(np.float64)
    want = A @ np.linalg.solve(A.T @ A, A.T @ arg[lo:hi].astype(np.float64))
    scale = max(1.0, float(np.linalg.norm(want)))
    assert np.linalg.norm(got[lo:hi] - want) <= 1e-4 * scale, name
    orc = oracle_prox_eval(desc, arg, tau_diag, tau)
    assert np.linalg.norm(got[lo:hi] - orc[lo:hi]) <= 1e-4 * scale, name
    if ref_driver.available():
        ref = ref_driver.run_prox(desc, arg, tau_diag, tau)
        assert np.linalg.norm(got[lo:hi] - ref[lo:hi]) <= 1e-4 * scale, name


# ---- ind_epi_conjquad_1d: the north star's ProxEpiConjQuadr (source external to the reference tree) -------------
CONJ_CASES = cases.prox_epi_conjquad_cases()


@pytest.mark.parametrize("name", sorted(CONJ_CASES))
def test_epi_conjquad_matches_oracle_and_bruteforce(ctx, name):
    """PARITY UNPINNED (no reference source): the CUDA kernel against the oracle restatement (same float
    expressions) and, on a sample, against the independent double-precision brute-force projection."""
    import refmath
    desc, n = CONJ_CASES[name]
    arg, tau_diag, tau = _inputs(name, n)
    arg = (1.5 * arg).astype(np.float32)
    got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau)
    want = oracle_prox_eval(desc, arg, tau_diag, tau)
    lo, hi = desc[1], desc[1] + desc[2]
    scale = max(1.0, float(np.abs(want[lo:hi]).max()))
    assert np.abs(got[lo:hi] - want[lo:hi]).max() <= 2e-5 * scale, name
    count, il, co = desc[4]
    G, A = got[lo:hi], arg[lo:hi]
    gx, gy = (G[0::2], G[1::2]) if il else (G[:count], G[count:])
    ax, ay = (A[0::2], A[1::2]) if il else (A[:count], A[count:])
    at = lambda k, i: float(np.atleast_1d(co[k])[i if np.atleast_1d(co[k]).size > 1 else 0])
    for i in range(0, count, 7):
        wx, wy = refmath.project_epi_conjquad_1d_bruteforce(float(ax[i]), float(ay[i]), at(0, i), at(1, i), at(2, i),
                                                            at(3, i), at(4, i))
        s = max(1.0, abs(wx), abs(wy))
        assert abs(gx[i] - wx) <= 2e-4 * s and abs(gy[i] - wy) <= 2e-4 * s, (name, i)


def test_epi_conjquad_in_a_pdhg_solve_matches_oracle(ctx):
    """Sublabel-style dual constraint inside PDHG: K = identity block, f* = ind_epi_conjquad_1d on (x, y) pairs,
    g = quadratic; the unfused schedule against the oracle running the same description."""
    r = np.random.default_rng(3)
    n = 600
    a = r.uniform(0.3, 2.0, n // 2).astype(np.float32)
    b = r.uniform(-1, 1, n // 2).astype(np.float32)
    c = r.uniform(-0.5, 0.5, n // 2).astype(np.float32)
    lo = r.uniform(-1, 0, n // 2).astype(np.float32)
    hi = (lo + r.uniform(0.2, 1.5, n // 2)).astype(np.float32)
    f = r.standard_normal(n).astype(np.float32)
    desc = dict(nrows=n, ncols=n, blocks=[("diags", 0, 0, [n, n, [1.0], [0]])],
                prox_g=[("elem_operation:1d:square", 0, n, True, [n, 1, False, cases.coeffs(a=1, b=f, c=2.0)])],
                prox_fstar=[("ind_epi_conjquad_1d", 0, n, False, [n // 2, False, [a, b, c, lo, hi]])],
                scaling=("alpha", 1.0))
    got = run_cuda(ctx, desc, 80, stepsize="alg1", residual_iter=4)
    want = run_oracle(desc, 80, stepsize="alg1", residual_iter=4)
    assert_parity(got, want, iter_tol=5e-5, label="epi_conjquad in PDHG")


# ---- ind_halfspace, ind_soc -------------------------------------------------------------
PROJ_CASES = cases.prox_projection_cases()


@pytest.mark.parametrize("name", sorted(PROJ_CASES))
def test_projection_prox_matches_oracle_and_reference(ctx, name):
    desc, n = PROJ_CASES[name]
    arg, tau_diag, tau = _inputs(name, n)
    got = pb.create_prox(ctx, desc).Eval(arg, tau_diag, tau)
    want = oracle_prox_eval(desc, arg, tau_diag, tau)
    lo, hi = desc[1], desc[1] + desc[2]
    scale = max(1.0, float(np.abs(want[lo:hi]).max()))
    assert np.abs(got[lo:hi] - want[lo:hi]).max() <= 2e-5 * scale, name
    if ref_driver.available():
        ref = ref_driver.run_prox(desc, arg, tau_diag, tau)
        assert np.abs(got[lo:hi] - ref[lo:hi]).max() <= 1e-5 * scale, name


def test_projection_prox_errors(ctx):
    with pytest.raises(pb.ProstError) as e:
        pb.ProxIndSOC(ctx, 0, 10, 3, False, False, alpha=2.0)
    assert "Only alpha = 1" in str(e.value)
    with pytest.raises(pb.ProstError) as e:
        pb.ProxIndHalfspace(ctx, 0, 10, 3, False, False, np.ones(7, np.float32), np.ones(1, np.float32))
    assert "Coefficient a has to have dimension count*dim or dim" in str(e.value)


# ---- Kronecker blocks dense_kron_id / id_kron_dense / sparse_kron_id / id_kron_sparse ------------------------------------------
KRON_CASES = cases.linop_kron_cases()


@pytest.mark.parametrize("name", sorted(KRON_CASES))
def test_kron_blocks_forward_adjoint(ctx, name):
    from oracle_binding import OracleProblem
    blocks = KRON_CASES[name]
    op = pb.create_linop(ctx, blocks)
    m, n = op.nrows, op.ncols
    r = np.random.default_rng(zlib.crc32(name.encode()))
    x, y = r.standard_normal(n).astype(np.float32), r.standard_normal(m).astype(np.float32)
    orc = OracleProblem(blocks=blocks)
    fwd, adj = op.Eval(x), op.EvalAdjoint(y)
    assert np.abs(fwd - orc.linop(x, False)).max() <= 1e-5 * max(1.0, float(np.abs(fwd).max())), name
    assert np.abs(adj - orc.linop(y, True)).max() <= 1e-5 * max(1.0, float(np.abs(adj).max())), name
    # the reference tests compare with kron() at 1e-3
    M = np.zeros((m, n))
    for (bname, row, col, (K, d)) in blocks:
        K = K.toarray() if hasattr(K, "toarray") else K
        full = np.kron(np.eye(d), K) if bname.startswith("id_") else np.kron(K, np.eye(d))
        M[row:row + full.shape[0], col:col + full.shape[1]] += full
    assert np.abs(fwd - M @ x).max() < 1e-3 and np.abs(adj - M.T @ y).max() < 1e-3, name
    assert abs(float(y @ fwd) - float(adj @ x)) <= 1e-3 * max(1.0, abs(float(y @ fwd))), name      # <Kx, y> = <x, K^T y>
    assert np.allclose(op.row_sums(1.0), np.abs(M).sum(1), rtol=1e-5, atol=1e-6)
    assert np.allclose(op.col_sums(1.0), np.abs(M).sum(0), rtol=1e-5, atol=1e-6)
    if ref_driver.available():
        ref_f = ref_driver.run_linop(blocks, x, False)["res"]
        ref_a = ref_driver.run_linop(blocks, y, True)["res"]
        assert np.abs(fwd - ref_f).max() <= 1e-5 * max(1.0, float(np.abs(ref_f).max())), name
        assert np.abs(adj - ref_a).max() <= 1e-5 * max(1.0, float(np.abs(ref_a).max())), name
