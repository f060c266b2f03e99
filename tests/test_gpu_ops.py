"""GPU suite: operator applies and proxes through the C ABI (host-vector entry points, the
analogue of mex eval_linop / eval_prox, prost.cpp:157-276) against the CPU oracle and the
reference tests' closed forms."""
import zlib

import numpy as np
import pytest

import cases
import refmath
import prost_b200 as pb
from oracle_binding import OracleProblem, oracle_prox_eval

pytestmark = pytest.mark.gpu

LINOPS = cases.linop_cases(small=False)
PROXES = cases.prox_cases(small=False)


def close(a, b, tol=1e-5):
    """|a - b| <= tol * max(1, |b|) elementwise; returns the fraction of violations."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    bad = np.abs(a - b) > tol * np.maximum(1.0, np.abs(b))
    return bad.mean() if bad.size else 0.0


@pytest.mark.parametrize("name", sorted(LINOPS))
def test_linop_forward_adjoint(ctx, name):
    blocks = LINOPS[name]
    op = pb.create_linop(ctx, blocks)
    orc = OracleProblem(blocks=blocks)
    m, n = orc.linop_size()
    assert (op.nrows, op.ncols) == (m, n)
    r = np.random.default_rng(3)
    x, y = r.random(n).astype(np.float32), r.random(m).astype(np.float32)
    fwd, adj = op.Eval(x), op.EvalAdjoint(y)
    exact = not any(b[0] in ("dense", "sparse") for b in blocks)
    if exact:
        # stencils and diagonals do the same float operations in the same order as the oracle;
        # only FMA contraction of the diagonal products (29 terms per row) differs
        tol = 1e-5 if any(b[0] == "diags" for b in blocks) else 1e-6
        assert close(fwd, orc.linop(x, False), tol) == 0
        assert close(adj, orc.linop(y, True), tol) == 0
    else:
        assert close(fwd, orc.linop(x, False), 1e-4) == 0
        assert close(adj, orc.linop(y, True), 1e-4) == 0
    K = cases.linop_matrix(blocks)                      # the reference tests' own bound
    assert np.linalg.norm(fwd - K @ x.astype(np.float64)) <= 1e-3
    assert np.linalg.norm(adj - K.T @ y.astype(np.float64)) <= 1e-3
    for alpha in (1.0, 0.5, 2.0):
        assert np.array_equal(op.row_sums(alpha), orc.row_sums(alpha))
        assert np.array_equal(op.col_sums(alpha), orc.col_sums(alpha))


def test_linop_overlap_is_rejected(ctx):
    op = pb.LinearOperator(ctx)
    op.AddBlock(pb.BlockZero(ctx, 0, 0, 10, 10))
    op.AddBlock(pb.BlockGradient2D(ctx, 5, 5, 4, 4, 1))
    with pytest.raises(pb.ProstError) as e:
        op.Initialize()
    assert "overlapping" in str(e.value)


@pytest.mark.parametrize("name", sorted(PROXES))
@pytest.mark.parametrize("invert", [False, True])
def test_prox_matches_oracle(ctx, name, invert):
    desc, n = PROXES[name]
    r = np.random.default_rng(zlib.crc32(name.encode()))
    arg = (2 * r.standard_normal(n)).astype(np.float32)
    tau_diag = r.uniform(0.5, 1.5, n).astype(np.float32)
    tau = 0.7
    prox = pb.create_prox(ctx, desc)
    got = prox.Eval(arg, tau_diag, tau, invert)
    want = oracle_prox_eval(desc, arg, tau_diag, tau, invert)
    lo, hi = desc[1], desc[1] + desc[2]
    frac = close(got[lo:hi], want[lo:hi], 2e-5)
    jumpy = any(k in name for k in ("l0", "truncquad", "trunclin", "lq"))
    assert frac <= (2e-3 if jumpy else 0.0), (name, frac)


def test_prox_norm2_ball_reference_test(ctx):
    """test_prox_sum_norm2.m verbatim: N = 6000, d = 7, inf-norm 1e-5."""
    N, d = 6000, 7
    r = np.random.default_rng(1)
    P = (-2 + 4 * r.random((N, d))).astype(np.float32)
    prox = pb.ProxElemOperationNorm2(ctx, "ind_leq0", 0, N, d, False, False, cases.coeffs(a=1, b=1, c=1))
    Q = prox.Eval(P.T.ravel(), np.ones(N * d), 1.0).reshape(d, N).T
    assert np.abs(Q - refmath.norm2_ball(P.astype(np.float64))).max() < 1e-5


def test_prox_simplex_reference_test(ctx):
    """test_prox_sum_ind_simplex.m verbatim: N = 1000, d = 289 planar vs projsplx; plus the
    register-resident sizes (d <= 64)."""
    for N, d in ((1000, 289), (4000, 32), (2000, 64), (3000, 3)):
        r = np.random.default_rng(d)
        P = (-2 + 4 * r.random((N, d))).astype(np.float32)
        prox = pb.ProxElemOperationIndSimplex(ctx, 0, N, d, False, False)
        Q = prox.Eval(P.T.ravel(), np.ones(N * d), 1.0).reshape(d, N).T
        assert np.abs(Q - refmath.projsplx_rows(P)).max() < 1e-5


def test_prox_permute_is_bit_exact(ctx):
    """Index work must be bit-exact (north star): gather/scatter round trip and equivalence with
    the un-permuted separable prox (test_prox_permute.m)."""
    N = 100003
    r = np.random.default_rng(8)
    arg = r.standard_normal(N).astype(np.float32)
    perm = r.permutation(N).astype(np.int32)
    ident = pb.ProxPermute(ctx, pb.ProxZero(ctx, 0, N), perm)
    assert np.array_equal(ident.Eval(arg, np.ones(N), 1.0), arg)
    c = cases.coeffs(c=0.5)
    plain = pb.ProxElemOperation1D(ctx, "abs", 0, N, 1, False, True, c)
    permuted = pb.ProxPermute(ctx, pb.ProxElemOperation1D(ctx, "abs", 0, N, 1, False, True, c), perm)
    assert np.array_equal(permuted.Eval(arg, np.ones(N), 1.0), plain.Eval(arg, np.ones(N), 1.0))


def test_prox_errors(ctx):
    with pytest.raises(pb.ProstError):
        pb.ProxIndEpiQuad(ctx, 0, 10, 2, False, False, [-1.0], np.zeros(10), [0.0])      # a <= 0
    with pytest.raises(pb.ProstError):
        pb.ProxIndEpiQuad(ctx, 0, 10, 2, False, False, [1.0], np.zeros(7), [0.0])        # |b| != count*(dim-1)
    with pytest.raises(pb.ProstError):
        pb.ProxPermute(ctx, pb.ProxZero(ctx, 0, 10), np.arange(9))                       # wrong length
    with pytest.raises(pb.ProstError):
        pb.ProxElemOperation1D(ctx, "nope", 0, 4, 1, False, True, cases.coeffs())
