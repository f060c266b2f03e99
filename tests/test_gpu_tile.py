"""GPU suite: the whole-iteration tiled kernel (prost_b200/csrc/pb_tile.cu, one HBM pass per PDHG
iteration, 28 B/pixel instead of 44) must be BIT-IDENTICAL to the two-pass specialised kernels
(fuse=3), and through them within the north-star bars of the oracle / reference.  Sizes cover
partial tiles in both directions (tiles are 32 columns x 128 rows), single-row-vector images, the
iteration-0 and residual-refresh hand-over to the two-pass kernels, every step-size rule, several
Function1D members on both proxes and several label planes."""
import copy

import numpy as np
import pytest

import prost_b200 as pb
from pdhg_util import TOL, assert_parity, run_cuda, run_oracle
from prost_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def same_residuals(a, b):
    """The tiled residual-refresh pass folds its double partial sums in a different order than the two-pass
    kernels: the float residuals may differ in the last digit."""
    return all(abs(a["res"][k] - b["res"][k]) <= 2e-6 * max(abs(b["res"][k]), 1e-30) for k in b["res"])

TOL4 = dict(tol_rel_primal=1e-4, tol_rel_dual=1e-4, tol_abs_primal=1e-4, tol_abs_dual=1e-4)


def tiled_iterations(ctx, desc, **opts):
    prob = pb.create_problem(ctx, desc)
    be = pb.BackendPDHG(ctx, prob, pb.pdhg_options(scale_steps_operator=0, fuse=1, **opts),
                        pb.solver_options(verbose=0, max_iters=10, **TOL))
    prob.Initialize()
    be.Initialize()
    d = be.profile_detail(10)
    return d["n_tile"] + d["n_tile_check"]


def with_g(desc, fn, **coeff):
    """swap the Function1D member / weights of prox_g"""
    d = copy.deepcopy(desc)
    name, idx, size, ds, (count, dim, il, c) = d["prox_g"][0]
    names = ["a", "b", "c", "d", "e", "alpha", "beta"]
    for k, v in coeff.items():
        c[names.index(k)] = np.atleast_1d(np.asarray(v, np.float32))
    d["prox_g"][0] = (f"elem_operation:1d:{fn}", idx, size, ds, [count, dim, il, c])
    return d


def with_f(desc, fn, **coeff):
    d = copy.deepcopy(desc)
    name, idx, size, ds, (count, dim, il, c) = d["prox_fstar"][0]
    names = ["a", "b", "c", "d", "e", "alpha", "beta"]
    for k, v in coeff.items():
        c[names.index(k)] = np.atleast_1d(np.asarray(v, np.float32))
    d["prox_fstar"][0] = (f"elem_operation:norm2:{fn}", idx, size, ds, [count, dim, il, c])
    return d


def channelwise_tv(nx, ny, nc):
    """nc label planes, Norm2 per voxel over (gx, gy): channel-by-channel TV (grid.y = label)"""
    d = syn.tvl1(nx, ny, nc=nc)
    N = nx * ny * nc
    name, idx, size, ds, (count, dim, il, c) = d["prox_fstar"][0]
    d["prox_fstar"][0] = (name, idx, size, ds, [N, 2, il, c])
    return d


SHAPES = [(32, 128), (64, 256), (70, 260), (33, 132), (31, 124), (1, 4), (5, 8), (96, 4), (200, 1000)]


@pytest.mark.parametrize("shape", SHAPES)
def test_tile_bit_identical_to_two_pass(ctx, shape):
    desc = syn.rof(*shape)
    opts = dict(stepsize="alg1", residual_iter=4)
    assert tiled_iterations(ctx, desc, **opts) > 0, "tiled path not selected"
    a = run_cuda(ctx, desc, 37, fuse=1, **opts)
    b = run_cuda(ctx, desc, 37, fuse=3, **opts)
    for k in ("x", "y", "z", "w"):
        assert np.array_equal(a[k], b[k]), k
    assert same_residuals(a, b) and a["steps"] == b["steps"]


@pytest.mark.parametrize("stepsize", ["alg1", "alg2", "goldstein", "boyd"])
def test_tile_all_stepsizes(ctx, stepsize):
    desc = syn.rof(70, 260)
    opts = dict(stepsize=stepsize, residual_iter=3, alg2_gamma=0.5)
    a = run_cuda(ctx, desc, 150, fuse=1, tol=TOL4, **opts)
    b = run_cuda(ctx, desc, 150, fuse=3, tol=TOL4, **opts)
    for k in ("x", "y", "z", "w"):
        assert np.array_equal(a[k], b[k]), k
    assert same_residuals(a, b) and a["steps"] == b["steps"]
    want = run_oracle(desc, 150, tol=TOL4, **opts)
    loose = stepsize == "alg2"
    assert_parity(a, want, iter_tol=5e-5 if loose else 1e-5, res_tol=5e-3 if loose else 1e-4, label=stepsize)


@pytest.mark.parametrize("variant", ["abs_g", "huber_g_vec_b", "general_weights", "abs_f_ball", "scalar_b",
                                     "channels3", "warm_start", "residual_iter_1", "residual_never"])
def test_tile_variants(ctx, variant):
    desc = syn.rof(70, 132)
    opts = dict(stepsize="boyd", residual_iter=5)
    x0 = y0 = None
    if variant == "abs_g":
        desc = with_g(desc, "abs", c=1.0)
    elif variant == "huber_g_vec_b":
        desc = with_g(desc, "huber", c=2.0, alpha=0.05)
    elif variant == "general_weights":
        desc = with_g(desc, "square", a=1.5, c=3.0, d=0.1, e=0.2)
    elif variant == "abs_f_ball":
        desc = with_f(desc, "abs", a=0.7, b=0.1, c=2.0)
    elif variant == "scalar_b":
        desc = with_g(desc, "square", b=0.25)
    elif variant == "channels3":
        desc = channelwise_tv(40, 132, 3)
    elif variant == "warm_start":
        r = np.random.default_rng(0)
        x0 = r.random(desc["ncols"]).astype(np.float32)
        y0 = (0.3 * r.standard_normal(desc["nrows"])).astype(np.float32)
    elif variant == "residual_iter_1":
        opts["residual_iter"] = 1          # every iteration refreshes: only the CHECK variant of the tiled kernel runs
    elif variant == "residual_never":
        opts["residual_iter"] = -1         # only iteration 0 refreshes (size_t % int wrap)
    n_tile = tiled_iterations(ctx, desc, **opts)
    assert n_tile > 0
    a = run_cuda(ctx, desc, 60, fuse=1, x0=x0, y0=y0, **opts)
    b = run_cuda(ctx, desc, 60, fuse=3, x0=x0, y0=y0, **opts)
    for k in ("x", "y", "z", "w"):
        assert np.array_equal(a[k], b[k]), (variant, k)
    want = run_oracle(desc, 60, x0=x0, y0=y0, **opts)
    assert_parity(a, want, label=variant)


def test_tile_not_selected_when_structure_differs(ctx):
    assert tiled_iterations(ctx, syn.tvl1(32, 128, nc=3), stepsize="alg1", residual_iter=4) == 0   # per-pixel groups
    assert tiled_iterations(ctx, syn.tv3d(8, 16, 4), stepsize="alg1", residual_iter=4) == 0        # 3-D gradient
    assert tiled_iterations(ctx, syn.rof(32, 126), stepsize="alg1", residual_iter=4) == 0           # ny % 4 != 0


def test_tile_metric_config_row_checksums(ctx):
    """ROF 2048 x 4096 (half the metric config's columns, same column length): tiled vs two-pass bit-identical,
    compared through a checksum of checksums so that the test stays cheap on the host."""
    desc = syn.rof(2048, 4096)
    opts = dict(stepsize="alg1", residual_iter=10)
    a = run_cuda(ctx, desc, 45, fuse=1, **opts)
    b = run_cuda(ctx, desc, 45, fuse=3, **opts)
    for k in ("x", "y"):
        va, vb = a[k].view(np.uint32).astype(np.uint64), b[k].view(np.uint32).astype(np.uint64)
        assert int(va.sum()) == int(vb.sum()) and int((va * np.arange(1, va.size + 1, dtype=np.uint64)).sum()) == \
            int((vb * np.arange(1, vb.size + 1, dtype=np.uint64)).sum()), k
    assert same_residuals(a, b)
