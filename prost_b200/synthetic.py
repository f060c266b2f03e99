"""Synthetic inputs and problem descriptions for the BASELINE.json configurations.

Inputs come from a counter-based hash RNG so that any slab of any array
can be generated independently and bit-identically on every rank:

    u(stream, i) = (splitmix64(42 + stream * 2^40 + i) >> 40) / 2^24   in [0, 1)

Problem descriptions use the reference's mex registry vocabulary (see prost_b200/factory.py) and
the parameters of the reference's examples (matlab/examples/example_rof_primaldual.m,
example_tvl1.m, example_multilabel_fast.m).
"""
import numpy as np

SEED = 42
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(stream, idx):
    """u(stream, i) for an integer array of counters ``idx``; float32 in [0, 1)."""
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = np.uint64(SEED) + (np.uint64(stream) << np.uint64(40)) + idx
    return ((splitmix64(key) >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)


def normal(stream, idx):
    """N(0,1) by Box-Muller on streams (stream, stream + 1)."""
    u1 = (uniform(stream, idx).astype(np.float64) * (1 << 24) + 0.5) / float(1 << 24)
    u2 = uniform(stream + 1, idx).astype(np.float64)
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)


def image(nx, ny, nc=1, sigma_n=0.1, stream=1, x0=0, x1=None):
    """Noisy test image, column-major ``y + x*ny + c*nx*ny`` (block_gradient2d.cu:59), columns
    [x0, x1) only.  Returned flat, length (x1-x0)*ny*nc, planar by channel."""
    x1 = nx if x1 is None else x1
    xs = np.arange(x0, x1, dtype=np.float64)[:, None]          # slow axis
    ys = np.arange(ny, dtype=np.float64)[None, :]
    out = np.empty((nc, x1 - x0, ny), dtype=np.float32)
    lin = (np.arange(x0, x1, dtype=np.uint64)[:, None] * np.uint64(ny) + np.arange(ny, dtype=np.uint64)[None, :])
    for c in range(nc):
        base = 0.5 + 0.25 * np.sin(2 * np.pi * 3 * xs / nx) * np.cos(2 * np.pi * 2 * ys / ny) \
            + 0.1 * (c + 1) / nc * np.sin(2 * np.pi * 5 * (xs + ys) / (nx + ny))
        g = normal(stream + 2 * c, lin + np.uint64(c) * np.uint64(nx * ny))
        out[c] = np.clip(base + sigma_n * g, 0.0, 1.0).astype(np.float32)
    return out.reshape(-1)


def salt_and_pepper(f, stream=20, frac=0.25, nx=None, ny=None, x0=0, x1=None):
    """25 % salt-and-pepper noise as in example_tvl1.m:11-14.  With ``nx, ny, x0, x1`` the array is the column
    slab [x0, x1) of a planar nx x ny x nc image and the hash counters are the GLOBAL element indices, so a slab
    equals the slice of the whole noisy image."""
    if nx is None:
        idx = np.arange(f.size, dtype=np.uint64)
    else:
        x1 = nx if x1 is None else x1
        w = x1 - x0
        nc = f.size // (w * ny)
        lin = (np.arange(x0, x1, dtype=np.uint64)[:, None] * np.uint64(ny) + np.arange(ny, dtype=np.uint64)[None, :])
        idx = (np.arange(nc, dtype=np.uint64)[:, None, None] * np.uint64(nx * ny) + lin[None]).reshape(-1)
    u = uniform(stream, idx)
    out = f.copy()
    out[u < frac / 2] = 1.0
    out[(u >= frac / 2) & (u < frac)] = 0.0
    return out


def _coeffs(a=1, b=0, c=1, d=0, e=0, alpha=0, beta=0):
    return [np.atleast_1d(np.asarray(v, dtype=np.float32)) for v in (a, b, c, d, e, alpha, beta)]


# ---------------------------------------------------------------------------------------------
# Problem descriptions
# ---------------------------------------------------------------------------------------------
def rof(nx, ny, lam=10.0, f=None):
    """C1 / metric config: ROF denoising  min_u (lam/2)|u-f|^2 + |grad u|_{2,1}
    (example_rof_primaldual.m:11-26)."""
    N = nx * ny
    f = image(nx, ny) if f is None else f
    return dict(
        nrows=2 * N, ncols=N,
        blocks=[("gradient2d", 0, 0, [nx, ny, 1, False])],
        prox_g=[("elem_operation:1d:square", 0, N, True, [N, 1, False, _coeffs(a=1, b=f, c=lam)])],
        prox_fstar=[("elem_operation:norm2:ind_leq0", 0, 2 * N, False, [N, 2, False, _coeffs(a=1, b=1, c=1)])],
        scaling=("alpha", 1.0), data=dict(f=f))


def tvl1(nx, ny, nc=3, lam=1.0, f=None):
    """C2: TV-L1 colour denoising with diagonal preconditioning (example_tvl1.m)."""
    N = nx * ny
    if f is None:
        f = salt_and_pepper(image(nx, ny, nc))
    return dict(
        nrows=2 * N * nc, ncols=N * nc,
        blocks=[("gradient2d", 0, 0, [nx, ny, nc, False])],
        prox_g=[("elem_operation:1d:abs", 0, N * nc, True, [N * nc, 1, False, _coeffs(a=1, b=f, c=lam)])],
        prox_fstar=[("elem_operation:norm2:ind_leq0", 0, 2 * N * nc, False,
                     [N, 2 * nc, False, _coeffs(a=1, b=1, c=1)])],
        scaling=("alpha", 1.0), data=dict(f=f))


def tv3d(nx, ny, L, lam=10.0, f=None):
    """C4: 3-D TV denoising with BlockGradient3D (Dirichlet in the third direction)."""
    N = nx * ny * L
    f = image(nx, ny, L) if f is None else f
    return dict(
        nrows=3 * N, ncols=N,
        blocks=[("gradient3d", 0, 0, [nx, ny, L, False])],
        prox_g=[("elem_operation:1d:square", 0, N, True, [N, 1, False, _coeffs(a=1, b=f, c=lam)])],
        prox_fstar=[("elem_operation:norm2:ind_leq0", 0, 3 * N, False, [N, 3, False, _coeffs(a=1, b=1, c=1)])],
        scaling=("alpha", 1.0), data=dict(f=f))


def lifting(nx, ny, L, lam=1.0, x0=0, x1=None):
    """C3: lifted multilabel stand-in built from in-tree operators only:
    K = [grad2d over L label planes ; identity], g = simplex over labels, f* = norm-ball on the
    gradient rows + epigraph of a quadratic on the identity rows.

    ``x0, x1``: only the column slab [x0, x1) of the nx-column problem (what one rank of the slab
    decomposition owns); every coefficient is a function of the GLOBAL pixel index, so the slab equals
    ``distributed.shard_description`` of the global description without ever building that."""
    x1 = nx if x1 is None else x1
    w = x1 - x0
    N, Nw = nx * ny, w * ny
    NL = Nw * L
    f = image(nx, ny, x0=x0, x1=x1)
    lab = (np.arange(L, dtype=np.float32) / L)[:, None]
    rho = ((lab - f[None, :]) ** 2).astype(np.float32).reshape(-1)          # unary cost, planar by label
    # pairs (x_i, y_i) = (r[i], r[NL/2 + i]): i runs over the first NL/2 entries of the label-planar vector,
    # i.e. label planes 0 .. L/2-1 (and half of plane (L-1)/2 when L is odd); b_i is a hash of the GLOBAL index
    pix = (np.arange(x0, x1, dtype=np.uint64)[:, None] * np.uint64(ny) + np.arange(ny, dtype=np.uint64)[None, :]).reshape(-1)
    j = np.arange(NL // 2, dtype=np.uint64)
    gidx = (j // np.uint64(Nw)) * np.uint64(N) + pix[(j % np.uint64(Nw)).astype(np.int64)]
    b = (2.0 * uniform(30, gidx) - 1.0).astype(np.float32)
    c = -rho[: NL // 2]
    return dict(
        nrows=3 * NL, ncols=NL,
        blocks=[("gradient2d", 0, 0, [w, ny, L, False]),
                ("diags", 2 * NL, 0, [NL, NL, [1.0], [0]])],
        prox_g=[("elem_operation:ind_simplex", 0, NL, False, [Nw, L, False])],
        prox_fstar=[("elem_operation:norm2:ind_leq0", 0, 2 * NL, False,
                     [Nw, 2 * L, False, _coeffs(a=1.0 / lam, b=1, c=1)]),
                    ("ind_epi_quad", 2 * NL, NL, False, [NL // 2, 2, False, [[1.0], b, c]])],
        scaling=("alpha", 1.0), data=dict(f=f))


def lasso(m, n, nnz_per_row=12, dense=0, lam=0.1):
    """C5: ADMM LASSO  min_x lam|x|_1 + (1/2)|Kx - b|^2 with a random sparse K (exactly
    ``nnz_per_row`` entries per row at hashed columns, N(0,1)/sqrt(nnz_per_row)) and an optional
    ``dense`` x ``dense`` BlockDense appended below it."""
    import scipy.sparse as sp
    rows = np.repeat(np.arange(m, dtype=np.int64), nnz_per_row)
    k = np.arange(m * nnz_per_row, dtype=np.uint64)
    cols = (splitmix64(np.uint64(SEED) + (np.uint64(40) << np.uint64(40)) + k) % np.uint64(n)).astype(np.int64)
    vals = normal(41, k) / np.sqrt(nnz_per_row)
    K = sp.csr_matrix((vals, (rows, cols)), shape=(m, n), dtype=np.float32)   # duplicates are summed
    K.sum_duplicates()
    xs = np.where(uniform(43, np.arange(n, dtype=np.uint64)) < 0.01,
                  normal(44, np.arange(n, dtype=np.uint64)), 0).astype(np.float32)
    blocks = [("sparse", 0, 0, [K.tocsc()])]
    rows_total = m
    bvec = K @ xs
    if dense:
        D = (normal(46, np.arange(dense * dense, dtype=np.uint64)) / 64.0).reshape(dense, dense)
        blocks.append(("dense", m, 0, [D]))
        bvec = np.concatenate([bvec, D @ xs[:dense]])
        rows_total += dense
    bvec = (bvec + 0.01 * normal(48, np.arange(rows_total, dtype=np.uint64))).astype(np.float32)
    return dict(
        nrows=rows_total, ncols=n, blocks=blocks,
        prox_g=[("elem_operation:1d:abs", 0, n, True, [n, 1, False, _coeffs(a=1, b=0, c=lam)])],
        prox_f=[("elem_operation:1d:square", 0, rows_total, True,
                 [rows_total, 1, False, _coeffs(a=1, b=bvec, c=1)])],
        scaling=("identity",), data=dict(b=bvec, x_true=xs))
