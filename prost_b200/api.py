"""Host-side mirror of the reference's operator/plugin interface over the C ABI.

Class and method names follow the reference's C++ API (include/prost/*.hpp) so that the parity
tests read like the reference's own: ``Problem.AddBlock / AddProx_g / SetScalingAlpha /
Initialize``, ``Solver.Initialize / Solve``, ``LinearOperator.Eval / EvalAdjoint``, ``Prox.Eval``.
This module is plumbing only: all arithmetic happens in ``libprost_b200.so`` on the GPU.
"""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import ADMMOptions, PDHGOptions, ProstError, SolverOptions, check, lib

FUNCTIONS_1D = ["zero", "abs", "square", "ind_leq0", "ind_geq0", "ind_eq0", "ind_box01", "max_pos0",
                "l0", "huber", "lq", "lq_plus_eps", "truncquad", "trunclin"]


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).ravel())


def _fp(a):
    return a.ctypes.data_as(_capi.c_float_p)


class _PinnedBlock:
    """One block of the library's pinned-host pool (pb_host_alloc); goes back to the pool when collected."""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(lib.pb_host_alloc(int(nbytes), C.byref(self.ptr)))

    def __del__(self):
        if getattr(self, "ptr", None) and self.ptr.value and lib is not None:
            lib.pb_host_free(self.ptr)
            self.ptr = C.c_void_p()


def pinned_empty(n, dtype=np.float32):
    """float32 vector of n elements in pinned host memory; the array keeps its block alive."""
    dt = np.dtype(dtype)
    blk = _PinnedBlock(max(int(n), 1) * dt.itemsize)
    raw = (C.c_char * (max(int(n), 1) * dt.itemsize)).from_address(blk.ptr.value)
    raw._owner = blk
    return np.frombuffer(raw, dtype=dt, count=int(n))


def function_id(name):
    fid = lib.pb_function1d_from_name(name.encode())
    if fid < 0:
        raise ProstError(-1, f"unknown Function1D '{name}'")
    return fid


class Context:
    """One GPU + stream (pb_context).  ``stream`` may be a raw cudaStream_t (int), e.g.
    ``torch.cuda.current_stream().cuda_stream``."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        check(lib.pb_context_create(int(device), C.c_void_p(stream or 0), C.byref(self._h)))
        self.device = device

    def synchronize(self):
        check(lib.pb_context_synchronize(self._h))

    @property
    def stream(self):
        return lib.pb_context_stream(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib.pb_context_destroy(self._h)
            self._h = None


class Comm:
    """pb_comm: one rank of the slab decomposition (one process per GPU).  ``unique_id`` is the 128
    bytes rank 0 got from :func:`Comm.unique_id`, distributed by the caller (see
    prost_b200.distributed.init_comm for the torch.distributed way)."""

    ID_BYTES = 128

    def __init__(self, ctx, rank, world, unique_id):
        assert len(unique_id) == self.ID_BYTES
        self.ctx = ctx
        self._h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), self.ID_BYTES)
        check(lib.pb_comm_create(ctx._h, int(rank), int(world), buf, C.byref(self._h)))

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(Comm.ID_BYTES)
        check(lib.pb_comm_unique_id(buf))
        return buf.raw

    rank = property(lambda s: lib.pb_comm_rank(s._h))
    world = property(lambda s: lib.pb_comm_world(s._h))
    peer_to_peer = property(lambda s: bool(lib.pb_comm_peer_to_peer(s._h)))

    def barrier(self):
        check(lib.pb_comm_barrier(self._h))

    def allreduce_sum(self, values):
        a = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel())
        check(lib.pb_comm_allreduce_sum(self._h, a.ctypes.data_as(_capi.c_double_p), a.size))
        return a

    def close(self):
        if getattr(self, "_h", None):
            lib.pb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class _Handle:
    _destroy = None

    def __init__(self, ctx):
        self.ctx = ctx
        self._h = C.c_void_p()

    def __del__(self):
        if getattr(self, "_h", None) and self._destroy:
            getattr(lib, self._destroy)(self._h)
            self._h = None


# ---------------------------------------------------------------------------------------------
# Blocks (include/prost/linop/*.hpp)
# ---------------------------------------------------------------------------------------------
class Block(_Handle):
    _destroy = "pb_block_destroy"

    row = property(lambda s: lib.pb_block_row(s._h))
    col = property(lambda s: lib.pb_block_col(s._h))
    nrows = property(lambda s: lib.pb_block_nrows(s._h))
    ncols = property(lambda s: lib.pb_block_ncols(s._h))

    def row_sum(self, row, alpha):
        return lib.pb_block_row_sum(self._h, row, alpha)

    def col_sum(self, col, alpha):
        return lib.pb_block_col_sum(self._h, col, alpha)


class BlockGradient2D(Block):
    def __init__(self, ctx, row, col, nx, ny, L, label_first=False):
        super().__init__(ctx)
        check(lib.pb_block_create_gradient2d(ctx._h, row, col, nx, ny, L, int(label_first), C.byref(self._h)))


class BlockGradient3D(Block):
    def __init__(self, ctx, row, col, nx, ny, L, label_first=False):
        super().__init__(ctx)
        check(lib.pb_block_create_gradient3d(ctx._h, row, col, nx, ny, L, int(label_first), C.byref(self._h)))


class BlockDiags(Block):
    def __init__(self, ctx, row, col, nrows, ncols, offsets, factors):
        super().__init__(ctx)
        ofs = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64).ravel())
        fac = _f32(factors)
        assert ofs.size == fac.size
        check(lib.pb_block_create_diags(ctx._h, row, col, nrows, ncols, ofs.size,
                                        ofs.ctypes.data_as(_capi.c_i64_p), _fp(fac), C.byref(self._h)))


class BlockSparse(Block):
    """BlockSparse::CreateFromCSC; ``A`` is anything scipy.sparse can turn into CSC."""

    def __init__(self, ctx, row, col, A):
        import scipy.sparse as sp
        super().__init__(ctx)
        A = sp.csc_matrix(A)
        A.sort_indices()
        val = _f32(A.data)
        ptr = np.ascontiguousarray(A.indptr.astype(np.int32))
        ind = np.ascontiguousarray(A.indices.astype(np.int32))
        check(lib.pb_block_create_sparse_csc(ctx._h, row, col, A.shape[0], A.shape[1], A.nnz, _fp(val),
                                             ptr.ctypes.data_as(_capi.c_i32_p),
                                             ind.ctypes.data_as(_capi.c_i32_p), C.byref(self._h)))


class BlockDense(Block):
    """BlockDense::CreateFromColFirstData; ``A`` is a 2-D array (stored column-major)."""

    def __init__(self, ctx, row, col, A):
        super().__init__(ctx)
        A = np.asarray(A, dtype=np.float32)
        data = np.ascontiguousarray(A.T).ravel()          # column-major
        check(lib.pb_block_create_dense(ctx._h, row, col, A.shape[0], A.shape[1], _fp(data), C.byref(self._h)))


class BlockDenseKronId(Block):
    """BlockDenseKronId::CreateFromColFirstData: kron(K, I_diaglength); ``K`` is a small 2-D array."""

    def __init__(self, ctx, row, col, K, diaglength):
        super().__init__(ctx)
        K = np.asarray(K, dtype=np.float32)
        data = np.ascontiguousarray(K.T).ravel()          # column-major
        check(lib.pb_block_create_dense_kron_id(ctx._h, int(diaglength), row, col, K.shape[0], K.shape[1], _fp(data),
                                                C.byref(self._h)))


class BlockIdKronDense(Block):
    """BlockIdKronDense::CreateFromColFirstData: kron(I_diaglength, K)."""

    def __init__(self, ctx, row, col, K, diaglength):
        super().__init__(ctx)
        K = np.asarray(K, dtype=np.float32)
        data = np.ascontiguousarray(K.T).ravel()
        check(lib.pb_block_create_id_kron_dense(ctx._h, int(diaglength), row, col, K.shape[0], K.shape[1], _fp(data),
                                                C.byref(self._h)))


class _BlockSparseKron(Block):
    _create = None

    def __init__(self, ctx, row, col, K, diaglength):
        import scipy.sparse as sp
        super().__init__(ctx)
        K = sp.csc_matrix(K)
        K.sort_indices()
        val = _f32(K.data)
        ptr = np.ascontiguousarray(K.indptr.astype(np.int32))
        ind = np.ascontiguousarray(K.indices.astype(np.int32))
        check(getattr(lib, self._create)(ctx._h, row, col, int(diaglength), K.shape[0], K.shape[1], K.nnz, _fp(val),
                                         ptr.ctypes.data_as(_capi.c_i32_p), ind.ctypes.data_as(_capi.c_i32_p),
                                         C.byref(self._h)))


class BlockSparseKronId(_BlockSparseKron):
    """BlockSparseKronId::CreateFromCSC: kron(K, I_diaglength) for a sparse factor."""
    _create = "pb_block_create_sparse_kron_id"


class BlockIdKronSparse(_BlockSparseKron):
    """BlockIdKronSparse::CreateFromCSC: kron(I_diaglength, K) for a sparse factor."""
    _create = "pb_block_create_id_kron_sparse"


class BlockZero(Block):
    def __init__(self, ctx, row, col, nrows, ncols):
        super().__init__(ctx)
        check(lib.pb_block_create_zero(ctx._h, row, col, nrows, ncols, C.byref(self._h)))


class LinearOperator(_Handle):
    _destroy = "pb_linop_destroy"

    def __init__(self, ctx):
        super().__init__(ctx)
        check(lib.pb_linop_create(ctx._h, C.byref(self._h)))
        self._blocks = []

    def AddBlock(self, block):
        check(lib.pb_linop_add_block(self._h, block._h))
        self._blocks.append(block)

    def Initialize(self):
        check(lib.pb_linop_initialize(self._h))

    nrows = property(lambda s: lib.pb_linop_nrows(s._h))
    ncols = property(lambda s: lib.pb_linop_ncols(s._h))

    def _eval_host(self, rhs, transpose):
        rhs = _f32(rhs)
        nin = self.nrows if transpose else self.ncols
        if rhs.size != nin:
            raise ProstError(-1, f"rhs has {rhs.size} elements, operator expects {nin}")
        out = np.empty(self.ncols if transpose else self.nrows, dtype=np.float32)
        ms = C.c_double()
        check(lib.pb_linop_eval_host(self._h, _fp(out), _fp(rhs), int(transpose), C.byref(ms)))
        self.last_ms = ms.value
        return out

    def Eval(self, rhs):
        """LinearOperator::Eval(std::vector&, const std::vector&) (linearoperator.cu:172-194)."""
        return self._eval_host(rhs, False)

    def EvalAdjoint(self, rhs):
        return self._eval_host(rhs, True)

    def row_sums(self, alpha):
        out = np.empty(self.nrows, dtype=np.float32)
        check(lib.pb_linop_row_sums(self._h, alpha, _fp(out)))
        return out

    def col_sums(self, alpha):
        out = np.empty(self.ncols, dtype=np.float32)
        check(lib.pb_linop_col_sums(self._h, alpha, _fp(out)))
        return out


# ---------------------------------------------------------------------------------------------
# Proxes (include/prost/prox/*.hpp)
# ---------------------------------------------------------------------------------------------
class Prox(_Handle):
    _destroy = "pb_prox_destroy"

    index = property(lambda s: lib.pb_prox_index(s._h))
    size = property(lambda s: lib.pb_prox_size(s._h))
    diagsteps = property(lambda s: bool(lib.pb_prox_diagsteps(s._h)))

    def Eval(self, arg, tau_diag, tau, invert_tau=False):
        """Prox::Eval(std::vector& result, arg, tau_diag, tau) (prox.cu:45-71)."""
        arg = _f32(arg)
        td = _f32(tau_diag)
        assert arg.size == td.size
        res = np.empty_like(arg)
        ms = C.c_double()
        check(lib.pb_prox_eval_host(self._h, _fp(res), _fp(arg), _fp(td), arg.size, float(tau),
                                    int(invert_tau), C.byref(ms)))
        self.last_ms = ms.value
        return res


def _coeff_arrays(coeffs):
    arrs = [_f32(c) for c in coeffs]
    assert len(arrs) == 7, "a, b, c, d, e, alpha, beta"
    ptrs = (_capi.c_float_p * 7)(*[_fp(a) for a in arrs])
    lens = (C.c_size_t * 7)(*[a.size for a in arrs])
    return arrs, ptrs, lens


class ProxElemOperation1D(Prox):
    """ProxElemOperation<T, ElemOperation1D<T, Function1D*>>."""

    def __init__(self, ctx, function, index, count, dim, interleaved, diagsteps, coeffs):
        super().__init__(ctx)
        keep, ptrs, lens = _coeff_arrays(coeffs)
        check(lib.pb_prox_create_elem_1d(ctx._h, index, count, dim, int(interleaved), int(diagsteps),
                                         function_id(function), ptrs, lens, C.byref(self._h)))


class ProxElemOperationNorm2(Prox):
    """ProxElemOperation<T, ElemOperationNorm2<T, Function1D*>>."""

    def __init__(self, ctx, function, index, count, dim, interleaved, diagsteps, coeffs):
        super().__init__(ctx)
        keep, ptrs, lens = _coeff_arrays(coeffs)
        check(lib.pb_prox_create_elem_norm2(ctx._h, index, count, dim, int(interleaved), int(diagsteps),
                                            function_id(function), ptrs, lens, C.byref(self._h)))


class ProxElemOperationIndSimplex(Prox):
    def __init__(self, ctx, index, count, dim, interleaved, diagsteps):
        super().__init__(ctx)
        check(lib.pb_prox_create_ind_simplex(ctx._h, index, count, dim, int(interleaved), int(diagsteps),
                                             C.byref(self._h)))


class ProxIndEpiQuad(Prox):
    def __init__(self, ctx, index, count, dim, interleaved, diagsteps, a, b, c):
        super().__init__(ctx)
        a, b, c = _f32(a), _f32(b), _f32(c)
        check(lib.pb_prox_create_ind_epi_quad(ctx._h, index, count, dim, int(interleaved), int(diagsteps),
                                              _fp(a), a.size, _fp(b), b.size, _fp(c), c.size,
                                              C.byref(self._h)))


class ProxElemOperationIndSum(Prox):
    """ProxElemOperation<T, ElemOperationIndSum<T>>: per-group projection onto sum_i x_i = 1."""

    def __init__(self, ctx, index, count, dim, interleaved, diagsteps):
        super().__init__(ctx)
        check(lib.pb_prox_create_ind_sum(ctx._h, index, count, dim, int(interleaved), int(diagsteps),
                                         C.byref(self._h)))


class ProxIndSum(Prox):
    """ProxIndSum<T>(index, size, count, dim, inds, sum[, count2, dim2, inds2, sum2]) (prox_ind_sum.hpp:37-62):
    index-list groups projected onto sum = ``sum`` in the metric of the step sizes; count = len(inds) / dim like in
    the mex factory (factory.cpp:459-481)."""

    def __init__(self, ctx, index, size, dim, inds, total, dim2=None, inds2=None, total2=0.0):
        super().__init__(ctx)
        u64p = C.POINTER(C.c_ulonglong)
        a = np.ascontiguousarray(np.asarray(inds, dtype=np.uint64).ravel())
        two = inds2 is not None
        b = np.ascontiguousarray(np.asarray(inds2 if two else [], dtype=np.uint64).ravel())
        d2 = int(dim2) if two else 0
        check(lib.pb_prox_create_ind_sum_indexed(
            ctx._h, index, size, a.size // int(dim), int(dim), a.ctypes.data_as(u64p), float(total),
            (b.size // d2) if two and d2 else 0, d2, b.ctypes.data_as(u64p) if two else None, float(total2),
            C.byref(self._h)))


class ProxElemOperationSpectral(Prox):
    """ProxElemOperation<T, ElemOperationSingularNx2 / Eigen2x2 / Eigen3x3 / EigenNxN> (elem_operation_singular_nx2.hpp,
    elem_operation_eigen_*.hpp).  ``kind``: "singular_nx2", "eigen_2x2", "eigen_3x3", "eigen_nxn"; ``function`` is
    a Function1D name, or for singular_nx2 also "ind_l1_ball" / "moreau:ind_l1_ball" (function_2d.hpp)."""
    KINDS = {"singular_nx2": 0, "eigen_2x2": 1, "eigen_3x3": 2, "eigen_nxn": 3, "mass4": 4, "ind_comass4_ball": 5,
             "mass5": 6, "ind_comass5_ball": 7}

    def __init__(self, ctx, kind, function, index, count, dim, interleaved, diagsteps, coeffs):
        super().__init__(ctx)
        if self.KINDS[kind] >= 4:       # mass / comass norms: one optional coefficient (the cost of mass4)
            cost = coeffs[0] if coeffs else [1.0]
            coeffs, function = [cost, [0.0], [1.0], [0.0], [0.0], [0.0], [0.0]], "zero"
        keep, ptrs, lens = _coeff_arrays(coeffs)
        fn2d = {"ind_l1_ball": 1, "moreau:ind_l1_ball": 2}.get(function, 0)
        fn1d = 0 if fn2d else function_id(function[len("sum_1d:"):] if function.startswith("sum_1d:") else function)
        check(lib.pb_prox_create_spectral(ctx._h, self.KINDS[kind], index, count, dim, int(interleaved), int(diagsteps),
                                          fn1d, fn2d, ptrs, lens, C.byref(self._h)))


class ProxIndRange(Prox):
    """ProxIndRange<T>(index, size, diagsteps) with setA / setAA (prox_ind_range.hpp:37-50): projection onto the range
    of the sparse matrix ``A``; ``AA`` = A^T A (dense) is computed here when not given."""

    def __init__(self, ctx, index, size, diagsteps, A, AA=None):
        import scipy.sparse as sp
        super().__init__(ctx)
        A = sp.csc_matrix(A)
        A.sort_indices()
        if AA is None:
            AA = (A.T @ A).toarray()
        val = _f32(A.data)
        ptr = np.ascontiguousarray(A.indptr.astype(np.int32))
        ind = np.ascontiguousarray(A.indices.astype(np.int32))
        aa = np.ascontiguousarray(np.asarray(AA, dtype=np.float32).T).ravel()       # column-major
        check(lib.pb_prox_create_ind_range(ctx._h, index, size, int(diagsteps), A.shape[0], A.shape[1], A.nnz, _fp(val),
                                           ptr.ctypes.data_as(_capi.c_i32_p), ind.ctypes.data_as(_capi.c_i32_p),
                                           _fp(aa), C.byref(self._h)))


class ProxIndEpiConjQuad1D(Prox):
    """ProxIndEpiConjQuad1D (the north star's "ProxEpiConjQuadr"; source external to the reference tree, parity
    unpinned): per (x, y) pair the projection onto the epigraph of the conjugate of a u^2 + b u + c on
    [alpha, beta].  Every coefficient is a scalar or one value per pair."""

    def __init__(self, ctx, index, count, interleaved, diagsteps, a, b, c, alpha, beta):
        super().__init__(ctx)
        keep = [_f32(np.atleast_1d(v)) for v in (a, b, c, alpha, beta)]
        ptrs = (_capi.c_float_p * 5)(*[_fp(v) for v in keep])
        lens = (C.c_size_t * 5)(*[v.size for v in keep])
        check(lib.pb_prox_create_ind_epi_conjquad_1d(ctx._h, index, count, int(interleaved), int(diagsteps), ptrs, lens,
                                                     C.byref(self._h)))


class ProxIndHalfspace(Prox):
    """ProxIndHalfspace<T>(index, count, dim, interleaved, diagsteps, a, b): projection onto <a, x> <= b."""

    def __init__(self, ctx, index, count, dim, interleaved, diagsteps, a, b):
        super().__init__(ctx)
        a, b = _f32(a), _f32(b)
        check(lib.pb_prox_create_ind_halfspace(ctx._h, index, count, dim, int(interleaved), int(diagsteps),
                                               _fp(a), a.size, _fp(b), b.size, C.byref(self._h)))


class ProxIndSOC(Prox):
    """ProxIndSOC<T>(index, count, dim, interleaved, diagsteps, alpha): projection onto the second-order cone."""

    def __init__(self, ctx, index, count, dim, interleaved, diagsteps, alpha=1.0):
        super().__init__(ctx)
        check(lib.pb_prox_create_ind_soc(ctx._h, index, count, dim, int(interleaved), int(diagsteps), float(alpha),
                                         C.byref(self._h)))


class ProxTransform(Prox):
    """ProxTransform<T>(inner, a, b, c, d, e): prox of c f(a x - b) + <d, x> + (e/2)|x|^2 through the prox of f
    (prox_transform.hpp:38-44); every coefficient is a scalar or one value per element."""

    def __init__(self, ctx, inner, a=1.0, b=0.0, c=1.0, d=0.0, e=0.0):
        super().__init__(ctx)
        self._inner = inner
        keep = [_f32(np.atleast_1d(v)) for v in (a, b, c, d, e)]
        ptrs = (_capi.c_float_p * 5)(*[_fp(v) for v in keep])
        lens = (C.c_size_t * 5)(*[v.size for v in keep])
        check(lib.pb_prox_create_transform(ctx._h, inner._h, ptrs, lens, C.byref(self._h)))


class ProxMoreau(Prox):
    def __init__(self, ctx, conjugate):
        super().__init__(ctx)
        self._inner = conjugate
        check(lib.pb_prox_create_moreau(ctx._h, conjugate._h, C.byref(self._h)))


class ProxPermute(Prox):
    def __init__(self, ctx, base, perm):
        super().__init__(ctx)
        self._inner = base
        perm = np.ascontiguousarray(np.asarray(perm, dtype=np.int32).ravel())
        check(lib.pb_prox_create_permute(ctx._h, base._h, perm.ctypes.data_as(_capi.c_int_p), perm.size,
                                         C.byref(self._h)))


class ProxZero(Prox):
    def __init__(self, ctx, index, size):
        super().__init__(ctx)
        check(lib.pb_prox_create_zero(ctx._h, index, size, C.byref(self._h)))


# ---------------------------------------------------------------------------------------------
# Problem / Backend / Solver
# ---------------------------------------------------------------------------------------------
class Problem(_Handle):
    """prost::Problem (include/prost/problem.hpp:64-114)."""
    _destroy = "pb_problem_destroy"

    def __init__(self, ctx):
        super().__init__(ctx)
        check(lib.pb_problem_create(ctx._h, C.byref(self._h)))
        self._keep = []

    def AddBlock(self, block):
        check(lib.pb_problem_add_block(self._h, block._h))
        self._keep.append(block)

    def AddProx_g(self, prox):
        check(lib.pb_problem_add_prox_g(self._h, prox._h))
        self._keep.append(prox)

    def AddProx_f(self, prox):
        check(lib.pb_problem_add_prox_f(self._h, prox._h))
        self._keep.append(prox)

    def AddProx_gstar(self, prox):
        check(lib.pb_problem_add_prox_gstar(self._h, prox._h))
        self._keep.append(prox)

    def AddProx_fstar(self, prox):
        check(lib.pb_problem_add_prox_fstar(self._h, prox._h))
        self._keep.append(prox)

    def SetDimensions(self, nrows, ncols):
        check(lib.pb_problem_set_dimensions(self._h, nrows, ncols))

    def SetScalingAlpha(self, alpha):
        check(lib.pb_problem_set_scaling_alpha(self._h, alpha))

    def SetScalingIdentity(self):
        check(lib.pb_problem_set_scaling_identity(self._h))

    def SetScalingCustom(self, left, right):
        left, right = _f32(left), _f32(right)
        check(lib.pb_problem_set_scaling_custom(self._h, _fp(left), left.size, _fp(right), right.size))

    def Initialize(self):
        check(lib.pb_problem_initialize(self._h))

    def Dualize(self):
        check(lib.pb_problem_dualize(self._h))

    nrows = property(lambda s: lib.pb_problem_nrows(s._h))
    ncols = property(lambda s: lib.pb_problem_ncols(s._h))

    def normest(self, tol=1e-6, max_iters=100, x0=None):
        out = C.c_float()
        x0a = _f32(x0) if x0 is not None else None
        check(lib.pb_problem_normest(self._h, tol, max_iters, _fp(x0a) if x0a is not None else None,
                                     C.byref(out)))
        return out.value

    def scaling(self):
        left = np.empty(self.nrows, dtype=np.float32)
        right = np.empty(self.ncols, dtype=np.float32)
        check(lib.pb_problem_get_scaling(self._h, _fp(left), _fp(right)))
        return left, right


def solver_options(**kw):
    o = SolverOptions()
    lib.pb_solver_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


_STEPS = {"alg1": 1, "alg2": 2, "goldstein": 3, "boyd": 4}


def pdhg_options(**kw):
    o = PDHGOptions()
    lib.pb_pdhg_default_options(C.byref(o))
    for k, v in kw.items():
        if k == "stepsize":
            o.stepsize_variant = _STEPS[v]
            continue
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def admm_options(**kw):
    o = ADMMOptions()
    lib.pb_admm_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class Backend(_Handle):
    """prost::Backend (include/prost/backend/backend.hpp:37-95)."""
    _destroy = "pb_backend_destroy"

    def Initialize(self, x0=None, y0=None):
        x0a = _f32(x0) if x0 is not None else None
        y0a = _f32(y0) if y0 is not None else None
        check(lib.pb_backend_initialize(self._h, _fp(x0a) if x0a is not None else None,
                                        x0a.size if x0a is not None else 0,
                                        _fp(y0a) if y0a is not None else None,
                                        y0a.size if y0a is not None else 0))

    def PerformIteration(self, n=1):
        check(lib.pb_backend_iterate(self._h, n))

    def profile(self, n=1):
        """Average device ms per iteration of (primal pass, dual pass, finalize); advances n iterations."""
        out = (C.c_float * 3)()
        check(lib.pb_backend_profile(self._h, n, out))
        return float(out[0]), float(out[1]), float(out[2])

    def profile_detail(self, n=1):
        """dict(primal_ms, dual_ms, finalize_ms, tile_ms, n_two_pass, n_tile, tile_check_ms, n_tile_check):
        see pb_backend_profile_detail."""
        out = (C.c_float * 8)()
        check(lib.pb_backend_profile_detail(self._h, n, out))
        keys = ["primal_ms", "dual_ms", "finalize_ms", "tile_ms", "n_two_pass", "n_tile", "tile_check_ms",
                "n_tile_check"]
        return dict(zip(keys, [float(v) for v in out]))

    def residuals(self):
        out = (C.c_float * 6)()
        check(lib.pb_backend_residuals(self._h, out))
        keys = ["primal_residual", "dual_residual", "primal_var_norm", "dual_var_norm", "eps_primal", "eps_dual"]
        return dict(zip(keys, [float(v) for v in out]))

    def stepsizes(self):
        out = (C.c_double * 3)()
        check(lib.pb_backend_stepsizes(self._h, out))
        return float(out[0]), float(out[1]), float(out[2])

    iteration = property(lambda s: lib.pb_backend_iteration(s._h))
    is_fused = property(lambda s: bool(lib.pb_backend_is_fused(s._h)))
    launch_count = property(lambda s: lib.pb_backend_launch_count(s._h))
    one_pass_iterations = property(lambda s: lib.pb_backend_one_pass_iterations(s._h))
    gpu_mem_amount = property(lambda s: lib.pb_backend_gpu_mem_amount(s._h))

    def current_solution(self, with_constraints=True):
        n, m = self.problem.ncols, self.problem.nrows
        x = np.empty(n, dtype=np.float32)
        y = np.empty(m, dtype=np.float32)
        z = np.empty(m, dtype=np.float32) if with_constraints else None
        w = np.empty(n, dtype=np.float32) if with_constraints else None
        check(lib.pb_backend_current_solution(self._h, _fp(x), _fp(z) if z is not None else None, _fp(y),
                                              _fp(w) if w is not None else None))
        return x, z, y, w

    def device_iterates(self):
        dx, dy = C.c_void_p(), C.c_void_p()
        check(lib.pb_backend_device_iterates(self._h, C.byref(dx), C.byref(dy)))
        return dx.value, dy.value


class BackendPDHG(Backend):
    def __init__(self, ctx, problem, opts=None, sopts=None, comm=None):
        super().__init__(ctx)
        self.problem = problem
        self.opts = opts or pdhg_options()
        self.sopts = sopts or solver_options()
        check(lib.pb_pdhg_create(ctx._h, problem._h, C.byref(self.opts), C.byref(self.sopts), C.byref(self._h)))
        if comm is not None:
            self.SetSlab(comm)

    def SetSlab(self, comm):
        """The problem is this rank's block of image columns (pb_backend_set_slab)."""
        check(lib.pb_backend_set_slab(self._h, comm._h))
        self.comm = comm


class BackendADMM(Backend):
    def __init__(self, ctx, problem, opts=None, sopts=None, comm=None):
        super().__init__(ctx)
        self.problem = problem
        self.opts = opts or admm_options()
        self.sopts = sopts or solver_options()
        check(lib.pb_admm_create(ctx._h, problem._h, C.byref(self.opts), C.byref(self.sopts), C.byref(self._h)))
        if comm is not None:
            self.SetRowShards(comm)

    def SetRowShards(self, comm):
        """The problem is this rank's block of ROWS of K and of the f-side proxes (distributed.shard_rows); K^T r
        is summed over the ranks with NCCL, sums over rows inside the reduction kernels (pb_backend_set_slab)."""
        check(lib.pb_backend_set_slab(self._h, comm._h))
        self.comm = comm


class Solver:
    """prost::Solver (include/prost/solver.hpp:85-99, src/solver.cu)."""
    CONVERGED, STOPPED_MAX_ITERS, STOPPED_USER = 0, 1, 2

    def __init__(self, problem, backend):
        self.problem = problem
        self.backend = backend
        self.opts = backend.sopts
        self.x0 = self.y0 = None
        self._stop = None
        self._interm = None
        self.iterations = 0

    def SetOptions(self, opts, x0=None, y0=None):
        self.opts = opts
        self.x0, self.y0 = x0, y0

    def SetStoppingCallback(self, cb):
        self._stop = cb

    def SetIntermCallback(self, cb):
        self._interm = cb

    def Initialize(self):
        try:
            self.problem.Initialize()
        except ProstError as e:
            raise ProstError(e.status, f"Failed to initialize the problem. Reason: {e}") from None
        x0, y0 = self.x0, self.y0
        if self.opts.solve_dual_problem:
            self.problem.Dualize()
            x0, y0 = y0, x0
        try:
            # Solver::Initialize pushes its options into the backend (solver.cu:88-90): the solver's
            # tolerances govern the stopping test even if the backend was constructed with others
            check(lib.pb_backend_set_solver_options(self.backend._h, C.byref(self.opts)))
            self.backend.sopts = self.opts
            self.backend.Initialize(x0, y0)
        except ProstError as e:
            raise ProstError(e.status, f"Failed to initialize the backend. Reason: {e}") from None

    def Solve(self):
        n, m = self.problem.ncols, self.problem.nrows
        # result vectors in pinned memory (pool with exact-size reuse): one DMA each instead of a staged copy
        # into cold pageable pages; PB_PINNED_RESULTS=0 gives plain numpy arrays like the reference's std::vector
        alloc = np.empty if os.environ.get("PB_PINNED_RESULTS", "1") == "0" else pinned_empty
        x, w = alloc(n, np.float32), alloc(n, np.float32)
        y, z = alloc(m, np.float32), alloc(m, np.float32)

        def stop(_user):
            return int(bool(self._stop())) if self._stop else 0

        def interm(_user, it, p, npr, d, nd):
            if not self._interm:
                return 0
            pa = np.ctypeslib.as_array(p, shape=(npr,))
            da = np.ctypeslib.as_array(d, shape=(nd,))
            return int(bool(self._interm(it, pa, da)))

        # no user callback -> NULL (the C loop then knows that nothing can stop it between two events)
        scb = _capi.STOPPING_CB(stop) if self._stop else _capi.STOPPING_CB()
        icb = _capi.INTERM_CB(interm)
        result, iters = C.c_int(), C.c_int()
        check(lib.pb_solver_solve(self.backend._h, C.byref(self.opts), scb, icb, None, _fp(x), _fp(z), _fp(y),
                                  _fp(w), C.byref(result), C.byref(iters)))
        self.iterations = iters.value
        dual = bool(self.opts.solve_dual_problem)
        if dual:
            self.problem.Dualize()
        # cur_primal_sol() & co swap roles when the dual problem was solved (solver.cu:216-250)
        self.cur_primal_sol, self.cur_dual_sol = (y, x) if dual else (x, y)
        self.cur_primal_constr_sol, self.cur_dual_constr_sol = (w, z) if dual else (z, w)
        return result.value
