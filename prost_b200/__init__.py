"""prost_b200 -- B200-native primal-dual solver core behind prost's operator/prox/backend API.

The package is a thin ctypes mirror of the C ABI in include/prost_b200.h; the hot path (fused
PDHG passes, operator applies, separable proxes, residual reductions) is hand-written CUDA for
sm_100a in prost_b200/csrc.  There is no CPU fallback.
"""
from .api import (ADMMOptions, Backend, Comm, BackendADMM, BackendPDHG, Block, BlockDense, BlockDenseKronId,
                  BlockIdKronDense, BlockIdKronSparse, BlockSparseKronId, BlockDiags,
                  BlockGradient2D, BlockGradient3D, BlockSparse, BlockZero, Context, LinearOperator,
                  PDHGOptions, Problem, ProstError, Prox, ProxElemOperation1D, ProxElemOperationIndSimplex,
                  ProxElemOperationIndSum, ProxElemOperationSpectral,
                  ProxElemOperationNorm2, ProxIndEpiQuad, ProxIndEpiConjQuad1D, ProxIndHalfspace, ProxIndRange, ProxIndSum, ProxIndSOC, ProxMoreau, ProxPermute, ProxTransform, ProxZero, Solver,
                  SolverOptions, admm_options, pdhg_options, solver_options)
from .factory import create_block, create_linop, create_problem, create_prox
from ._capi import LIB_PATH, lib


def version():
    return lib.pb_version().decode()


def device_count():
    return lib.pb_device_count()


def release_cached_memory():
    """Returns the device buffers cached from destroyed problems / backends to the driver."""
    lib.pb_release_cached_memory()
