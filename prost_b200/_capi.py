"""ctypes binding of the C ABI declared in include/prost_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (``make -C prost_b200/csrc``)
into ``prost_b200/lib/libprost_b200.so``.  There is no fallback: if the library is missing,
importing this module raises, and every compute entry point fails without a CUDA device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libprost_b200.so")


class ProstError(RuntimeError):
    """Mirror of prost::Exception (include/prost/exception.hpp:29-41)."""

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first "
        "(python -c 'import __graft_entry__ as g; g.build()' or make -C prost_b200/csrc)")

lib = C.CDLL(LIB_PATH)

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_size_p = C.POINTER(C.c_size_t)
handle = C.c_void_p
handle_p = C.POINTER(C.c_void_p)


class SolverOptions(C.Structure):
    """pb_solver_options == Solver<T>::Options scalars (include/prost/solver.hpp:39-70)."""
    _fields_ = [("tol_rel_primal", C.c_float), ("tol_rel_dual", C.c_float),
                ("tol_abs_primal", C.c_float), ("tol_abs_dual", C.c_float),
                ("max_iters", C.c_int), ("num_cback_calls", C.c_int),
                ("verbose", C.c_int), ("solve_dual_problem", C.c_int)]


class PDHGOptions(C.Structure):
    """pb_pdhg_options == BackendPDHG<T>::Options (backend_pdhg.hpp:57-82) + fuse/normest_x0."""
    _fields_ = [("tau0", C.c_double), ("sigma0", C.c_double),
                ("residual_iter", C.c_int), ("scale_steps_operator", C.c_int),
                ("alg2_gamma", C.c_float),
                ("arg_alpha0", C.c_float), ("arg_nu", C.c_float), ("arg_delta", C.c_float),
                ("arb_delta", C.c_float), ("arb_tau", C.c_float),
                ("stepsize_variant", C.c_int), ("fuse", C.c_int),
                ("normest_x0", c_float_p)]


class ADMMOptions(C.Structure):
    """pb_admm_options == BackendADMM<T>::Options (backend_admm.hpp:38-63)."""
    _fields_ = [("rho0", C.c_double), ("alpha", C.c_double),
                ("cg_tol_pow", C.c_double), ("cg_tol_min", C.c_double), ("cg_tol_max", C.c_double),
                ("cg_max_iter", C.c_int), ("residual_iter", C.c_int),
                ("arb_delta", C.c_float), ("arb_tau", C.c_float), ("arb_gamma", C.c_float)]


STOPPING_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)
INTERM_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_float_p, C.c_size_t, c_float_p, C.c_size_t)

# name -> (restype, argtypes); every symbol include/prost_b200.h declares
SIGNATURES = {
    "pb_version": (C.c_char_p, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_device_count": (C.c_int, []),
    "pb_context_create": (C.c_int, [C.c_int, C.c_void_p, handle_p]),
    "pb_context_destroy": (None, [handle]),
    "pb_context_synchronize": (C.c_int, [handle]),
    "pb_release_cached_memory": (None, []),
    "pb_context_stream": (C.c_void_p, [handle]),
    "pb_context_device": (C.c_int, [handle]),
    "pb_host_alloc": (C.c_int, [C.c_size_t, handle_p]),
    "pb_host_free": (None, [handle]),
    "pb_malloc": (C.c_int, [handle, C.c_size_t, handle_p]),
    "pb_free": (C.c_int, [handle, C.c_void_p]),
    "pb_memcpy_h2d": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pb_memcpy_d2h": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.c_size_t]),
    "pb_block_create_gradient2d": (C.c_int, [handle] + [C.c_size_t] * 5 + [C.c_int, handle_p]),
    "pb_block_create_gradient3d": (C.c_int, [handle] + [C.c_size_t] * 5 + [C.c_int, handle_p]),
    "pb_block_create_diags": (C.c_int, [handle] + [C.c_size_t] * 5 + [c_i64_p, c_float_p, handle_p]),
    "pb_block_create_sparse_csc": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                             c_float_p, c_i32_p, c_i32_p, handle_p]),
    "pb_block_create_dense": (C.c_int, [handle] + [C.c_size_t] * 4 + [c_float_p, handle_p]),
    "pb_block_create_sparse_kron_id": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                                 c_float_p, c_i32_p, c_i32_p, handle_p]),
    "pb_block_create_id_kron_sparse": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                                 c_float_p, c_i32_p, c_i32_p, handle_p]),
    "pb_block_create_dense_kron_id": (C.c_int, [handle] + [C.c_size_t] * 5 + [c_float_p, handle_p]),
    "pb_block_create_id_kron_dense": (C.c_int, [handle] + [C.c_size_t] * 5 + [c_float_p, handle_p]),
    "pb_block_create_zero": (C.c_int, [handle] + [C.c_size_t] * 4 + [handle_p]),
    "pb_block_destroy": (None, [handle]),
    "pb_block_row": (C.c_size_t, [handle]),
    "pb_block_col": (C.c_size_t, [handle]),
    "pb_block_nrows": (C.c_size_t, [handle]),
    "pb_block_ncols": (C.c_size_t, [handle]),
    "pb_block_row_sum": (C.c_float, [handle, C.c_size_t, C.c_float]),
    "pb_block_col_sum": (C.c_float, [handle, C.c_size_t, C.c_float]),
    "pb_block_gpu_mem_amount": (C.c_size_t, [handle]),
    "pb_linop_create": (C.c_int, [handle, handle_p]),
    "pb_linop_destroy": (None, [handle]),
    "pb_linop_add_block": (C.c_int, [handle, handle]),
    "pb_linop_initialize": (C.c_int, [handle]),
    "pb_linop_nrows": (C.c_size_t, [handle]),
    "pb_linop_ncols": (C.c_size_t, [handle]),
    "pb_linop_eval": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.c_float, C.c_int]),
    "pb_linop_eval_host": (C.c_int, [handle, c_float_p, c_float_p, C.c_int, c_double_p]),
    "pb_linop_row_sum": (C.c_float, [handle, C.c_size_t, C.c_float]),
    "pb_linop_col_sum": (C.c_float, [handle, C.c_size_t, C.c_float]),
    "pb_linop_row_sums": (C.c_int, [handle, C.c_float, c_float_p]),
    "pb_linop_col_sums": (C.c_int, [handle, C.c_float, c_float_p]),
    "pb_function1d_from_name": (C.c_int, [C.c_char_p]),
    "pb_prox_create_elem_1d": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                         C.c_int, C.POINTER(c_float_p), c_size_p, handle_p]),
    "pb_prox_create_elem_norm2": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                            C.c_int, C.POINTER(c_float_p), c_size_p, handle_p]),
    "pb_prox_create_ind_simplex": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                             handle_p]),
    "pb_prox_create_ind_sum": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, handle_p]),
    "pb_prox_create_ind_sum_indexed": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                                 C.POINTER(C.c_ulonglong), C.c_float, C.c_size_t, C.c_size_t,
                                                 C.POINTER(C.c_ulonglong), C.c_float, handle_p]),
    "pb_prox_create_spectral": (C.c_int, [handle, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.POINTER(c_float_p), c_size_p, handle_p]),
    "pb_prox_create_ind_range": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p,
                                           c_i32_p, c_i32_p, c_float_p, handle_p]),
    "pb_prox_create_ind_epi_conjquad_1d": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                                     C.POINTER(c_float_p), c_size_p, handle_p]),
    "pb_prox_create_ind_halfspace": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                               c_float_p, C.c_size_t, c_float_p, C.c_size_t, handle_p]),
    "pb_prox_create_ind_soc": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_float,
                                         handle_p]),
    "pb_prox_create_ind_epi_quad": (C.c_int, [handle, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                              c_float_p, C.c_size_t, c_float_p, C.c_size_t, c_float_p,
                                              C.c_size_t, handle_p]),
    "pb_prox_create_transform": (C.c_int, [handle, handle, C.POINTER(c_float_p), c_size_p, handle_p]),
    "pb_prox_create_moreau": (C.c_int, [handle, handle, handle_p]),
    "pb_prox_create_permute": (C.c_int, [handle, handle, c_int_p, C.c_size_t, handle_p]),
    "pb_prox_create_zero": (C.c_int, [handle, C.c_size_t, C.c_size_t, handle_p]),
    "pb_prox_destroy": (None, [handle]),
    "pb_prox_index": (C.c_size_t, [handle]),
    "pb_prox_size": (C.c_size_t, [handle]),
    "pb_prox_diagsteps": (C.c_int, [handle]),
    "pb_prox_gpu_mem_amount": (C.c_size_t, [handle]),
    "pb_prox_eval": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int]),
    "pb_prox_eval_host": (C.c_int, [handle, c_float_p, c_float_p, c_float_p, C.c_size_t, C.c_float,
                                    C.c_int, c_double_p]),
    "pb_problem_create": (C.c_int, [handle, handle_p]),
    "pb_problem_destroy": (None, [handle]),
    "pb_problem_add_block": (C.c_int, [handle, handle]),
    "pb_problem_add_prox_g": (C.c_int, [handle, handle]),
    "pb_problem_add_prox_f": (C.c_int, [handle, handle]),
    "pb_problem_add_prox_gstar": (C.c_int, [handle, handle]),
    "pb_problem_add_prox_fstar": (C.c_int, [handle, handle]),
    "pb_problem_set_dimensions": (C.c_int, [handle, C.c_size_t, C.c_size_t]),
    "pb_problem_set_scaling_alpha": (C.c_int, [handle, C.c_float]),
    "pb_problem_set_scaling_identity": (C.c_int, [handle]),
    "pb_problem_set_scaling_custom": (C.c_int, [handle, c_float_p, C.c_size_t, c_float_p, C.c_size_t]),
    "pb_problem_initialize": (C.c_int, [handle]),
    "pb_problem_dualize": (C.c_int, [handle]),
    "pb_problem_nrows": (C.c_size_t, [handle]),
    "pb_problem_ncols": (C.c_size_t, [handle]),
    "pb_problem_gpu_mem_amount": (C.c_size_t, [handle]),
    "pb_problem_normest": (C.c_int, [handle, C.c_float, C.c_int, c_float_p, c_float_p]),
    "pb_problem_get_scaling": (C.c_int, [handle, c_float_p, c_float_p]),
    "pb_solver_default_options": (None, [C.POINTER(SolverOptions)]),
    "pb_pdhg_default_options": (None, [C.POINTER(PDHGOptions)]),
    "pb_admm_default_options": (None, [C.POINTER(ADMMOptions)]),
    "pb_pdhg_create": (C.c_int, [handle, handle, C.POINTER(PDHGOptions), C.POINTER(SolverOptions), handle_p]),
    "pb_admm_create": (C.c_int, [handle, handle, C.POINTER(ADMMOptions), C.POINTER(SolverOptions), handle_p]),
    "pb_backend_destroy": (None, [handle]),
    "pb_backend_set_solver_options": (C.c_int, [handle, C.POINTER(SolverOptions)]),
    "pb_backend_initialize": (C.c_int, [handle, c_float_p, C.c_size_t, c_float_p, C.c_size_t]),
    "pb_backend_iterate": (C.c_int, [handle, C.c_int]),
    "pb_backend_profile": (C.c_int, [handle, C.c_int, c_float_p]),
    "pb_backend_profile_detail": (C.c_int, [handle, C.c_int, c_float_p]),
    "pb_backend_residuals": (C.c_int, [handle, c_float_p]),
    "pb_backend_stepsizes": (C.c_int, [handle, c_double_p]),
    "pb_backend_iteration": (C.c_size_t, [handle]),
    "pb_backend_current_solution": (C.c_int, [handle, c_float_p, c_float_p, c_float_p, c_float_p]),
    "pb_backend_gpu_mem_amount": (C.c_size_t, [handle]),
    "pb_backend_is_fused": (C.c_int, [handle]),
    "pb_backend_launch_count": (C.c_ulonglong, [handle]),
    "pb_backend_one_pass_iterations": (C.c_ulonglong, [handle]),
    "pb_backend_device_iterates": (C.c_int, [handle, handle_p, handle_p]),
    "pb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "pb_comm_create": (C.c_int, [handle, C.c_int, C.c_int, C.c_void_p, handle_p]),
    "pb_comm_destroy": (None, [handle]),
    "pb_comm_rank": (C.c_int, [handle]),
    "pb_comm_world": (C.c_int, [handle]),
    "pb_comm_peer_to_peer": (C.c_int, [handle]),
    "pb_comm_barrier": (C.c_int, [handle]),
    "pb_comm_allreduce_sum": (C.c_int, [handle, c_double_p, C.c_size_t]),
    "pb_backend_set_slab": (C.c_int, [handle, handle]),
    "pb_ring_trace_read": (C.c_uint, [C.POINTER(C.c_ulonglong), C.c_uint]),
    "pb_solver_solve": (C.c_int, [handle, C.POINTER(SolverOptions), STOPPING_CB, INTERM_CB, C.c_void_p,
                                  c_float_p, c_float_p, c_float_p, c_float_p, c_int_p, c_int_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)        # AttributeError here == the library does not export the ABI
    _fn.restype = _res
    _fn.argtypes = _args


def check(status):
    """Turn a pb_status into a ProstError carrying pb_last_error()."""
    if status != 0:
        raise ProstError(status, lib.pb_last_error().decode("utf-8", "replace"))
