"""ctypes binding of libtripsb200.so (the C ABI declared in include/tripsb200.h).

There is no CPU fallback: if the shared library is missing, or the device is not sm_100, every entry point
raises.  `lib()` loads lazily so that host-only logic (argument checks, the NumPy projected problem) can be
imported and tested on a machine without a GPU.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtripsb200.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/tripsb200.h one to one
SIGNATURES = {
    "tb200_last_error": (ctypes.c_char_p, []),
    "tb200_version": (c_int, []),
    "tb200_device_info": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_require_sm100": (c_int, []),
    "tb200_fp64_peak_instructions": (c_i64, [c_int]),
    "tb200_fp64_peak_run": (c_int, [c_int, c_ptr, c_ptr]),
    "tb200_spmv_workspace_len": (c_i64, [c_i64]),
    "tb200_spmv_launches": (c_int, [c_int]),
    "tb200_spmv_set_variant": (c_int, [c_int]),
    "tb200_spmv_csr_f64": (c_int, [c_int, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_spmv_csr_f32s": (c_int, [c_int, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_spmv_sell_f64": (c_int, [c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_spmv_sell_f32s": (c_int, [c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_gk_step_sell_f64": (c_int, [c_i64, c_i64] + [c_ptr] * 18),
    "tb200_reduce_finalize": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "tb200_reduce_workspace_len": (c_i64, []),
    "tb200_vec_div": (c_int, [c_i64, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr]),
    "tb200_vec_axpy": (c_int, [c_i64, c_dbl, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_vec_dot_partials": (c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_vec_norm2": (c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_vec_dot": (c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_vec_diffnorm2": (c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_vec_binary": (c_int, [c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_irls_weights": (c_int, [c_i64, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr]),
    "tb200_basis_workspace_len": (c_i64, [c_i64]),
    "tb200_basis_dots": (c_int, [c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_basis_combine": (c_int, [c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_gram_workspace_len": (c_i64, [c_i64]),
    "tb200_gram_set_bulk": (None, [c_int]),
    "tb200_gram_set_block": (None, [c_int]),
    "tb200_basis_set_vec2": (None, [c_int]),
    "tb200_weighted_gram": (c_int, [c_i64, c_i64, c_ptr, c_i64, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_weighted_gram_panel": (c_int, [c_i64, c_i64, c_ptr, c_i64, c_ptr, c_int, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_gram_factor_dd": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_count_rows": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_fill_rows": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_count_cols": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_fill_cols": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    "tb200_ctfan_count_rows": (c_int, [c_dbl, c_dbl, c_dbl, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ctfan_fill_rows": (c_int, [c_dbl, c_dbl, c_dbl, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    "tb200_ctfan_count_cols": (c_int, [c_dbl, c_dbl, c_dbl, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ctfan_fill_cols": (c_int, [c_dbl, c_dbl, c_dbl, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_geometry": (c_int, [c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_count_rows_first": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_fill_rows_aligned": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr]),
    "tb200_ct_fill_rows_aligned_vals": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_spmv_sell_f64": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr,
                                       c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_gk_step_sell_ct_f64": (c_int, [c_int, c_int, c_int, c_int] + [c_ptr] * 22),
    "tb200_ct_forward_f64": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_forward_rays_workspace_len": (c_i64, [c_int, c_int]),
    "tb200_ct_forward_set_tuning": (c_int, [c_dbl, c_int]),
    "tb200_ct_forward_rays_plan": (c_int, [c_int, c_int, c_ptr, c_ptr]),
    "tb200_ct_forward_rays_f64": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ctfan_forward_rays_f64": (c_int, [c_dbl, c_dbl, c_dbl, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr,
                                             c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_backproject_workspace_len": (c_i64, [c_int, c_int]),
    "tb200_ct_backproject_f64": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_backproject_rows_f64": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_backproject_sharded_f64": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_dbl,
                                                 c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_ct_forward_rays_sharded_f64": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_dbl, c_ptr, c_ptr,
                                                  c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_comm_mailbox_bytes": (c_i64, []),
    "tb200_comm_handle_bytes": (c_int, []),
    "tb200_comm_max_ranks": (c_int, []),
    "tb200_comm_init": (c_int, [c_int, c_int, c_i64, c_ptr, c_ptr]),
    "tb200_comm_connect": (c_int, [c_ptr, c_ptr]),
    "tb200_comm_arena": (c_ptr, [c_ptr, c_int]),
    "tb200_comm_destroy": (c_int, [c_ptr]),
    "tb200_comm_allreduce_dd": (c_int, [c_ptr, c_int, c_i64, c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    "tb200_comm_allreduce_scale": (c_int, [c_ptr, c_int, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "tb200_comm_push": (c_int, [c_ptr, ctypes.c_uint, c_i64, c_ptr, c_i64, c_ptr]),
    "tb200_halo_exchange": (c_int, [c_ptr, c_int, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_comm_scale": (c_int, [c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "tb200_gk_step_ct_f64": (c_int, [c_int, c_int, c_int, c_int] + [c_ptr] * 17),
    "tb200_correlate2d_f64": (c_int, [c_int, c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "tb200_fd_rows": (c_i64, [c_int, c_int, c_int, c_int]),
    "tb200_fd_apply": (c_int, [c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr]),
    "tb200_fd_adjoint": (c_int, [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_cd2d_apply": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_dbl, c_dbl, c_ptr]),
    "tb200_cd2d_adjoint": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "tb200_fd1d_apply": (c_int, [c_i64, c_ptr, c_ptr, c_ptr]),
    "tb200_fd1d_adjoint": (c_int, [c_i64, c_ptr, c_ptr, c_ptr]),
}

_lib = None
_sm100_ok = False


class TB200Error(RuntimeError):
    """Non-zero status from libtripsb200."""


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TB200Error(
                f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(or __graft_entry__.build()). trips_b200 has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    return lib().tb200_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise TB200Error(f"libtripsb200 {what} failed with status {rc}: {last_error()}")


def require_device():
    """Raise unless the current CUDA device is a compute-capability 10.x part."""
    global _sm100_ok
    if not _sm100_ok:
        check(lib().tb200_require_sm100(), "tb200_require_sm100")
        _sm100_ok = True


# launch accounting for bench.py ("gpu_launches": kernels of this library enqueued in a timed region)
launch_count = 0


def count(n=1):
    global launch_count
    launch_count += n
