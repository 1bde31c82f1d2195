"""ctypes binding of include/abft_cuda.h (libabftcuda.so, built in-tree).

There is no fallback: if the library is missing, `load()` raises; if no CUDA
device is usable, context creation fails with the library's error text.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABFT_CUDA_LIB") or os.path.join(HERE, "libabftcuda.so")   # override: kernel experiments

MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c", "secdedv"]
FATAL_KINDS = {1, 4, 5, 6, 7, 8, 9, 11, 12, 13}

# every extern "C" symbol include/abft_cuda.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "abft_cuda_abi_version", "abft_cuda_last_error", "abft_cuda_mode_name", "abft_cuda_mode_from_name",
    "abft_cuda_device_count", "abft_cuda_format_event", "abft_cuda_event_is_fatal",
    "abft_cuda_ctx_create", "abft_cuda_ctx_destroy", "abft_cuda_ctx_mode", "abft_cuda_ctx_sync",
    "abft_cuda_ctx_stream", "abft_cuda_phases_enable", "abft_cuda_phases_read",
    "abft_cuda_matrix_create", "abft_cuda_matrix_create_part", "abft_cuda_matrix_destroy",
    "abft_cuda_plan_tiles", "abft_cuda_plan_matrix", "abft_cuda_matrix_rows", "abft_cuda_matrix_nnz", "abft_cuda_matrix_read", "abft_cuda_matrix_poke_rowptr",
    "abft_cuda_matrix_poke_col", "abft_cuda_matrix_poke_value",
    "abft_cuda_vector_create", "abft_cuda_vector_wrap", "abft_cuda_vector_view", "abft_cuda_vector_destroy",
    "abft_cuda_vector_len", "abft_cuda_vector_flip_bit", "abft_cuda_vector_dev_ptr", "abft_cuda_vector_map", "abft_cuda_vector_unmap",
    "abft_cuda_vector_copy", "abft_cuda_vector_download_async", "abft_cuda_vector_upload_async", "abft_cuda_host_alloc",
    "abft_cuda_host_free",
    "abft_cuda_spmv", "abft_cuda_dot", "abft_cuda_calc_xr", "abft_cuda_calc_p",
    "abft_cuda_dot_async", "abft_cuda_calc_xr_async", "abft_cuda_scalar_dev_ptr", "abft_cuda_scalar_host",
    "abft_cuda_flip_bit", "abft_cuda_poll_events", "abft_cuda_events_pending",
    "abft_cuda_cg_solve", "abft_cuda_cg_host_loop", "abft_cuda_dist_cg_host_loop", "abft_cuda_launch_count", "abft_cuda_matrix_tiles",
    "abft_cuda_comm_create", "abft_cuda_comm_destroy", "abft_cuda_comm_window_handle", "abft_cuda_vector_ipc_handle",
    "abft_cuda_comm_window_ptr", "abft_cuda_comm_attach_self", "abft_cuda_comm_attach_peer",
    "abft_cuda_comm_attach_peer_ptrs", "abft_cuda_comm_set_plan", "abft_cuda_dist_push", "abft_cuda_dist_calc_p",
    "abft_cuda_dist_spmv", "abft_cuda_dist_dot", "abft_cuda_dist_calc_xr", "abft_cuda_dist_cg_solve",
    "abft_cuda_dist_dot_async", "abft_cuda_dist_calc_xr_async",
    "abft_cuda_coo_matrix_create", "abft_cuda_coo_matrix_destroy", "abft_cuda_coo_matrix_nnz",
    "abft_cuda_coo_matrix_read", "abft_cuda_coo_spmv", "abft_cuda_coo_flip_bit", "abft_cuda_coo_format_event",
]


class Event(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_uint32), ("bit", C.c_uint32)]


def build(force: bool = False) -> str:
    """Compile libabftcuda.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", HERE, "-s", "lib"], check=True)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `make -C {HERE} lib` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, f64 = C.c_void_p, C.c_int, C.c_uint32, C.c_double
    u32p, f64p = C.POINTER(C.c_uint32), C.POINTER(C.c_double)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("abft_cuda_abi_version", i32)
    sig("abft_cuda_last_error", C.c_char_p)
    sig("abft_cuda_mode_name", C.c_char_p, i32)
    sig("abft_cuda_mode_from_name", i32, C.c_char_p)
    sig("abft_cuda_device_count", i32)
    sig("abft_cuda_format_event", i32, C.POINTER(Event), C.c_char_p, i32)
    sig("abft_cuda_event_is_fatal", i32, i32)
    sig("abft_cuda_ctx_create", vp, i32, i32)
    sig("abft_cuda_ctx_destroy", None, vp)
    sig("abft_cuda_ctx_mode", i32, vp)
    sig("abft_cuda_ctx_sync", i32, vp)
    sig("abft_cuda_ctx_stream", vp, vp)
    sig("abft_cuda_phases_enable", i32, vp, i32)
    sig("abft_cuda_phases_read", i32, vp, vp, i32)
    sig("abft_cuda_matrix_create", vp, vp, vp, vp, vp, i32, i32)
    sig("abft_cuda_matrix_create_part", vp, vp, vp, vp, vp, i32, i32, i32, i32, C.c_int64)
    sig("abft_cuda_matrix_destroy", None, vp, vp)
    sig("abft_cuda_plan_tiles", i32, vp, i32, i32, vp, i32)
    sig("abft_cuda_plan_matrix", i32, vp, i32, i32, vp, i32, C.POINTER(C.c_int))
    sig("abft_cuda_matrix_rows", i32, vp)
    sig("abft_cuda_matrix_nnz", i32, vp)
    sig("abft_cuda_matrix_tiles", i32, vp)
    sig("abft_cuda_matrix_read", i32, vp, vp, vp, vp, vp)
    sig("abft_cuda_matrix_poke_rowptr", i32, vp, vp, i32, u32)
    sig("abft_cuda_matrix_poke_col", i32, vp, vp, i32, u32)
    sig("abft_cuda_matrix_poke_value", i32, vp, vp, i32, f64)
    sig("abft_cuda_vector_create", vp, vp, i32)
    sig("abft_cuda_vector_wrap", vp, vp, vp, i32)
    sig("abft_cuda_vector_view", vp, vp, vp, i32, i32)
    sig("abft_cuda_vector_destroy", None, vp, vp)
    sig("abft_cuda_vector_len", i32, vp)
    sig("abft_cuda_vector_flip_bit", i32, vp, vp, i32, i32)
    sig("abft_cuda_vector_dev_ptr", vp, vp)
    sig("abft_cuda_vector_map", f64p, vp, vp)
    sig("abft_cuda_vector_unmap", i32, vp, vp, vp)
    sig("abft_cuda_vector_copy", i32, vp, vp, vp)
    sig("abft_cuda_vector_download_async", i32, vp, vp, vp)
    sig("abft_cuda_vector_upload_async", i32, vp, vp, vp)
    sig("abft_cuda_host_alloc", vp, C.c_size_t)
    sig("abft_cuda_host_free", None, vp)
    sig("abft_cuda_spmv", i32, vp, vp, vp, vp)
    sig("abft_cuda_dot", i32, vp, vp, vp, f64p)
    sig("abft_cuda_calc_xr", i32, vp, vp, vp, vp, vp, f64, f64p)
    sig("abft_cuda_calc_p", i32, vp, vp, vp, f64)
    sig("abft_cuda_dot_async", i32, vp, vp, vp, i32)
    sig("abft_cuda_calc_xr_async", i32, vp, vp, vp, vp, vp, f64, i32)
    sig("abft_cuda_scalar_dev_ptr", vp, vp, i32)
    sig("abft_cuda_scalar_host", f64, vp, i32)
    sig("abft_cuda_flip_bit", i32, vp, vp, C.c_int64, i32)
    sig("abft_cuda_poll_events", i32, vp, C.POINTER(Event), i32, C.POINTER(C.c_int))
    sig("abft_cuda_events_pending", i32, vp)
    sig("abft_cuda_cg_solve", i32, vp, vp, vp, vp, vp, vp, vp, i32, f64, i32, f64p)
    sig("abft_cuda_cg_host_loop", i32, vp, vp, vp, vp, vp, vp, i32, f64, f64p, vp)
    sig("abft_cuda_dist_cg_host_loop", i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, f64, f64p, vp)
    sig("abft_cuda_launch_count", C.c_uint64, vp)
    i32p = C.POINTER(C.c_int)
    sig("abft_cuda_comm_create", vp, vp, i32, i32)
    sig("abft_cuda_comm_destroy", None, vp, vp)
    sig("abft_cuda_comm_window_handle", i32, vp, vp, C.c_char_p)
    sig("abft_cuda_vector_ipc_handle", i32, vp, vp, C.c_char_p)
    sig("abft_cuda_comm_window_ptr", vp, vp)
    sig("abft_cuda_comm_attach_self", i32, vp, vp, vp)
    sig("abft_cuda_comm_attach_peer", i32, vp, vp, i32, C.c_char_p, C.c_char_p)
    sig("abft_cuda_comm_attach_peer_ptrs", i32, vp, vp, i32, vp, vp, i32)
    sig("abft_cuda_comm_set_plan", i32, vp, vp, i32p, i32p, C.c_uint, i32)
    sig("abft_cuda_dist_push", i32, vp, vp, vp)
    sig("abft_cuda_dist_calc_p", i32, vp, vp, vp, vp, f64)
    sig("abft_cuda_dist_spmv", i32, vp, vp, vp, vp, vp)
    sig("abft_cuda_dist_dot", i32, vp, vp, vp, vp, f64p)
    sig("abft_cuda_dist_calc_xr", i32, vp, vp, vp, vp, vp, vp, f64, f64p)
    sig("abft_cuda_dist_dot_async", i32, vp, vp, vp, vp, i32p)
    sig("abft_cuda_dist_calc_xr_async", i32, vp, vp, vp, vp, vp, vp, f64, i32p)
    sig("abft_cuda_dist_cg_solve", i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f64, i32, f64p)
    sig("abft_cuda_coo_matrix_create", vp, vp, vp, vp, vp, i32, i32)
    sig("abft_cuda_coo_matrix_destroy", None, vp, vp)
    sig("abft_cuda_coo_matrix_nnz", i32, vp)
    sig("abft_cuda_coo_matrix_read", i32, vp, vp, vp)
    sig("abft_cuda_coo_spmv", i32, vp, vp, vp, vp)
    sig("abft_cuda_coo_flip_bit", i32, vp, vp, i32, i32)
    sig("abft_cuda_coo_format_event", i32, C.POINTER(Event), C.c_char_p, i32)
    _lib = L
    return L


def last_error() -> str:
    return load().abft_cuda_last_error().decode(errors="replace")
