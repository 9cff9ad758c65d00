"""ctypes binding of liblj_b200.so -- the C ABI declared in include/lj_b200.h.

This is the only way Python code in this repository reaches the CUDA kernels, so a parity
test written against this module exercises exactly what a C++/cgo/JNI caller would bind.
There is no CPU fallback: if the shared library is missing or no CUDA device is present the
import / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LJ_B200_LIB", os.path.join(HERE, "liblj_b200.so"))  # override: A/B builds

# enums of include/lj_b200.h
LJ_OK, LJ_ERR_CUDA, LJ_ERR_BAD_ARG, LJ_ERR_CAPACITY, LJ_ERR_OVERFLOW32, LJ_ERR_NO_DEVICE, \
    LJ_ERR_INVALID_LIST = range(7)
LJ_AOS_D3, LJ_AOS_D4, LJ_SOA_D, LJ_AOS_F4, LJ_AOS_F3 = range(5)
LJ_LIST_CSR, LJ_LIST_ELL, LJ_LIST_ELL_ROWS = 0, 1, 2
LJ_VARIANT_AUTO, LJ_VARIANT_SUBWARP, LJ_VARIANT_TILE_TMA, LJ_VARIANT_NEWTON3, LJ_VARIANT_CLUSTER, \
    LJ_VARIANT_CELLTILE = range(6)
LJ_PREC_FP64, LJ_PREC_MIXED = 0, 1
LJ_PART_ALL, LJ_PART_INTERIOR, LJ_PART_BOUNDARY = 0, 1, 2
LJ_LIST_SORT_ROWS = 1
LJ_LIST_CLUSTERS = 2
LJ_LIST_TILES = 8
LJ_LIST_TILES_WIDE = 16
LJ_LIST_PER_PARTICLE_SEARCH = 4


class LjBuf(C.Structure):
    _fields_ = [("host", C.c_void_p), ("dev", C.c_void_p), ("bytes", C.c_size_t)]


class LjForceArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("p", C.c_void_p), ("pn", C.c_int64), ("dt", C.c_double),
        ("cl2", C.c_double), ("list", C.c_void_p), ("number_of_partners", C.c_void_p),
        ("pointer", C.c_void_p), ("layout", C.c_int32), ("list_layout", C.c_int32),
        ("variant", C.c_int32), ("group", C.c_int32), ("precision", C.c_int32),
        ("pointer64", C.c_int32), ("threads_per_block", C.c_int32), ("list_scalar", C.c_int32),
        ("plane_stride", C.c_int64), ("row_begin", C.c_int64), ("row_end", C.c_int64),
        ("list_entries", C.c_int64), ("ell_width", C.c_int64), ("mirror_token", C.c_uint64),
    ]


class LjHaloSeg(C.Structure):
    _fields_ = [("local_dst", C.c_void_p), ("peer_src", C.c_void_p), ("bytes", C.c_size_t),
                ("wait_flag", C.c_void_p), ("wait_value", C.c_int32), ("done_value", C.c_int32),
                ("done_flag", C.c_void_p)]


class LjDecompArgs(C.Structure):
    _fields_ = [("ngpus", C.c_int32), ("devices", C.c_void_p), ("q_xyz_host", C.c_void_p), ("pn", C.c_int64),
                ("slab_begin", C.c_void_p), ("halo_rows", C.c_int64), ("search_len", C.c_double),
                ("cutoff", C.c_double), ("dt", C.c_double), ("precision", C.c_int32), ("list_flags", C.c_int32)]


class LjListArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("pn", C.c_int64), ("layout", C.c_int32), ("half", C.c_int32),
        ("plane_stride", C.c_int64), ("search_len", C.c_double),
        ("number_of_partners", C.c_void_p), ("pointer", C.c_void_p), ("sorted_list", C.c_void_p),
        ("capacity", C.c_int64), ("pointer64", C.c_int32), ("flags", C.c_int32),
        ("row_begin", C.c_int64), ("row_end", C.c_int64),
    ]


class LjMeasureArgs(C.Structure):
    _fields_ = [
        ("q_host", C.c_void_p), ("p_host", C.c_void_p), ("pn", C.c_int64), ("layout", C.c_int32),
        ("half", C.c_int32), ("plane_stride", C.c_int64), ("dt", C.c_double), ("cl2", C.c_double),
        ("search_len", C.c_double), ("loop", C.c_int32), ("rebuild_every", C.c_int32),
        ("variant", C.c_int32), ("group", C.c_int32), ("precision", C.c_int32),
        ("threads_per_block", C.c_int32), ("use_graph", C.c_int32), ("list_flags", C.c_int32),
        ("list_host", C.c_void_p), ("number_of_partners_host", C.c_void_p),
        ("pointer_host", C.c_void_p), ("number_of_pairs_in", C.c_int64),
        ("number_of_pairs", C.c_int64), ("max_partners", C.c_int32), ("list_builds", C.c_int32),
        ("seconds_total", C.c_double), ("seconds_kernel", C.c_double),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
    ]


# every symbol include/lj_b200.h declares: (restype, argtypes)
_vp, _i32, _i64, _sz, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_double
PROTOTYPES = {
    "lj_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "lj_ctx_destroy": (C.c_int, [_vp]),
    "lj_sync": (C.c_int, [_vp, _vp]),
    "lj_last_error_string": (C.c_char_p, [_vp]),
    "lj_status_string": (C.c_char_p, [C.c_int]),
    "lj_launch_count": (_i64, [_vp]),
    "lj_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "lj_kernel_timing_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "lj_ctx_stream": (_vp, [_vp]),
    "lj_device_sm_count": (C.c_int, [_vp]),
    "lj_buf_allocate": (C.c_int, [_vp, _sz, C.POINTER(LjBuf), _vp]),
    "lj_buf_deallocate": (C.c_int, [_vp, C.POINTER(LjBuf), _vp]),
    "lj_buf_host2dev": (C.c_int, [_vp, C.POINTER(LjBuf), _sz, _sz, _vp]),
    "lj_buf_dev2host": (C.c_int, [_vp, C.POINTER(LjBuf), _sz, _sz, _vp]),
    "lj_buf_set_val32": (C.c_int, [_vp, C.POINTER(LjBuf), _sz, _sz, C.c_uint32, _vp]),
    "lj_dev_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp), _vp]),
    "lj_dev_free": (C.c_int, [_vp, _vp, _vp]),
    "lj_upload": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "lj_download": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "lj_force_step": (C.c_int, [_vp, C.POINTER(LjForceArgs), _vp]),
    "lj_force_step_part": (C.c_int, [_vp, C.POINTER(LjForceArgs), _i32, _vp]),
    "lj_force_loop": (C.c_int, [_vp, C.POINTER(LjForceArgs), C.c_int, C.c_int, _vp]),
    "lj_build_list": (C.c_int, [_vp, C.POINTER(LjListArgs), C.POINTER(_i64), _vp]),
    "lj_force_loop_soa6": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(LjForceArgs), C.c_int, _vp]),
    "lj_build_list_soa6": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(LjListArgs), C.POINTER(_i64), _vp]),
    "lj_list_invalidate": (C.c_int, [_vp]),
    "lj_list_mirror": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _dbl, _vp, _vp, _i32, _vp, _i64, _i32, C.POINTER(_i64), _vp]),
    "lj_list_mirror_token": (C.c_uint64, [_vp]),
    "lj_list_result": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i32), _vp]),
    "lj_build_ell": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _i64, C.POINTER(_i32), _vp]),
    "lj_build_ell_rows": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _i64, C.POINTER(_i32), _vp]),
    "lj_shuffle_rows": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, C.c_uint32, _vp]),
    "lj_validate_list": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _vp]),
    "lj_init_fcc": (_i64, [_dbl, _dbl, _vp, _i64, C.POINTER(_i32)]),
    "lj_paircache_write_text": (C.c_int, [C.c_char_p, _i64, _i64, _vp, _vp, _vp]),
    "lj_paircache_read_text": (C.c_int, [C.c_char_p, _i64, C.POINTER(_i64), C.POINTER(_i64), _vp, _vp, _i64, _vp, _i64]),
    "lj_pairdat_write": (C.c_int, [C.c_char_p, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "lj_pairdat_read": (C.c_int, [C.c_char_p, _i64, _i64, _i64, C.POINTER(_i64), _vp, _vp, _vp, _i64]),
    "lj_measure": (C.c_int, [_vp, C.POINTER(LjMeasureArgs)]),
    "lj_drift": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i64, _dbl, _vp]),
    "lj_max_displacement2": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i64, C.POINTER(_dbl), _vp]),
    "lj_energy": (C.c_int, [_vp, C.POINTER(LjForceArgs), C.POINTER(_dbl), C.POINTER(_dbl), _vp]),
    "lj_ipc_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "lj_ipc_free": (C.c_int, [_vp, _vp]),
    "lj_ipc_export": (C.c_int, [_vp, _vp, C.c_char_p]),
    "lj_ipc_open": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "lj_ipc_close": (C.c_int, [_vp, _vp]),
    "lj_halo_pull": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "lj_decomp_plan_fcc": (C.c_int, [_dbl, _dbl, _i32, _dbl, _vp, C.POINTER(_i64)]),
    "lj_decomp_create": (C.c_int, [C.POINTER(_vp), C.POINTER(LjDecompArgs)]),
    "lj_decomp_step": (C.c_int, [_vp, _i32, _i32, _i32]),
    "lj_decomp_md": (C.c_int, [_vp, _i32, _i32, _i32]),
    "lj_decomp_rebuild": (C.c_int, [_vp]),
    "lj_decomp_sync": (C.c_int, [_vp]),
    "lj_decomp_gather": (C.c_int, [_vp, _vp, _vp]),
    "lj_decomp_pairs": (_i64, [_vp]),
    "lj_decomp_launch_count": (_i64, [_vp]),
    "lj_decomp_last_error": (C.c_char_p, [_vp]),
    "lj_decomp_destroy": (C.c_int, [_vp]),
    "lj_flag_set": (C.c_int, [_vp, _vp, _i32, _vp]),
    "lj_flag_wait": (C.c_int, [_vp, _vp, _i32, _vp]),
    "lj_halo_pull_sync": (C.c_int, [_vp, C.POINTER(LjHaloSeg), _i32, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the product library and type every exported function.  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                LIB_PATH + " is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or make -C lj_gpu_b200/csrc).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
