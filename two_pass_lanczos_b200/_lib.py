"""ctypes binding of libtplanczos.so -- the C ABI declared in include/tplanczos.h.

There is NO fallback: if the shared library is missing this module raises, and every compute call
returns TPL_ERR_CUDA when no GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os

from .error import raise_for_status

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtplanczos.so")

c_dp = C.POINTER(C.c_double)
c_u64p = C.POINTER(C.c_uint64)
c_u32p = C.POINTER(C.c_uint32)
c_szp = C.POINTER(C.c_size_t)

FTK_FN = C.CFUNCTYPE(C.c_int, c_dp, C.c_size_t, c_dp, C.c_size_t, c_dp, c_szp, C.c_void_p)
STEP_CB = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, c_dp, c_dp, C.c_void_p)

# name -> (restype, argtypes); kept in sync with include/tplanczos.h (tests/test_abi.py checks it)
SIGNATURES = {
    "tpl_last_error_message": (C.c_char_p, []),
    "tpl_version": (C.c_char_p, []),
    "tpl_load_kkt": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "tpl_write_kkt_binary": (C.c_int, [C.c_char_p, C.c_size_t, C.c_size_t, c_u32p, c_u32p, c_dp, C.c_size_t]),
    "tpl_kkt_save_binary": (C.c_int, [C.c_void_p, C.c_char_p]),
    "tpl_load_kkt_binary": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "tpl_kkt_free": (None, [C.c_void_p]),
    "tpl_kkt_num_nodes": (C.c_size_t, [C.c_void_p]),
    "tpl_kkt_num_arcs": (C.c_size_t, [C.c_void_p]),
    "tpl_kkt_num_costs": (C.c_size_t, [C.c_void_p]),
    "tpl_kkt_nnz": (C.c_size_t, [C.c_void_p]),
    "tpl_kkt_csc": (C.c_int, [C.c_void_p, c_szp, c_szp, C.POINTER(c_u64p), C.POINTER(c_u64p), C.POINTER(c_dp)]),
    "tpl_kkt_incidence": (C.c_int, [C.c_void_p, C.POINTER(c_u32p), C.POINTER(c_u32p), C.POINTER(c_dp), c_szp,
                                    C.POINTER(C.c_int)]),
    "tpl_op_from_csc": (C.c_int, [C.c_size_t, c_u64p, c_u64p, c_dp, C.c_int, C.POINTER(C.c_void_p)]),
    "tpl_op_from_dense": (C.c_int, [C.c_size_t, c_dp, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "tpl_op_from_dense_hermitian": (C.c_int, [C.c_size_t, c_dp, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "tpl_op_is_complex": (C.c_int, [C.c_void_p]),
    "tpl_op_from_diagonal": (C.c_int, [C.c_size_t, c_dp, C.c_int, C.POINTER(C.c_void_p)]),
    "tpl_op_from_kkt": (C.c_int, [C.c_size_t, C.c_size_t, c_u32p, c_u32p, c_dp, C.c_size_t, C.c_int,
                                  C.POINTER(C.c_void_p)]),
    "tpl_op_from_kkt_system": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "tpl_op_free": (None, [C.c_void_p]),
    "tpl_op_nrows": (C.c_size_t, [C.c_void_p]),
    "tpl_op_check_len": (C.c_int, [C.c_void_p, C.c_size_t]),
    "tpl_op_format": (C.c_int, [C.c_void_p]),
    "tpl_op_device": (C.c_int, [C.c_void_p]),
    "tpl_op_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tpl_op_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tpl_op_last_timing": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp]),
    "tpl_op_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "tpl_op_matrix_bytes": (C.c_uint64, [C.c_void_p]),
    "tpl_op_device_bytes": (C.c_uint64, [C.c_void_p]),
    "tpl_op_set_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "tpl_op_kernel_shape": (C.c_char_p, [C.c_void_p]),
    "tpl_tiles_plan": (C.c_int, [C.c_size_t, C.c_size_t, c_u32p, c_u32p, C.c_int, C.c_uint32, C.c_int, c_u64p]),
    "tpl_blocks_plan": (C.c_int, [C.c_size_t, C.c_size_t, c_u32p, c_u32p, c_dp, C.c_size_t, C.c_int, C.c_size_t, C.c_int, c_u64p]),
    "tpl_cells_plan": (C.c_int, [C.c_size_t, C.c_size_t, c_u32p, c_u32p, C.c_int, C.c_size_t, c_u64p]),
    "tpl_op_layout_check": (C.c_int, [C.c_void_p, c_u32p, c_u32p, c_dp, C.c_size_t, c_u64p]),
    "tpl_op_trace_enable": (C.c_int, [C.c_void_p, C.c_size_t]),
    "tpl_op_trace_read": (C.c_int, [C.c_void_p, c_u64p, C.c_size_t, c_szp, c_szp, c_szp]),
    "tpl_pass_one": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, c_dp, c_dp, c_szp, c_dp]),
    "tpl_pass_two": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, c_dp, C.c_size_t, C.c_double, c_dp, C.c_size_t,
                               C.c_void_p, C.c_void_p, C.c_size_t]),
    "tpl_standard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, c_dp, c_dp, c_szp, c_dp,
                               STEP_CB, C.c_void_p]),
    "tpl_lanczos": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, FTK_FN, C.c_void_p, C.c_void_p]),
    "tpl_lanczos_two_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, FTK_FN, C.c_void_p, C.c_void_p]),
    "tpl_lanczos_sweep": (C.c_int, [C.c_void_p, C.c_void_p, c_szp, C.c_size_t, FTK_FN, C.c_void_p, C.c_void_p, C.c_size_t]),
    "tpl_lanczos_two_pass_sweep": (C.c_int, [C.c_void_p, C.c_void_p, c_szp, C.c_size_t, FTK_FN, C.c_void_p, C.c_void_p, C.c_size_t]),
    "tpl_ftk_inv": (C.c_int, [c_dp, C.c_size_t, c_dp, C.c_size_t, c_dp, c_szp, C.c_void_p]),
    "tpl_ftk_exp": (C.c_int, [c_dp, C.c_size_t, c_dp, C.c_size_t, c_dp, c_szp, C.c_void_p]),
    "tpl_ftk_inv_residuals": (C.c_int, [c_dp, C.c_size_t, c_dp, C.c_size_t, C.c_double, c_dp]),
    "tpl_lanczos_two_pass_inv_adaptive": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p, c_szp, c_dp]),
    "tpl_ftk_square": (C.c_int, [c_dp, C.c_size_t, c_dp, C.c_size_t, c_dp, c_szp, C.c_void_p]),
    "tpl_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "tpl_op_from_kkt_sharded": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, c_u32p, c_u32p, c_dp,
                                          C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8),
                                          C.POINTER(C.c_void_p)]),
    "tpl_op_fabric_export": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint8)]),
    "tpl_op_fabric_import": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint8), C.c_int]),
    "tpl_op_shard_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), c_szp, c_szp]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m two_pass_lanczos_b200.build` "
                "(the engine is CUDA-only; there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code: int):
    if code != 0:
        raise_for_status(code, load().tpl_last_error_message().decode(errors="replace"))
