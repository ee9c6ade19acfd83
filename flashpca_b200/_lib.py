"""ctypes binding of the C ABI declared in include/flashpca_b200.h.

There is deliberately no fallback: if libflashpca_b200.so is missing this
module raises, and if no CUDA device is usable every call fails.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libflashpca_b200.so")

_c = ctypes
_vp, _u64, _u32, _i, _d = _c.c_void_p, _c.c_uint64, _c.c_uint32, _c.c_int, _c.c_double
_dp = _c.POINTER(_c.c_double)

# name -> (restype, argtypes); mirrors include/flashpca_b200.h one to one
SIGNATURES = {
    "fpb_abi_version": (_i, []),
    "fpb_last_error": (_c.c_char_p, [_vp]),
    "fpb_create": (_i, [_c.POINTER(_vp), _vp, _u64, _u64, _i, _vp, _i]),
    "fpb_create_from_file": (_i, [_c.POINTER(_vp), _c.c_char_p, _u64, _u64, _u64, _i, _vp, _i]),
    "fpb_create_streaming": (_i, [_c.POINTER(_vp), _c.c_char_p, _u64, _u64, _u64, _u64, _i, _vp, _i]),
    "fpb_create_synthetic": (_i, [_c.POINTER(_vp), _u64, _u64, _u64, _vp, _vp, _u32, _u32, _u64,
                                  _i, _i]),
    "fpb_create_dense": (_i, [_c.POINTER(_vp), _vp, _u64, _u64, _i, _i]),
    "fpb_get_dense": (_i, [_vp, _vp]),
    "fpb_destroy": (None, [_vp]),
    "fpb_rows": (_u64, [_vp]),
    "fpb_cols": (_u64, [_vp]),
    "fpb_nsnps": (_u64, [_vp]),
    "fpb_stream": (_vp, [_vp]),
    "fpb_get_meansd": (_i, [_vp, _vp]),
    "fpb_get_trace": (_i, [_vp, _dp]),
    "fpb_get_bed": (_i, [_vp, _vp]),
    "fpb_perform_op": (_i, [_vp, _vp, _vp]),
    "fpb_perform_op_multi": (_i, [_vp, _vp, _u32, _vp]),
    "fpb_crossprod": (_i, [_vp, _vp, _vp]),
    "fpb_crossprod_multi": (_i, [_vp, _vp, _u32, _vp]),
    "fpb_prod": (_i, [_vp, _vp, _vp]),
    "fpb_prod_multi": (_i, [_vp, _vp, _u32, _vp]),
    "fpb_perform_op_dev": (_i, [_vp, _vp, _vp]),
    "fpb_perform_op_multi_dev": (_i, [_vp, _vp, _u32, _vp]),
    "fpb_crossprod_multi_dev": (_i, [_vp, _vp, _u32, _vp]),
    "fpb_prod_multi_dev": (_i, [_vp, _vp, _u32, _vp]),
    "fpb_sync": (_i, [_vp]),
    "fpb_comm_unique_id": (_i, [_vp]),
    "fpb_comm_init": (_i, [_vp, _vp, _i, _i]),
    "fpb_comm_kind": (_i, [_vp]),
    "fpb_comm_link_local": (_i, [_vp, _i]),
    "fpb_pca": (_i, [_vp, _u32, _u32, _u32, _d, _vp, _vp, _c.POINTER(_u32), _c.POINTER(_u32),
                     _c.POINTER(_u32)]),
    "fpb_pca_block": (_i, [_vp, _u32, _u32, _u32, _d, _vp, _vp, _c.POINTER(_u32), _c.POINTER(_u32)]),
    "fpb_pca_residual": (_i, [_vp, _d, _vp, _u32]),
    "fpb_pca_op_times": (_u32, [_vp, _vp, _u32]),
    "fpb_pca_phase_times": (None, [_vp, _vp]),
    "fpb_time_perform_op": (_i, [_vp, _vp, _vp, _u32, _c.POINTER(_c.c_float), _vp]),
    "fpb_time_perform_op_steps": (_i, [_vp, _vp, _vp, _u32, _vp]),
    "fpb_launch_count": (_u64, [_vp]),
    "fpb_path_info": (_c.c_uint, [_vp]),
    "fpb_device_memory": (_i, [_i, _c.POINTER(_u64), _c.POINTER(_u64)]),
    "fpb_fused_debug": (_i, [_vp, _vp, _u64]),
}

# fpb_path_info bits (include/flashpca_b200.h)
PATH_DENSE, PATH_TENSOR, PATH_TMA, PATH_SINGLE_COPY, PATH_FUSED, PATH_STREAMING = 1, 2, 4, 8, 16, 32

_lib = None


def load() -> ctypes.CDLL:
    """Load the native library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "flashpca_b200: native library %s is missing; run `python -m flashpca_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class FpbError(RuntimeError):
    """Raised where upstream throws std::runtime_error (data.cpp:160,188,287)."""


def check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = load().fpb_last_error(handle)
        raise FpbError(msg.decode() if msg else "flashpca_b200 error")
