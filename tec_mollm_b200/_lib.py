"""ctypes binding of ``libtecgat.so`` (C ABI declared in ``include/tecgat.h``).

There is NO fallback: if the library is missing or a call fails, a ``RuntimeError`` is raised.
The library is built in-tree by ``build.py`` (``python -m tec_mollm_b200.build`` or
``__graft_entry__.build()``) with ``nvcc -gencode arch=compute_100a,code=sm_100a``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TECGAT_LIB") or os.path.join(_HERE, "lib", "libtecgat.so")  # TECGAT_LIB: A/B builds (tools/tune_edge_bwd.py)

F32, BF16 = 0, 1
MODE_SHARED, MODE_LITERAL = 0, 1
PROJ_TC, PROJ_FFMA = 0, 1
ABI_VERSION = 4

_lock = threading.Lock()
_lib = None

_vp, _i32, _i64, _u64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double

# name -> (restype, argtypes); every int-returning entry is error-checked by ``call``
_SIGNATURES = {
    "tecgat_abi_version": (C.c_int, []),
    "tecgat_last_error": (C.c_char_p, []),
    "tecgat_plan_create": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, C.POINTER(_vp)]),
    "tecgat_plan_destroy": (C.c_int, [_vp]),
    "tecgat_plan_info": (C.c_int, [_vp, C.POINTER(_i64)]),
    "tecgat_plan_export": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tecgat_project_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "tecgat_project_bwd_workspace": (_i64, [_i64, _i32, _i32, _i32]),
    "tecgat_project_bwd": (C.c_int, [_vp] * 11 + [_i64, _i32, _i32, _i32, _i32, _vp]),
    "tecgat_project_bwd_acc_supported": (C.c_int, [_i32, _i32]),
    "tecgat_project_bwd_acc": (C.c_int, [_vp] * 11 + [_i64, _i32, _i32, _i32, _vp]),
    "tecgat_edge_fwd": (C.c_int, [_vp] * 7 + [_i32, _i32, _i32, _f32, _f32, _u64, _i32, _i32, _vp]),
    "tecgat_edge_bwd_workspace": (_i64, [_vp, _i32, _i32, _i32]),
    "tecgat_edge_bwd": (C.c_int, [_vp] * 13 + [_i32, _i32, _i32, _f32, _f32, _u64, _i32, _i32, _vp]),
    "tecgat_forward": (C.c_int, [_vp] * 12 + [_i32, _i32, _i32, _i32, _f32, _f32, _u64, _vp, _i32, _i32, _i32, _vp]),
    "tecgat_forward_into": (C.c_int, [_vp] * 12 + [_i32, _i32, _i32, _i32, _f32, _f32, _u64, _vp, _i32, _i32, _i32, _vp, _i64, _vp]),
    "tecgat_backward_workspace": (_i64, [_vp, _i32, _i32, _i32, _i32, _i32]),
    "tecgat_backward_fused_supported": (C.c_int, [_i32, _i32, _i32]),
    "tecgat_backward": (C.c_int, [_vp] * 14 + [_i32] + [_vp] * 6 + [_i32, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _u64, _vp,
                                  _i32, _i32, _i32, _vp]),
    "tecgat_seed_advance": (C.c_int, [_vp, _vp, _vp]),
    "tecgat_launch_count": (_i64, []),
    "tecgat_phase_timing": (C.c_int, [_i32]),
    "tecgat_phase_times": (C.c_int, [_vp]),
    "tecgat_embed_fwd": (C.c_int, [_vp] * 8 + [_i32] * 8 + [_vp]),
    "tecgat_embed_bwd_workspace": (_i64, [_i32, _i32, _i32]),
    "tecgat_embed_bwd": (C.c_int, [_vp] * 8 + [_i32] * 9 + [_vp]),
    "tecgat_gn_gelu_fwd": (C.c_int, [_vp] * 6 + [_i64, _i32, _i32, _i32, _i32, _f32, _i32, _i32, _vp]),
    "tecgat_gn_gelu_bwd_workspace": (_i64, [_i64, _i32, _i32]),
    "tecgat_gn_gelu_bwd": (C.c_int, [_vp] * 10 + [_i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tecgat_residual_permute_fwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tecgat_residual_permute_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tecgat_dropout_mask_host": (C.c_int, [_u64, _i64, _i64, _i32, _f32, _i64, _vp]),
    "tecgraph_distance_rows": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _f64, _vp, _vp]),
    "tecgraph_edges_count": (C.c_int, [_vp, _vp, _i64, _f64, _f64, _vp, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64)]),
    "tecgraph_edges_fill": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tecgraph_ctx_destroy": (C.c_int, [_vp]),
    "tecgraph_ctx_stats": (C.c_int, [_vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"tec_mollm_b200: CUDA library {LIB_PATH} is missing -- build it with "
                "`python -m tec_mollm_b200.build` (needs nvcc). There is no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI drifted
            fn.restype = res
            fn.argtypes = args
        got = handle.tecgat_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"tec_mollm_b200: libtecgat ABI {got} != expected {ABI_VERSION}; rebuild the library")
        _lib = handle
    return _lib


def call(name: str, *args):
    """Call an int-returning entry point and turn a non-zero code into ``RuntimeError``."""
    handle = lib()
    rc = getattr(handle, name)(*args)
    if rc != 0:
        msg = handle.tecgat_last_error()
        raise RuntimeError(f"{name} failed (code {rc}): {msg.decode(errors='replace') if msg else '?'}")
    return rc
