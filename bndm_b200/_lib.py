"""ctypes binding of libbndm_b200.so (the C ABI declared in include/bndm_b200.h).

There is deliberately NO CPU / eager fallback here: if the library is missing or the
tensors are not on a CUDA device the call raises.  (The CPU restatement lives in oracle/
and is test infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbndm_b200.so")

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE, ERR_ARCH = 0, -1, -2, -3, -4, -5

SRC_DRAW, SRC_IMAGE = 0, 1
GEMM_AUTO, GEMM_SIMT, FORCE_DENSE, GEMM_GEMV, GEMM_TC = 0, 16, 32, 64, 128

# name -> (restype, argtypes); mirrors include/bndm_b200.h one to one
_P = C.c_void_p
SIGNATURES = {
    "bndm_version": (C.c_int, []),
    "bndm_last_error": (C.c_char_p, []),
    "bndm_device_is_sm100": (C.c_int, []),
    "bndm_prepare_L": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "bndm_reserve_columns": (C.c_int, [_P, C.c_int, _P]),
    "bndm_L_is_lower_triangular": (C.c_int, [_P]),
    "bndm_workspace_bytes": (C.c_int64, [_P]),
    "bndm_free_L": (C.c_int, [_P]),
    "bndm_profile_enable": (C.c_int, [_P, C.c_int]),
    "bndm_profile_last_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "bndm_get_noise_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_uint, _P]),
    "bndm_get_noise_shard_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, _P]),
    "bndm_get_noise_train_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_uint, _P]),
    "bndm_white128_reinterpret_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "bndm_white128_reinterpret_shard_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "bndm_iadb_step_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "bndm_iadb_step_sched_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "bndm_iadb_step_sched_dnhwc_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "bndm_ddim_step_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, _P]),
    "bndm_debug_set_trace": (C.c_int, [_P, _P]),
    "bndm_debug_set_policy": (C.c_int, [C.c_int, C.c_int]),
    "bndm_debug_streamk_check": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "bndm_debug_streamk_check_sub": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "bndm_debug_gemv_schedule_check": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bndm_groupnorm_nhwc_f32": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_float, C.c_int, _P]),
    "bndm_linear_tc_f32": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "bndm_shortcut_residual_tf32": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int64, C.c_int, _P]),
    "bndm_conv_in3x3_nhwc_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "bndm_upsample2x_nhwc_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "bndm_attention_small_f32": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "bndm_add_bias_nhwc_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, _P]),
    "bndm_snapshot_uint8_hwc": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "bndm_to_uint8_nhwc": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
}

_lock = threading.Lock()
_lib = None


class BndmError(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise BndmError(
                    f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; "
                    f"g.build()'`). bndm_b200 has no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.bndm_version() != 2:
                raise BndmError("libbndm_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def check(rc: int, what: str):
    """Maps C-ABI error codes onto the exceptions the reference raises at the same sites."""
    if rc == OK:
        return
    msg = load().bndm_last_error().decode("utf-8", "replace")
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == ERR_ARG:
        raise ValueError(f"{what}: {msg}")
    raise BndmError(f"{what}: {msg} (code {rc})")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda_f32(t, name):
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise BndmError(f"{name} is on {t.device}: bndm_b200 only runs on CUDA tensors (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (the reference casts with .float()), got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def current_stream(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
