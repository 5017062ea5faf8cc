"""ctypes binding of libsam2b200.so (the C ABI declared in include/sam2_b200.h).

There is NO fallback: if the library is missing or the device is not sm_100, every op raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsam2b200.so")
_lib = None


class Sam2B200Error(RuntimeError):
    pass


_SIGNATURES = {
    "sam2b200_version": (c_int, []),
    "sam2b200_last_error": (c_char_p, []),
    "sam2b200_launch_count": (c_longlong, []),
    "sam2b200_debug_set_timeline": (c_longlong, [c_void_p, c_longlong]),
    "sam2b200_debug_set_variant": (c_int, [c_int, c_int]),
    "sam2b200_check_device": (c_int, [c_int]),
    "sam2b200_rope_apply": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p]),
    "sam2b200_attn_default_nsplit": (c_int, [c_int, c_int, c_int]),
    "sam2b200_attn_fwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sam2b200_attn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "sam2b200_attn_bwd": (c_int, [c_void_p] * 11 + [c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                                   c_int, c_int, c_int, c_float, c_void_p]),
    "sam2b200_attn_fwd_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_int, c_int, c_int, c_float, c_int, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_attn_bwd_ex": (c_int, [c_void_p] * 11 + [c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                                      c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int,
                                                      c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_ln_fwd": (c_int, [c_void_p] * 9 + [c_longlong, c_float, c_int, c_int, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_ln_bwd_workspace_bytes": (c_size_t, [c_longlong]),
    "sam2b200_ln_bwd": (c_int, [c_void_p] * 13 + [c_longlong, c_int, c_int, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_ln_bwd_stages": (c_int, [c_void_p] * 13 + [c_longlong, c_int, c_int, c_float, c_void_p, c_uint, c_int, c_void_p]),
    "sam2b200_colsum_workspace_bytes": (c_size_t, [c_longlong, c_int]),
    "sam2b200_colsum": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int,
                                c_longlong, c_float, c_void_p]),
    "sam2b200_permute_rows": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                      c_void_p]),
    "sam2b200_dropout_inplace": (c_int, [c_void_p, c_longlong, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_dropout_mask": (c_int, [c_void_p, c_longlong, c_longlong, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_mask_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_longlong]),
    "sam2b200_mask_loss_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float,
                                       c_float, c_float, c_int, c_int, c_void_p]),
    "sam2b200_mask_loss_bwd": (c_int, [c_void_p] * 9 + [c_int, c_int, c_longlong, c_int, c_float,
                                                         c_float, c_float, c_int, c_int, c_void_p]),
    "sam2b200_mask_loss_fwd_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float,
                                              c_float, c_float, c_int, c_int, c_void_p]),
    "sam2b200_mask_loss_bwd_frames": (c_int, [c_void_p] * 9 + [c_int, c_int, c_longlong, c_int, c_float,
                                                                c_float, c_float, c_int, c_int, c_void_p]),
    "sam2b200_mask_loss_bwd_coef": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_float,
                                            c_float, c_float, c_void_p]),
    "sam2b200_attn_fwd_v64": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_float, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_attn_fwd_proj": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_float, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_attn_fwd_v64_proj": (c_int, [c_void_p] * 11 + [c_int, c_int, c_int, c_float, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_attn_bwd_v64": (c_int, [c_void_p] * 8 + [c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                      c_void_p, c_void_p, c_int, c_void_p, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_merged_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sam2b200_merged_loss_fwd": (c_int, [c_void_p] * 12 + [c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_void_p]),
    "sam2b200_merged_loss_bwd": (c_int, [c_void_p] * 13 + [c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_void_p]),
    "sam2b200_proj_rope": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                                   c_void_p, c_int, c_int, c_int, c_void_p]),
    "sam2b200_ln_proj": (c_int, [c_void_p] * 8 + [c_longlong, c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_uint, c_float, c_void_p, c_uint, c_void_p]),
    "sam2b200_ln_gelu_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_int, c_void_p]),
    "sam2b200_ln_gelu_bwd": (c_int, [c_void_p] * 7 + [c_longlong, c_int, c_float, c_int, c_void_p]),
    "sam2b200_dwconv7": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sam2b200_dwconv7_bwd_w_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sam2b200_dwconv7_bwd_w": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sam2b200_wgrad": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "sam2b200_gemm": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_longlong, c_int, c_int, c_void_p, c_void_p,
                              c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sam2b200_gemm_debug_timeline": (c_longlong, [c_void_p, c_longlong]),
    "sam2b200_gemm_plan": (c_int, [c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sam2b200_gemm_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_longlong, c_int,
                                 c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_uint, c_void_p, c_void_p, c_void_p]),
    "sam2b200_fold_grads": (c_int, [c_void_p] * 11),
    "sam2b200_wgrad2": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                c_longlong, c_void_p, c_void_p]),
    "sam2b200_mlp_dh": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p]),
    "sam2b200_bank_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sam2b200_bank_gather_packed": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                            c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load(build_if_missing: bool = False):
    """Load (once) and return the ctypes handle.  Raises Sam2B200Error if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build
            _build.build()
        else:
            raise Sam2B200Error(
                f"{LIB_PATH} not found: build it with `python -m sam2_video_training_b200.build` "
                "(there is no CPU / PyTorch fallback for this path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().sam2b200_last_error()
        raise Sam2B200Error(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr_array(ptrs):
    """Host array of device pointers (const T* const*)."""
    arr = (c_void_p * len(ptrs))(*ptrs)
    return arr
