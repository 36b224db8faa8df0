"""ctypes binding of libhwgat_b200.so (include/hwgat_b200.h).

There is no fallback: if the library is missing, or a call returns a non-zero
status, this raises.  Nothing here computes anything on the host.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HWGAT_B200_LIB") or os.path.join(HERE, "lib", "libhwgat_b200.so")   # override: A/B builds

F32, BF16 = 0, 1
LAYOUT_BFKD, LAYOUT_WINDOWS = 0, 1
ABI_VERSION = 23
FP32_MODES = {"ffma": 0, "x3": 1}      # HWGAT_FP32_FFMA / HWGAT_FP32_X3 (include/hwgat_b200.h)
FP32_DEFAULT = "x3"        # the package default; the C library itself starts in FFMA mode

# name -> (restype, argtypes); must list every symbol of include/hwgat_b200.h
SIGNATURES = {
    "hwgat_version": (c_int, []),
    "hwgat_error_string": (c_char_p, [c_int]),
    "hwgat_attn2_f32_workspace_bytes": (c_size_t, [c_int] * 5),
    "hwgat_attn2_fwd_f32": (c_int, [c_void_p] * 4 + [c_float] + [c_void_p] * 2 + [c_int] * 9 + [c_void_p]),
    "hwgat_attn2_bwd_f32": (c_int, [c_void_p] * 5 + [c_float] + [c_void_p] * 4 + [c_size_t] + [c_int] * 9 + [c_void_p]),
    "hwgat_ln_wpool_fwd": (c_int, [c_void_p] * 8 + [c_size_t, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p]),
    "hwgat_ln_wpool_bwd": (c_int, [c_void_p] * 10 + [c_int] * 5 + [c_void_p]),
    "hwgat_band_attn_workspace_bytes": (c_size_t, [c_int] * 6),
    "hwgat_band_attn_fwd": (c_int, [c_int] + [c_void_p] * 7 + [c_int] * 7 + [c_void_p]),
    "hwgat_band_attn_bwd": (c_int, [c_int] + [c_void_p] * 11 + [c_size_t] + [c_int] * 7 + [c_void_p]),
    "hwgat_launch_count": (c_ulonglong, []),
    "hwgat_set_deterministic": (c_int, [c_int]),
    "hwgat_set_fp32_mode": (c_int, [c_int]),
    "hwgat_adjacency_build": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hwgat_mask_build": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hwgat_mask_pack": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "hwgat_attn_workspace_bytes": (c_size_t, [c_int] * 7),
    "hwgat_attn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_size_t]
                       + [c_int] * 10 + [c_void_p]),
    "hwgat_attn_bwd_f32_kept": (c_int, [c_void_p] * 5 + [c_float] + [c_void_p] * 4 + [c_size_t] + [c_int] * 9 + [c_void_p]),
    "hwgat_attn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_size_t] + [c_int] * 10 + [c_void_p]),
    "hwgat_attn2_workspace_bytes": (c_size_t, [c_int] * 7),
    "hwgat_attn2_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_size_t]
                        + [c_int] * 9 + [c_float, c_ulonglong, c_ulonglong, c_void_p]),
    "hwgat_attn2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t] + [c_int] * 10 + [c_float, c_ulonglong, c_ulonglong,
                                                                                c_void_p]),
    "hwgat_attn_fwd_keep": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                    c_size_t] + [c_int] * 9 + [c_void_p]),
    "hwgat_ln_fwd": (c_int, [c_void_p] * 6 + [c_longlong, c_int, c_float, c_void_p]),
    "hwgat_ln_fwd_f32": (c_int, [c_void_p] * 6 + [c_longlong, c_int, c_float, c_void_p]),
    "hwgat_ln_bwd": (c_int, [c_void_p] * 9 + [c_longlong, c_int, c_void_p]),
    "hwgat_ln_bwd_f32": (c_int, [c_void_p] * 9 + [c_longlong, c_int, c_void_p]),
    "hwgat_bda_ln_fwd": (c_int, [c_void_p] * 9 + [c_longlong, c_int, c_float, c_float, c_ulonglong, c_ulonglong,
                                                  c_void_p]),
    "hwgat_bda_ln_fwd_f32": (c_int, [c_void_p] * 9 + [c_longlong, c_int, c_float, c_float, c_ulonglong, c_ulonglong,
                                                  c_void_p]),
    "hwgat_bda_ln_bwd": (c_int, [c_void_p] * 11 + [c_longlong, c_int, c_float, c_ulonglong, c_ulonglong, c_void_p]),
    "hwgat_bda_ln_bwd_f32": (c_int, [c_void_p] * 11 + [c_longlong, c_int, c_float, c_ulonglong, c_ulonglong, c_void_p]),
    "hwgat_bda_merge_fwd": (c_int, [c_void_p] * 4 + [c_longlong, c_int, c_int, c_int, c_float, c_ulonglong, c_ulonglong,
                                                      c_void_p]),
    "hwgat_bda_merge_fwd_f32": (c_int, [c_void_p] * 4 + [c_longlong, c_int, c_int, c_int, c_float, c_ulonglong, c_ulonglong,
                                                      c_void_p]),
    "hwgat_ln_bwd_unmerge": (c_int, [c_void_p] * 9 + [c_longlong, c_int, c_int, c_int, c_void_p]),
    "hwgat_ln_bwd_unmerge_f32": (c_int, [c_void_p] * 9 + [c_longlong, c_int, c_int, c_int, c_void_p]),
    "hwgat_linear_x3_supported": (c_int, [c_longlong, c_int, c_int]),
    "hwgat_linear_x3_fwd": (c_int, [c_void_p] * 5 + [c_longlong, c_int, c_int, c_void_p]),
    "hwgat_linear_x3_bwd": (c_int, [c_void_p] * 6 + [c_longlong, c_int, c_int, c_void_p]),
    "hwgat_linear_f32_fwd": (c_int, [c_void_p] * 4 + [c_int, c_int, c_int, c_void_p]),
    "hwgat_linear_f32_bwd": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_void_p]),
    "hwgat_smooth_ce_fwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_float, c_void_p]),
    "hwgat_smooth_ce_bwd": (c_int, [c_void_p] * 5 + [c_int, c_int, c_float, c_void_p]),
    "hwgat_bias_gelu_dropout_fwd": (c_int, [c_void_p] * 3 + [c_longlong, c_int, c_float, c_ulonglong, c_ulonglong,
                                                             c_void_p]),
    "hwgat_bias_gelu_dropout_fwd_f32": (c_int, [c_void_p] * 3 + [c_longlong, c_int, c_float, c_ulonglong, c_ulonglong,
                                                             c_void_p]),
    "hwgat_bias_gelu_dropout_bwd": (c_int, [c_void_p] * 5 + [c_longlong, c_int, c_float, c_ulonglong, c_ulonglong,
                                                             c_void_p]),
    "hwgat_bias_gelu_dropout_bwd_f32": (c_int, [c_void_p] * 5 + [c_longlong, c_int, c_float, c_ulonglong, c_ulonglong,
                                                             c_void_p]),
    "hwgat_embed_fwd": (c_int, [c_void_p] * 4 + [c_longlong, c_int, c_int, c_int, c_int, c_float, c_ulonglong,
                                                 c_ulonglong, c_void_p]),
    "hwgat_ln_pool_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "hwgat_ln_pool_fwd": (c_int, [c_void_p] * 7 + [c_size_t, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p]),
    "hwgat_ln_pool_bwd": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hwgat_ffn_fused_supported": (c_int, [c_longlong, c_int, c_int]),
    "hwgat_ffn_fwd": (c_int, [c_void_p] * 7 + [c_longlong, c_int, c_int, c_float, c_ulonglong, c_ulonglong,
                               c_void_p]),
    "hwgat_ffn_bwd_workspace_bytes": (c_size_t, [c_longlong, c_int, c_int]),
    "hwgat_ffn_bwd": (c_int, [c_void_p] * 11 + [c_size_t, c_longlong, c_int, c_int, c_void_p]),
    "hwgat_proj_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_void_p]),
    "hwgat_proj_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "hwgat_proj_bwd": (c_int, [c_void_p] * 6 + [c_size_t, c_longlong, c_int, c_int, c_void_p]),
    "hwgat_adamw_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double,
                                  c_double, c_double, c_longlong, c_float, c_void_p]),
    "hwgat_debug_gemm_nt": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_void_p]),
    "hwgat_debug_gemm_tn": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p]),
    "hwgat_debug_set_gemm_pair": (c_int, [c_int]),
    "hwgat_merge_fwd": (c_int, [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    "hwgat_merge_bwd": (c_int, [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
}

_lib = None


class HwgatError(RuntimeError):
    pass


# Entry points that take stream-ordered scratch inside the library (the bf16 planes of the x3 fp32 GEMMs): the pool
# cannot grow while PyTorch's caching allocator sits on the free memory, so a HWGAT_ERR_WORKSPACE from them is retried
# once after the cache has been released (every such call fails before it has written anything, or is idempotent).
_SCRATCH_CALLS = ("hwgat_attn_fwd", "hwgat_attn_bwd", "hwgat_attn_bwd_f32_kept", "hwgat_attn2_fwd_f32", "hwgat_attn2_bwd_f32",
                  "hwgat_band_attn_fwd", "hwgat_band_attn_bwd", "hwgat_linear_f32_fwd", "hwgat_linear_f32_bwd",
                  "hwgat_linear_x3_fwd", "hwgat_linear_x3_bwd")
ERR_WORKSPACE = 1003


def _with_scratch_retry(fn):
    def call(*args):
        status = fn(*args)
        if status == ERR_WORKSPACE:
            try:
                import torch
                torch.cuda.synchronize()
                torch.cuda.empty_cache()
            except Exception:      # noqa: BLE001 - no torch / no device: report the original status
                return status
            status = fn(*args)
        return status
    call.__name__ = getattr(fn, "__name__", "hwgat_call")
    return call


def load() -> ctypes.CDLL:
    """Load the library once.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HwgatError(
            f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
            "(run `python -m sl_hwgat_b200.build`). There is no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.hwgat_version()
    if got != ABI_VERSION:
        raise HwgatError(f"libhwgat_b200 ABI {got}, binding expects {ABI_VERSION}: rebuild the library")
    if os.environ.get("HWGAT_DETERMINISTIC", "0") not in ("", "0"):   # see ops.set_deterministic
        lib.hwgat_set_deterministic(1)
    for name in _SCRATCH_CALLS:
        setattr(lib, name, _with_scratch_retry(getattr(lib, name)))
    mode = os.environ.get("HWGAT_FP32", FP32_DEFAULT).lower()          # see ops.set_fp32_mode
    if mode not in FP32_MODES:
        raise HwgatError(f"HWGAT_FP32={mode!r}: expected one of {sorted(FP32_MODES)}")
    lib.hwgat_set_fp32_mode(FP32_MODES[mode])
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().hwgat_error_string(status)
        raise HwgatError(f"{what} failed with status {status}: {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().hwgat_launch_count())
