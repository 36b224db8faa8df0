"""Host side of the hot path: torch.autograd.Functions over the C ABI.

Every function here hands device pointers and the current CUDA stream to
libhwgat_b200.so (include/hwgat_b200.h).  Tensors are allocated by PyTorch
(device memory + caching allocator = plumbing); all arithmetic of the path runs
in the hand-written kernels.  A tensor that is not on a CUDA device is an error:
there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import BF16, F32, LAYOUT_BFKD, LAYOUT_WINDOWS, check

import os

WINDOW = 16       # the reference's keypoints per window (model_params.py:254)
WINDOWS = (16, 32, 64)   # window sizes the bf16 kernels cover: N = 2 W = 32, 64, 128 tokens per window
TEMPORAL_PATCH = 2  # frames per window (model_params.py:250)
HEAD_DIM = 64

# Which bf16 attention kernels run for the reference window (W = 16):
#   "fused" : K2 / K3 (attn_tc.cu) - QKV projection inside the attention kernel, 32 x 32 windows on mma.sync
#   "tc2"   : K2b / K3b (attn_core_tc2.cu) - projection GEMM + attention core with every product on tcgen05
# W = 32 / 64 always take "tc2" (the only kernels built for them).  HWGAT_ATTN_IMPL overrides for A/B runs.
#   "hybrid": K2 forward that also keeps its q, k, v rows + K3b backward on them (no recompute, no projection GEMM)
ATTN_IMPL = os.environ.get("HWGAT_ATTN_IMPL", "fused")
# with "fused": widths from which the hybrid is used (0 = never); HWGAT_HYBRID_MIN_D overrides for A/B runs
HYBRID_MIN_D = int(os.environ.get("HWGAT_HYBRID_MIN_D", "0"))


def _need_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.HwgatError(
                "sl_hwgat_b200 runs on sm_100a CUDA kernels only; got a tensor on "
                f"{t.device}. There is no CPU fallback.")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise _lib.HwgatError(f"unsupported dtype {t.dtype}: the kernels take float32 or bfloat16")


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _require_binary(t: Optional[torch.Tensor], what: str) -> None:
    """The kernels hold the two multiplicative masks of MSA.forward (HWGATE.py:102-108) as ONE bit per
    (query, key).  That is exact for 0/1 masks - what the reference ships - and wrong for weighted ones, which
    the reference would multiply into the logits: refuse those instead of silently binarising them.  Runs once
    per mask (the packed bits are cached by the callers)."""
    if t is not None and not bool(((t == 0) | (t == 1)).all().item()):
        raise NotImplementedError(
            f"{what} has entries other than 0 and 1: the packed-bitmask kernels only represent 0/1 "
            "multiplicative masks (the reference's skeleton adjacency and shifted-window mask)")


_CAST_CACHE: dict = {}


def set_deterministic(on: bool = True) -> bool:
    """Bit-reproducible PARAMETER gradients (outputs and input gradients always are): every token split of the
    weight-gradient GEMMs and every CTA of the bias / LayerNorm column sums writes its own partial, added in index
    order by a finish kernel (hwgat_set_deterministic, include/hwgat_b200.h; also env HWGAT_DETERMINISTIC=1).  The
    default keeps the fp32-atomic sums, whose order varies run to run at the 1e-6 level.  Process-wide; returns the
    previous setting.  The reference offers nothing comparable (its cuBLAS / eager backward is not run-to-run exact either)."""
    return bool(_lib.load().hwgat_set_deterministic(1 if on else 0))


def set_fp32_mode(mode: str) -> str:
    """How fp32 tensors (no autocast: the reference's own loop, utils.py:102) are multiplied.  "ffma": true-fp32 FFMA
    GEMMs, the 1e-5 parity mode.  "x3": every Linear of the blocks and the QKV projection of the attention run on
    tcgen05 as six bf16 products of hi / mid / lo planes with fp32 accumulation (gemm_x3.cu; ~2e-7 against fp64,
    several times faster).  Process-wide (hwgat_set_fp32_mode; also env HWGAT_FP32); returns the previous mode."""
    if mode not in _lib.FP32_MODES:
        raise ValueError(f"fp32 mode {mode!r}: expected one of {sorted(_lib.FP32_MODES)}")
    prev = _lib.load().hwgat_set_fp32_mode(_lib.FP32_MODES[mode])
    return {v: k for k, v in _lib.FP32_MODES.items()}[prev]


def linear_x3_active(n: int, d_in: int, d_out: int) -> bool:
    """True when linear_f32 of this shape runs on the tcgen05 x3 GEMM: mode "x3", not in deterministic mode, and a shape
    gemm_x3.cu takes (n, d_in, d_out multiples of 128: hwgat_linear_x3_supported)."""
    lib = _lib.load()
    return (lib.hwgat_set_fp32_mode(-1) == _lib.FP32_MODES["x3"] and not lib.hwgat_set_deterministic(-1)
            and bool(lib.hwgat_linear_x3_supported(n, d_in, d_out)))


def fp32_mode() -> str:
    return {v: k for k, v in _lib.FP32_MODES.items()}[_lib.load().hwgat_set_fp32_mode(-1)]


def invalidate_cast_cache() -> None:
    """Forget every cached low-precision parameter copy.  Needed after parameters were written through raw pointers,
    which does not bump torch's `_version` counter: sl_hwgat_b200.optim.AdamW calls this after every step."""
    _CAST_CACHE.clear()


def cast_cached(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Detached contiguous copy of a parameter in `dtype`, reused until the parameter is modified in place
    (`_version`) or re-allocated (`data_ptr`): inference re-casts nothing, training casts once per step
    instead of once per forward and once per backward."""
    src = t.detach()
    if src.dtype == dtype and src.is_contiguous():
        return src
    if src.is_cuda and torch.cuda.is_current_stream_capturing():
        return src.to(dtype).contiguous()      # inside a CUDA graph the cast is part of the graph (replays see updates)
    key = (id(t), dtype)
    tag = (src.data_ptr(), t._version, src.device, tuple(src.shape))
    hit = _CAST_CACHE.get(key)
    if hit is not None and hit[0] == tag and hit[2]() is t:
        return hit[1]
    out = src.to(dtype).contiguous()
    import weakref
    try:
        ref = weakref.ref(t, lambda _r, k=key: _CAST_CACHE.pop(k, None))
    except TypeError:
        return out
    _CAST_CACHE[key] = (tag, out, ref)
    return out


# --------------------------------------------------------------------------
# K1: adjacency and packed masks
# --------------------------------------------------------------------------

def adjacency_build(edges: Sequence[Sequence[Sequence[int]]], window: int, temporal_patch: int,
                    device) -> torch.Tensor:
    """(nW, TP*W, TP*W) float32 adjacency on `device`.
    Replaces HWGATEParams.get_adj_mat (model_params.py:373-400)."""
    lib = _lib.load()
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.HwgatError("adjacency_build needs a CUDA device (no CPU fallback)")
    nW = len(edges)
    n_edges = len(edges[0]) if nW else 0
    if any(len(e) != n_edges for e in edges):
        raise ValueError("every window must list the same number of edges")
    e = torch.tensor(edges, dtype=torch.int32).reshape(nW, n_edges, 2).to(device)
    N = window * temporal_patch
    adj = torch.empty((nW, N, N), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib.hwgat_adjacency_build(e.data_ptr(), n_edges, nW, window, temporal_patch, adj.data_ptr(),
                                        _stream()), "hwgat_adjacency_build")
    return adj


def mask_build(adj: torch.Tensor, frames: int, shift: int, window: int = WINDOW,
               temporal_patch: int = TEMPORAL_PATCH) -> torch.Tensor:
    """Packed mask (frames/TP * nW, N, N/32) uint32 (stored as int32) of one block:
    adjacency AND shifted-window mask.  Replaces HWGATE.py:169-187, 309, 102-108."""
    lib = _lib.load()
    _need_cuda(adj)
    adj = adj.contiguous().float()
    _require_binary(adj, "adj_mat")
    nW, N = adj.shape[0], adj.shape[1]
    bits = torch.empty((frames // temporal_patch * nW, N, N // 32), dtype=torch.int32, device=adj.device)
    with torch.cuda.device(adj.device):
        check(lib.hwgat_mask_build(adj.data_ptr(), nW, window, temporal_patch, frames, shift, bits.data_ptr(),
                                   _stream()), "hwgat_mask_build")
    return bits


def mask_pack(adj: Optional[torch.Tensor], mask: Optional[torch.Tensor], n_windows: int, N: int,
              device) -> torch.Tensor:
    """Pack caller-provided float masks: bits = (adj != 0) & (mask != 0)
    (the two multiplies of MSA.forward, HWGATE.py:102-108)."""
    lib = _lib.load()
    _need_cuda(adj, mask)
    adj_c = adj.contiguous().float() if adj is not None else None
    mask_c = mask.contiguous().float() if mask is not None else None
    _require_binary(adj_c, "adj_mat")
    _require_binary(mask_c, "attn_mask")
    bits = torch.empty((n_windows, N, N // 32), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib.hwgat_mask_pack(_ptr(adj_c), 0 if adj_c is None else adj_c.shape[0], _ptr(mask_c), n_windows, N,
                                  bits.data_ptr(), _stream()), "hwgat_mask_pack")
    return bits


# --------------------------------------------------------------------------
# K2 / K3: fused windowed graph attention
# --------------------------------------------------------------------------

class _WindowGraphAttention(torch.autograd.Function):
    """out = PV(softmax(mask(QK^T))) of every window, from the normalised
    residual stream; forward saves only its inputs (K3 recomputes)."""

    @staticmethod
    def forward(ctx, xn, w_qkv, b_qkv, bits, threshold, heads, shift, layout, frames, kps, hybrid=False):
        lib = _lib.load()
        _need_cuda(xn, w_qkv, b_qkv, bits)
        code = _dtype_code(xn)
        xn_c = xn.contiguous()
        d = xn_c.shape[-1]
        n_tok = xn_c.numel() // d
        if frames * kps == 0 or n_tok % (frames * kps) != 0:
            raise ValueError(f"token count {n_tok} is not a multiple of frames*keypoints = {frames * kps}")
        B = n_tok // (frames * kps)
        w_c = cast_cached(w_qkv, xn_c.dtype)
        b_c = cast_cached(b_qkv, torch.float32)
        out = torch.empty_like(xn_c)
        ws_bytes = lib.hwgat_attn_workspace_bytes(B, frames, kps, d, heads, code, 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xn_c.device) if ws_bytes else None
        # hybrid (bf16, gradients wanted): K2 also keeps the q, k, v rows it formed, and the backward is K3b's
        # tcgen05 core on them (no QKV recompute) followed by the same d_xn / d_w GEMMs
        hybrid = bool(hybrid) and code == BF16 and any(ctx.needs_input_grad[:3]) and B > 0
        qkv = torch.empty((n_tok, 3 * d), dtype=torch.bfloat16, device=xn_c.device) if hybrid else None
        with torch.cuda.device(xn_c.device):
            if hybrid:
                check(lib.hwgat_attn_fwd_keep(xn_c.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), bits.data_ptr(),
                                              float(threshold), out.data_ptr(), qkv.data_ptr(), _ptr(ws), ws_bytes, B,
                                              frames, kps, d, heads, WINDOW, TEMPORAL_PATCH, shift, layout, _stream()),
                      "hwgat_attn_fwd_keep")
            else:
                check(lib.hwgat_attn_fwd(xn_c.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), bits.data_ptr(),
                                         float(threshold), out.data_ptr(), _ptr(ws), ws_bytes, B, frames, kps, d, heads,
                                         WINDOW, TEMPORAL_PATCH, shift, layout, code, _stream()), "hwgat_attn_fwd")
        # fp32: the forward's workspace IS the projected qkv (n, 3d): kept for the backward instead of re-projecting
        # (12 d bytes per token; one x3 GEMM and its splits less per block)
        keep_ws = ws if (code == F32 and B > 0 and any(ctx.needs_input_grad[:3])) else None
        ctx.save_for_backward(xn_c, w_c, b_c, bits, qkv, keep_ws)
        ctx.meta = (float(threshold), heads, shift, layout, frames, kps, B, d, code, w_qkv.dtype, b_qkv.dtype)
        return out.view_as(xn)

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        xn_c, w_c, b_c, bits, qkv, kept = ctx.saved_tensors
        threshold, heads, shift, layout, frames, kps, B, d, code, w_dtype, b_dtype = ctx.meta
        g = d_out.to(xn_c.dtype).contiguous()
        d_xn = torch.empty_like(xn_c)
        d_w = torch.empty((3 * d, d), dtype=torch.float32, device=xn_c.device)
        d_b = torch.empty((3 * d,), dtype=torch.float32, device=xn_c.device)
        if kept is not None:      # fp32 with the forward's qkv kept
            ws = torch.empty(kept.numel(), dtype=torch.uint8, device=xn_c.device)
            with torch.cuda.device(xn_c.device):
                check(lib.hwgat_attn_bwd_f32_kept(g.data_ptr(), xn_c.data_ptr(), w_c.data_ptr(), kept.data_ptr(),
                                                  bits.data_ptr(), threshold, d_xn.data_ptr(), d_w.data_ptr(),
                                                  d_b.data_ptr(), ws.data_ptr(), ws.numel(), B, frames, kps, d, heads,
                                                  WINDOW, TEMPORAL_PATCH, shift, layout, _stream()),
                      "hwgat_attn_bwd_f32_kept")
            return (d_xn.view_as(d_out), d_w.to(w_dtype), d_b.to(b_dtype)) + (None,) * 8
        if qkv is not None:       # hybrid: K3b on the q, k, v kept by K2 (q, k column-permuted: qk_perm = 1)
            ws_bytes = lib.hwgat_attn2_workspace_bytes(B, frames, kps, d, heads, 1, 1)
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xn_c.device)
            with torch.cuda.device(xn_c.device):
                check(lib.hwgat_attn2_bwd(g.data_ptr(), xn_c.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), qkv.data_ptr(),
                                          bits.data_ptr(), threshold, d_xn.data_ptr(), d_w.data_ptr(), d_b.data_ptr(),
                                          ws.data_ptr(), ws.numel(), B, frames, kps, d, heads, WINDOW, TEMPORAL_PATCH,
                                          shift, layout, 1, 0.0, 0, 0, _stream()), "hwgat_attn2_bwd")
            return (d_xn.view_as(d_out), d_w.to(w_dtype), d_b.to(b_dtype)) + (None,) * 8
        ws_bytes = lib.hwgat_attn_workspace_bytes(B, frames, kps, d, heads, code, 1)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xn_c.device)
        with torch.cuda.device(xn_c.device):
            check(lib.hwgat_attn_bwd(g.data_ptr(), xn_c.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), bits.data_ptr(),
                                     threshold, d_xn.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), ws.data_ptr(),
                                     ws.numel(), B, frames, kps, d, heads, WINDOW, TEMPORAL_PATCH, shift, layout,
                                     code, _stream()), "hwgat_attn_bwd")
        return (d_xn.view_as(d_out), d_w.to(w_dtype), d_b.to(b_dtype)) + (None,) * 8


class _WindowGraphAttention2(torch.autograd.Function):
    """The same op on K2b / K3b (any window of W in {16, 32, 64} keypoints): projection GEMM + tcgen05 attention
    core.  The projected q, k, v rows are kept for the backward (3x the activation in bf16) unless `save_qkv` is
    False, in which case the backward re-projects them."""

    @staticmethod
    def forward(ctx, xn, w_qkv, b_qkv, bits, threshold, heads, shift, layout, frames, kps, window, save_qkv, attn_p):
        lib = _lib.load()
        _need_cuda(xn, w_qkv, b_qkv, bits)
        if xn.dtype == torch.float32:
            return _WindowGraphAttention2._forward_f32(ctx, lib, xn, w_qkv, b_qkv, bits, threshold, heads, shift, layout,
                                                       frames, kps, window, attn_p)
        if xn.dtype != torch.bfloat16:
            raise _lib.HwgatError(f"unsupported dtype {xn.dtype}: the kernels take float32 or bfloat16")
        xn_c = xn.contiguous()
        d = xn_c.shape[-1]
        n_tok = xn_c.numel() // d
        if frames * kps == 0 or n_tok % (frames * kps) != 0:
            raise ValueError(f"token count {n_tok} is not a multiple of frames*keypoints = {frames * kps}")
        B = n_tok // (frames * kps)
        w_c = cast_cached(w_qkv, torch.bfloat16)
        b_c = cast_cached(b_qkv, torch.float32)
        out = torch.empty_like(xn_c)
        qkv = torch.empty((n_tok, 3 * d), dtype=torch.bfloat16, device=xn_c.device)
        ws_bytes = lib.hwgat_attn2_workspace_bytes(B, frames, kps, d, heads, 0, 1)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xn_c.device)
        seed, off = _philox_stream(xn_c.device) if attn_p > 0 else (0, 0)
        with torch.cuda.device(xn_c.device):
            check(lib.hwgat_attn2_fwd(xn_c.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), bits.data_ptr(),
                                      float(threshold), out.data_ptr(), qkv.data_ptr(), ws.data_ptr(), ws.numel(), B,
                                      frames, kps, d, heads, window, TEMPORAL_PATCH, shift, layout, float(attn_p), seed,
                                      off, _stream()), "hwgat_attn2_fwd")
        keep = save_qkv and any(ctx.needs_input_grad[:3])
        ctx.save_for_backward(xn_c, w_c, b_c, bits, qkv if keep else None)
        ctx.meta = (float(threshold), heads, shift, layout, frames, kps, B, d, window, w_qkv.dtype, b_qkv.dtype,
                    float(attn_p), seed, off)
        return out.view_as(xn)

    @staticmethod
    def _forward_f32(ctx, lib, xn, w_qkv, b_qkv, bits, threshold, heads, shift, layout, frames, kps, window, attn_p):
        """the fp32 parity mode (attn_win_f32.cu): window_size 32 / 64, true fp32, qkv always kept"""
        if attn_p > 0:
            raise _lib.HwgatError("attention dropout is built into the bf16 kernels only (K2b / K3b)")
        xn_c = xn.contiguous()
        d = xn_c.shape[-1]
        n_tok = xn_c.numel() // d
        if frames * kps == 0 or n_tok % (frames * kps) != 0:
            raise ValueError(f"token count {n_tok} is not a multiple of frames*keypoints = {frames * kps}")
        B = n_tok // (frames * kps)
        w_c = cast_cached(w_qkv, torch.float32)
        b_c = cast_cached(b_qkv, torch.float32)
        out = torch.empty_like(xn_c)
        qkv = torch.empty((n_tok, 3 * d), dtype=torch.float32, device=xn_c.device)
        with torch.cuda.device(xn_c.device):
            check(lib.hwgat_attn2_fwd_f32(xn_c.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), bits.data_ptr(),
                                          float(threshold), out.data_ptr(), qkv.data_ptr(), B, frames, kps, d, heads,
                                          window, TEMPORAL_PATCH, shift, layout, _stream()), "hwgat_attn2_fwd_f32")
        ctx.save_for_backward(xn_c, w_c, b_c, bits, qkv if any(ctx.needs_input_grad[:3]) else None)
        ctx.meta = (float(threshold), heads, shift, layout, frames, kps, B, d, window, w_qkv.dtype, b_qkv.dtype,
                    0.0, 0, 0)
        return out.view_as(xn)

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        xn_c, w_c, b_c, bits, qkv = ctx.saved_tensors
        threshold, heads, shift, layout, frames, kps, B, d, window, w_dtype, b_dtype, attn_p, seed, off = ctx.meta
        if xn_c.dtype == torch.float32:
            g = d_out.float().contiguous()
            d_xn = torch.empty_like(xn_c)
            d_w = torch.empty((3 * d, d), dtype=torch.float32, device=xn_c.device)
            d_b = torch.empty((3 * d,), dtype=torch.float32, device=xn_c.device)
            ws_bytes = lib.hwgat_attn2_f32_workspace_bytes(B, frames, kps, d, 1)
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xn_c.device)
            with torch.cuda.device(xn_c.device):
                check(lib.hwgat_attn2_bwd_f32(g.data_ptr(), xn_c.data_ptr(), w_c.data_ptr(), qkv.data_ptr(),
                                              bits.data_ptr(), threshold, d_xn.data_ptr(), d_w.data_ptr(),
                                              d_b.data_ptr(), ws.data_ptr(), ws.numel(), B, frames, kps, d, heads,
                                              window, TEMPORAL_PATCH, shift, layout, _stream()), "hwgat_attn2_bwd_f32")
            return (d_xn.view_as(d_out), d_w.to(w_dtype), d_b.to(b_dtype)) + (None,) * 10
        g = d_out.to(torch.bfloat16).contiguous()
        d_xn = torch.empty_like(xn_c)
        d_w = torch.empty((3 * d, d), dtype=torch.float32, device=xn_c.device)
        d_b = torch.empty((3 * d,), dtype=torch.float32, device=xn_c.device)
        ws_bytes = lib.hwgat_attn2_workspace_bytes(B, frames, kps, d, heads, 1, 0 if qkv is None else 1)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xn_c.device)
        with torch.cuda.device(xn_c.device):
            check(lib.hwgat_attn2_bwd(g.data_ptr(), xn_c.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), _ptr(qkv),
                                      bits.data_ptr(), threshold, d_xn.data_ptr(), d_w.data_ptr(), d_b.data_ptr(),
                                      ws.data_ptr(), ws.numel(), B, frames, kps, d, heads, window, TEMPORAL_PATCH,
                                      shift, layout, 0, attn_p, seed, off, _stream()), "hwgat_attn2_bwd")
        return (d_xn.view_as(d_out), d_w.to(w_dtype), d_b.to(b_dtype)) + (None,) * 10


def window_graph_attention(xn: torch.Tensor, w_qkv: torch.Tensor, b_qkv: torch.Tensor, bits: torch.Tensor,
                           heads: int, shift: int = 0, threshold: Optional[float] = None,
                           layout: int = LAYOUT_BFKD, frames: Optional[int] = None,
                           kps: Optional[int] = None, window: int = WINDOW, impl: Optional[str] = None,
                           save_qkv: bool = True, attn_drop: float = 0.0) -> torch.Tensor:
    """Fused roll + window_partition + QKV + masked attention + window_reverse +
    roll back (HWGATE.py:197-201, 86-114, 207-215), without the output projection.

    xn: (B, F, K, d) for LAYOUT_BFKD, or (B*f*nW, 2*window, d) for LAYOUT_WINDOWS (then
    `frames` and `kps` must be given).  threshold None = eval mode.  window: keypoints per window
    (16 = the reference; 32 and 64 are bf16 only).  impl: "fused" | "tc2" | None (= ops.ATTN_IMPL), see above.
    attn_drop > 0: dropout on the attention probabilities (self.attn_drop, HWGATE.py:112; pass 0 in eval mode) - built
    into K2b / K3b only, so it selects them (bf16)."""
    if layout == LAYOUT_BFKD:
        frames, kps = xn.shape[1], xn.shape[2]
    elif frames is None or kps is None:
        raise ValueError("LAYOUT_WINDOWS needs frames and kps")
    thr = -1.0 if threshold is None else float(threshold)
    if window not in WINDOWS:
        raise _lib.HwgatError(f"window_size {window} is not supported by the sm_100a kernels (16, 32, 64; no fallback)")
    impl = impl or ATTN_IMPL
    if not 0.0 <= attn_drop < 1.0:
        raise ValueError("attn_drop must be in [0, 1)")
    if window != WINDOW or attn_drop > 0 or (impl == "tc2" and xn.dtype == torch.bfloat16):
        return _WindowGraphAttention2.apply(xn, w_qkv, b_qkv, bits, thr, heads, shift, layout, frames, kps, window,
                                            save_qkv, float(attn_drop))
    hybrid = impl == "hybrid" or (impl == "fused" and HYBRID_MIN_D and xn.shape[-1] >= HYBRID_MIN_D)
    return _WindowGraphAttention.apply(xn, w_qkv, b_qkv, bits, thr, heads, shift, layout, frames, kps, hybrid)


# --------------------------------------------------------------------------
# K15 / K16: frame-banded graph attention (the sibling models WGATE and GATE)
# --------------------------------------------------------------------------

BAND_WINDOWS = (16, 32)


def band_mask_pack(adj_mask: torch.Tensor, frames: int, window: int) -> torch.Tensor:
    """Pack the ADDITIVE mask of WGATE / GATE - 0 on graph edges, -10000 elsewhere (WGATE.py:190, GATE.py:142) - into
    the (nW, window, 3) uint32 band words of K15 / K16, after proving that the band is all there is.

    adj_mask: (nW, frames*kw, frames*kw) with the token order f*kw + k of WGATE's window_partition (WGATE.py:32-45),
    kw <= window keypoints per window; keypoints kw .. window-1 of the packed words are padding (no bits: the stream
    stores `window` keypoints per window).  GATE's (1, 1, F*29, F*29) buffer is the nW = 1, kw = 29 case.  Requirements, checked here once per mask (the result
    is cached by the callers) and refused loudly otherwise - there is no dense-attention fallback:
      * every entry is 0 or -10000;
      * entries between tokens more than one frame apart are -10000 (frame-banded);
      * the (f, f-1), (f, f) and (f, f+1) blocks do not depend on f (frame-invariant);
      * every token has at least one edge (so the -10000 terms vanish in the reference's fp32 softmax and the banded
        softmax equals it).
    Integer / bit work only; runs as a handful of torch ops on the mask's device at model set-up, not on the step."""
    m = adj_mask.detach()
    if m.dim() in (2, 4):
        m = m.reshape(-1, m.shape[-2], m.shape[-1])
    nW, N, N2 = m.shape
    if N != N2 or N % frames:
        raise _lib.HwgatError(f"adjacency mask {tuple(adj_mask.shape)} does not cover {frames} frames")
    kw = N // frames
    if kw > window or window not in BAND_WINDOWS:
        raise _lib.HwgatError(f"band attention: {kw} keypoints per window do not fit the kernels' window of {window}")
    edge = m == 0
    if not bool((edge | (m == -10000.0)).all()):
        raise _lib.HwgatError("band attention: the additive mask must hold only 0 (edge) and -10000 (no edge)")
    e = edge.reshape(nW, frames, kw, frames, kw).permute(0, 1, 3, 2, 4)          # (nW, fq, fk, i, j)
    fq = torch.arange(frames, device=m.device)
    far = (fq[:, None] - fq[None, :]).abs() > 1
    if bool(e[:, far].any()):
        raise _lib.HwgatError("band attention: the graph links tokens more than one frame apart; the sm_100a kernels "
                              "evaluate the frame band only and there is no dense fallback")
    blocks = []
    for r in (-1, 0, 1):                                                        # key frame = query frame + r
        q = fq[(fq + r >= 0) & (fq + r < frames)]
        if q.numel() == 0:
            blocks.append(torch.zeros(nW, kw, kw, dtype=torch.bool, device=m.device))
            continue
        blk = e[:, q, q + r]                                                    # (nW, n, i, j)
        if not bool((blk == blk[:, :1]).all()):
            raise _lib.HwgatError("band attention: the graph changes from frame to frame (not frame-invariant)")
        blocks.append(blk[:, 0])
    band = torch.stack(blocks, dim=2)                                           # (nW, i, 3, j)
    if not bool(e.any(dim=-1).any(dim=2).all()):
        raise _lib.HwgatError("band attention: a token without any edge (its softmax would spread over all tokens)")
    weights = (1 << torch.arange(kw, device=m.device, dtype=torch.int64))
    words = (band.to(torch.int64) * weights).sum(dim=-1)                        # (nW, kw, 3)
    out = torch.zeros(nW, window, 3, dtype=torch.int64, device=m.device)
    out[:, :kw] = words
    # uint32 payload in an int32 tensor (bit 31 = keypoint 31)
    out = torch.where(out >= 2 ** 31, out - 2 ** 32, out).to(torch.int32).contiguous()
    # the blocks between adjacent frames are (a subset of) the identity in every graph the reference builds: K15 / K16
    # then evaluate only their diagonal (`diag` of hwgat_band_attn_fwd); one host read at set-up, cached by the callers
    row_bit = weights[None, :, None]
    out.band_diag = bool(((words[:, :, (0, 2)] & ~row_bit) == 0).all().item())
    return out


class _BandGraphAttention(torch.autograd.Function):
    """QKV projection + frame-banded graph attention (K15) ; backward K16 + the three weight-side GEMMs.  bf16: the
    tcgen05 / TMA / HMMA kernels; fp32: the true-fp32 parity kernels (attn_f32.cu)."""

    @staticmethod
    def forward(ctx, xn, w_qkv, b_qkv, bits, heads, window, diag):
        lib = _lib.load()
        _need_cuda(xn, w_qkv, b_qkv, bits)
        if xn.dim() != 4:
            raise _lib.HwgatError("band attention takes the (B, F, K, d) stream")
        code = _dtype_code(xn)
        xn_c = xn.contiguous()
        B, F, K, d = xn_c.shape
        if bits.shape != (K // window, window, 3):
            raise ValueError(f"band words {tuple(bits.shape)} do not match K = {K}, window = {window}")
        w_c = cast_cached(w_qkv, xn_c.dtype)
        b_c = cast_cached(b_qkv, torch.float32)
        n_tok = B * F * K
        out = torch.empty_like(xn_c)
        qkv = torch.empty((n_tok, 3 * d), dtype=xn_c.dtype, device=xn_c.device)
        need = any(ctx.needs_input_grad[:3])
        lse = torch.empty((n_tok, heads), dtype=torch.float32, device=xn_c.device) if need else None
        with torch.cuda.device(xn_c.device):
            check(lib.hwgat_band_attn_fwd(code, xn_c.data_ptr(), w_c.data_ptr(), b_c.data_ptr(), bits.data_ptr(),
                                          out.data_ptr(), qkv.data_ptr(), _ptr(lse), B, F, K, d, heads, window,
                                          int(diag), _stream()), "hwgat_band_attn_fwd")
        if need:
            ctx.save_for_backward(xn_c, w_c, bits, qkv, out, lse)
        ctx.meta = (heads, window, w_qkv.dtype, b_qkv.dtype, int(diag), code)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        xn_c, w_c, bits, qkv, out, lse = ctx.saved_tensors
        heads, window, w_dtype, b_dtype, diag, code = ctx.meta
        B, F, K, d = xn_c.shape
        g = d_out.to(xn_c.dtype).contiguous()
        d_xn = torch.empty_like(xn_c)
        d_w = torch.empty((3 * d, d), dtype=torch.float32, device=xn_c.device)
        d_b = torch.empty((3 * d,), dtype=torch.float32, device=xn_c.device)
        ws_bytes = lib.hwgat_band_attn_workspace_bytes(code, B, F, K, d, 1)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xn_c.device)
        with torch.cuda.device(xn_c.device):
            check(lib.hwgat_band_attn_bwd(code, g.data_ptr(), xn_c.data_ptr(), w_c.data_ptr(), qkv.data_ptr(),
                                          out.data_ptr(), lse.data_ptr(), bits.data_ptr(), d_xn.data_ptr(),
                                          d_w.data_ptr(), d_b.data_ptr(), ws.data_ptr(), ws.numel(), B, F, K, d, heads,
                                          window, diag, _stream()), "hwgat_band_attn_bwd")
        return d_xn, d_w.to(w_dtype), d_b.to(b_dtype), None, None, None, None


def band_attention_supported(B: int, F: int, K: int, d: int, heads: int, window: int) -> bool:
    return (window in BAND_WINDOWS and K % window == 0 and d % 128 == 0 and d % heads == 0
            and d // heads in (16, 32, 64) and (B * F * K) % 128 == 0)


def band_graph_attention(xn: torch.Tensor, w_qkv: torch.Tensor, b_qkv: torch.Tensor, bits: torch.Tensor, heads: int,
                         window: int, diag: Optional[bool] = None) -> torch.Tensor:
    """window_partition + QKV + additive-masked attention over all frames + window_reverse of WGATE
    (WGATE.py:150-158, 87-106) / the masked full attention of GATE (GATE.py:49-66), without the output projection, on
    the (B, F, K, d) stream: bf16 = the timed kernels, fp32 = the 1e-5 parity kernels.  bits: band_mask_pack(...).  The attention evaluates the graph's frame band only
    (see band_mask_pack for why that equals the reference's dense softmax).  diag: the off-frame blocks are the
    identity (None: what band_mask_pack found; False forces the general path)."""
    if diag is None:
        diag = getattr(bits, "band_diag", False)
    return _BandGraphAttention.apply(xn, w_qkv, b_qkv, bits, heads, window, bool(diag))


# --------------------------------------------------------------------------
# K4: temporal merge
# --------------------------------------------------------------------------

class _TemporalMerge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        _need_cuda(x)
        code = _dtype_code(x)
        x_c = x.contiguous()
        B, F, K, d = x_c.shape
        out = torch.empty((B, F // TEMPORAL_PATCH, K, d * TEMPORAL_PATCH), dtype=x_c.dtype, device=x_c.device)
        with torch.cuda.device(x_c.device):
            check(lib.hwgat_merge_fwd(x_c.data_ptr(), out.data_ptr(), B, F, K, d, TEMPORAL_PATCH, code, _stream()),
                  "hwgat_merge_fwd")
        ctx.shape = (B, F, K, d, code)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        B, F, K, d, code = ctx.shape
        g_c = g.contiguous()
        d_x = torch.empty((B, F, K, d), dtype=g_c.dtype, device=g_c.device)
        with torch.cuda.device(g_c.device):
            check(lib.hwgat_merge_bwd(g_c.data_ptr(), d_x.data_ptr(), B, F, K, d, TEMPORAL_PATCH, code, _stream()),
                  "hwgat_merge_bwd")
        return d_x


def temporal_merge(x: torch.Tensor) -> torch.Tensor:
    """(B,F,K,d) -> (B,F/2,K,2d); replaces TemporalMerging.forward (HWGATE.py:55-63)."""
    if x.shape[1] % TEMPORAL_PATCH != 0:
        raise ValueError("frame count must be a multiple of the temporal patch size")
    return _TemporalMerge.apply(x)


# --------------------------------------------------------------------------
# K5-K7: bandwidth-bound fusions of the rest of the block (bf16 / autocast path)
# --------------------------------------------------------------------------

def _philox_stream(device) -> tuple:
    """(seed, offset) for one dropout call, taken from (and advancing) PyTorch's CUDA generator, so
    torch.manual_seed / torch.cuda.manual_seed make the masks reproducible.  The CPU generator - the
    one the training threshold is drawn from (HWGATE.py:96) - is not touched."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, off = gen.initial_seed(), gen.get_offset()
    gen.set_offset(off + 4)
    return seed & 0xFFFFFFFFFFFFFFFF, off


def _ew(lib, name: str, io: torch.dtype):
    """K5 - K7 entry point for activations of dtype `io`: bf16 (autocast path) or float32 (fp32 path, `_f32` forms)"""
    if io == torch.float32:
        return getattr(lib, name + "_f32")
    if io != torch.bfloat16:
        raise _lib.HwgatError(f"activation dtype {io}: the elementwise kernels take bfloat16 or float32")
    return getattr(lib, name)


class _LayerNormResidual(torch.autograd.Function):
    """(x) -> (x, LayerNorm(x) as bf16).  The first output is x itself: routing the residual branch
    through it lets backward add the residual gradient inside the LayerNorm-backward pass (K5')."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, io=torch.bfloat16):
        lib = _lib.load()
        _need_cuda(x, gamma, beta)
        if x.dtype != torch.float32:
            raise _lib.HwgatError("layer_norm_residual takes the fp32 residual stream")
        x_c = x.contiguous()
        d = x_c.shape[-1]
        n = x_c.numel() // d
        g_c, b_c = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        y = torch.empty(x_c.shape, dtype=io, device=x_c.device)
        mean = torch.empty(n, dtype=torch.float32, device=x_c.device)
        rstd = torch.empty(n, dtype=torch.float32, device=x_c.device)
        with torch.cuda.device(x_c.device):
            check(_ew(lib, "hwgat_ln_fwd", io)(x_c.data_ptr(), g_c.data_ptr(), b_c.data_ptr(), y.data_ptr(),
                                               mean.data_ptr(), rstd.data_ptr(), n, d, float(eps), _stream()),
                  "hwgat_ln_fwd")
        ctx.save_for_backward(x_c, g_c, mean, rstd)
        ctx.meta = (n, d, gamma.dtype, beta.dtype, io)
        return x_c.detach(), y

    @staticmethod
    def backward(ctx, g_res, g_y):
        lib = _lib.load()
        x_c, g_c, mean, rstd = ctx.saved_tensors
        n, d, gdt, bdt, io = ctx.meta
        if g_y is None:
            return g_res, None, None, None, None
        dy = g_y.to(io).contiguous()
        dres = g_res.float().contiguous() if g_res is not None else None
        dx = torch.empty_like(x_c)
        dgamma = torch.empty(d, dtype=torch.float32, device=x_c.device)
        dbeta = torch.empty(d, dtype=torch.float32, device=x_c.device)
        with torch.cuda.device(x_c.device):
            check(_ew(lib, "hwgat_ln_bwd", io)(dy.data_ptr(), _ptr(dres), x_c.data_ptr(), mean.data_ptr(),
                                               rstd.data_ptr(), g_c.data_ptr(), dx.data_ptr(), dgamma.data_ptr(),
                                               dbeta.data_ptr(), n, d, _stream()), "hwgat_ln_bwd")
        return dx, dgamma.to(gdt), dbeta.to(bdt), None, None


def layer_norm_residual(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                        io: torch.dtype = torch.bfloat16):
    """Returns (x, y): y = LayerNorm(x) in bf16 (norm1 / norm2 of the block + the autocast cast; io = torch.float32:
    in fp32, the fp32 path); use the returned x for the residual add."""
    return _LayerNormResidual.apply(x, gamma, beta, eps, io)


class _BiasDropoutAddLN(torch.autograd.Function):
    """x1 = res + dropout(a0 + bias); optionally y = LayerNorm(x1) as bf16 in the same pass (K6 / K6')."""

    @staticmethod
    def forward(ctx, res, a0, bias, gamma, beta, eps, p, io=torch.bfloat16):
        lib = _lib.load()
        _need_cuda(res, a0, bias, gamma, beta)
        res_c, a_c = res.contiguous(), a0.to(io).contiguous()
        if res_c.dtype != torch.float32 or res_c.shape != a_c.shape:
            raise _lib.HwgatError("bias_dropout_add_ln takes an fp32 residual and a same-shape branch")
        d = res_c.shape[-1]
        n = res_c.numel() // d
        dev = res_c.device
        has_ln = gamma is not None
        b_c = bias.detach().float().contiguous() if bias is not None else None
        g_c = gamma.detach().float().contiguous() if has_ln else None
        bt_c = beta.detach().float().contiguous() if has_ln else None
        seed, off = _philox_stream(dev) if p > 0 else (0, 0)
        x1 = torch.empty_like(res_c)
        y = torch.empty(res_c.shape, dtype=io, device=dev) if has_ln else None
        mean = torch.empty(n, dtype=torch.float32, device=dev) if has_ln else None
        rstd = torch.empty(n, dtype=torch.float32, device=dev) if has_ln else None
        with torch.cuda.device(dev):
            check(_ew(lib, "hwgat_bda_ln_fwd", io)(res_c.data_ptr(), a_c.data_ptr(), _ptr(b_c), _ptr(g_c), _ptr(bt_c),
                                                   x1.data_ptr(), _ptr(y), _ptr(mean), _ptr(rstd), n, d, float(eps),
                                                   float(p), seed, off, _stream()), "hwgat_bda_ln_fwd")
        if has_ln:
            ctx.save_for_backward(x1, g_c, mean, rstd)
        ctx.meta = (n, d, float(p), seed, off, has_ln, a0.dtype,
                    None if bias is None else bias.dtype, None if gamma is None else gamma.dtype,
                    None if beta is None else beta.dtype, io)
        if has_ln:
            return x1, y
        return x1, None

    @staticmethod
    def backward(ctx, g_x1, g_y):
        lib = _lib.load()
        n, d, p, seed, off, has_ln, adt, bdt, gdt, btdt, io = ctx.meta
        dev = (g_x1 if g_x1 is not None else g_y).device
        gx = g_x1.float().contiguous() if g_x1 is not None else None
        d_a0 = torch.empty((n, d), dtype=io, device=dev)
        bda_bwd = _ew(lib, "hwgat_bda_ln_bwd", io)
        dbias = torch.empty(d, dtype=torch.float32, device=dev) if bdt is not None else None
        if has_ln:
            x1, g_c, mean, rstd = ctx.saved_tensors
            dy = (g_y if g_y is not None else torch.zeros_like(x1, dtype=io)).to(io).contiguous()
            d_res = torch.empty_like(x1)
            dgamma = torch.empty(d, dtype=torch.float32, device=dev)
            dbeta = torch.empty(d, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                check(bda_bwd(_ptr(gx), dy.data_ptr(), x1.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                              g_c.data_ptr(), d_res.data_ptr(), d_a0.data_ptr(), _ptr(dbias),
                              dgamma.data_ptr(), dbeta.data_ptr(), n, d, p, seed, off, _stream()),
                      "hwgat_bda_ln_bwd")
            shape = x1.shape
            return (d_res, d_a0.view(shape).to(adt), None if dbias is None else dbias.to(bdt), dgamma.to(gdt),
                    dbeta.to(btdt), None, None, None)
        with torch.cuda.device(dev):
            check(bda_bwd(gx.data_ptr(), 0, 0, 0, 0, 0, 0, d_a0.data_ptr(), _ptr(dbias), 0, 0, n, d, p,
                          seed, off, _stream()), "hwgat_bda_ln_bwd")
        return gx, d_a0.view(gx.shape).to(adt), None if dbias is None else dbias.to(bdt), None, None, None, None, None


def bias_dropout_add_ln(res: torch.Tensor, a0: torch.Tensor, bias: Optional[torch.Tensor], norm, p: float,
                        training: bool, io: torch.dtype = torch.bfloat16):
    """x1 = res + dropout(a0 + bias) [Linear bias + proj_drop / ff.drop + shortcut, HWGATE.py:115-116, 134-135,
    217, 219] and, if `norm` (an nn.LayerNorm) is given, y = norm(x1) as bf16 in the same pass.  Returns (x1, y)."""
    if norm is None:
        return _BiasDropoutAddLN.apply(res, a0, bias, None, None, 0.0, p if training else 0.0, io)
    return _BiasDropoutAddLN.apply(res, a0, bias, norm.weight, norm.bias, norm.eps, p if training else 0.0, io)


class _BdaMergeLN(torch.autograd.Function):
    """Level boundary in one Function: x1 = res + dropout(a0 + bias) stored directly in TemporalMerging's layout
    (B, F/2, K, 2d) [K6 with the merge folded in], then y = LayerNorm_{2d}(x1) as bf16 [K5 of the next level's first
    norm1].  Backward: K5' stores its result un-merged [the merge's adjoint folded in], then K6' (no-LN form).
    Replaces K6 + K4 + K5 and K5' + K4' + K6' (HWGATE.py:134-135, 219, 55-63, 203)."""

    @staticmethod
    def forward(ctx, res, a0, bias, gamma, beta, eps, p, io=torch.bfloat16):
        lib = _lib.load()
        _need_cuda(res, a0, bias, gamma, beta)
        res_c, a_c = res.contiguous(), a0.to(io).contiguous()
        if res_c.dtype != torch.float32 or res_c.shape != a_c.shape or res_c.dim() != 4:
            raise _lib.HwgatError("bias_dropout_add_merge_ln takes an fp32 (B,F,K,d) residual and a same-shape branch")
        B, F, K, d = res_c.shape
        n, dev = B * F * K, res_c.device
        b_c = cast_cached(bias, torch.float32) if bias is not None else None
        g_c, bt_c = cast_cached(gamma, torch.float32), cast_cached(beta, torch.float32)
        seed, off = _philox_stream(dev) if p > 0 else (0, 0)
        xm = torch.empty((B, F // 2, K, 2 * d), dtype=torch.float32, device=dev)
        y = torch.empty((B, F // 2, K, 2 * d), dtype=io, device=dev)
        mean = torch.empty(n // 2, dtype=torch.float32, device=dev)
        rstd = torch.empty(n // 2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_ew(lib, "hwgat_bda_merge_fwd", io)(res_c.data_ptr(), a_c.data_ptr(), _ptr(b_c), xm.data_ptr(), n, d, F,
                                                      K, float(p), seed, off, _stream()), "hwgat_bda_merge_fwd")
            check(_ew(lib, "hwgat_ln_fwd", io)(xm.data_ptr(), g_c.data_ptr(), bt_c.data_ptr(), y.data_ptr(),
                                               mean.data_ptr(), rstd.data_ptr(), n // 2, 2 * d, float(eps), _stream()),
                  "hwgat_ln_fwd")
        ctx.save_for_backward(xm, g_c, mean, rstd)
        ctx.meta = (B, F, K, d, float(p), seed, off, a0.dtype, None if bias is None else bias.dtype, gamma.dtype,
                    beta.dtype, io)
        return xm, y

    @staticmethod
    def backward(ctx, g_xm, g_y):
        lib = _lib.load()
        xm, g_c, mean, rstd = ctx.saved_tensors
        B, F, K, d, p, seed, off, adt, bdt, gdt, btdt, io = ctx.meta
        n, dev = B * F * K, xm.device
        dy = (g_y if g_y is not None else torch.zeros_like(xm, dtype=io)).to(io).contiguous()
        gx = g_xm.float().contiguous() if g_xm is not None else None
        d_x1 = torch.empty((B, F, K, d), dtype=torch.float32, device=dev)
        dgamma = torch.empty(2 * d, dtype=torch.float32, device=dev)
        dbeta = torch.empty(2 * d, dtype=torch.float32, device=dev)
        d_a0 = torch.empty((B, F, K, d), dtype=io, device=dev)
        dbias = torch.empty(d, dtype=torch.float32, device=dev) if bdt is not None else None
        with torch.cuda.device(dev):
            check(_ew(lib, "hwgat_ln_bwd_unmerge", io)(dy.data_ptr(), _ptr(gx), xm.data_ptr(), mean.data_ptr(),
                                                       rstd.data_ptr(), g_c.data_ptr(), d_x1.data_ptr(),
                                                       dgamma.data_ptr(), dbeta.data_ptr(), n // 2, 2 * d, F // 2, K,
                                                       _stream()), "hwgat_ln_bwd_unmerge")
            check(_ew(lib, "hwgat_bda_ln_bwd", io)(d_x1.data_ptr(), 0, 0, 0, 0, 0, 0, d_a0.data_ptr(), _ptr(dbias), 0, 0,
                                                   n, d, p, seed, off, _stream()), "hwgat_bda_ln_bwd")
        return (d_x1, d_a0.to(adt), None if dbias is None else dbias.to(bdt), dgamma.to(gdt), dbeta.to(btdt), None,
                None, None)


def bias_dropout_add_merge_ln(res: torch.Tensor, a0: torch.Tensor, bias: Optional[torch.Tensor], next_norm, p: float,
                              training: bool, io: torch.dtype = torch.bfloat16):
    """The last residual add of a level, TemporalMerging and the next level's first LayerNorm:
    returns (x_merged fp32 (B,F/2,K,2d), next_norm(x_merged) as bf16)."""
    return _BdaMergeLN.apply(res, a0, bias, next_norm.weight, next_norm.bias, next_norm.eps, p if training else 0.0, io)


def merge_fold_supported(d: int, frames: int) -> bool:
    return d in (128, 256) and frames % 2 == 0


class _BiasGeluDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u0, bias, p, io=torch.bfloat16):
        lib = _lib.load()
        _need_cuda(u0, bias)
        u_c = u0.to(io).contiguous()
        cols = u_c.shape[-1]
        n = u_c.numel() // cols
        b_c = bias.detach().float().contiguous() if bias is not None else None
        seed, off = _philox_stream(u_c.device) if p > 0 else (0, 0)
        g = torch.empty_like(u_c)
        with torch.cuda.device(u_c.device):
            check(_ew(lib, "hwgat_bias_gelu_dropout_fwd", io)(u_c.data_ptr(), _ptr(b_c), g.data_ptr(), n, cols, float(p),
                                                              seed, off, _stream()), "hwgat_bias_gelu_dropout_fwd")
        ctx.save_for_backward(u_c, b_c)
        ctx.meta = (n, cols, float(p), seed, off, u0.dtype, None if bias is None else bias.dtype, io)
        return g

    @staticmethod
    def backward(ctx, dg):
        lib = _lib.load()
        u_c, b_c = ctx.saved_tensors
        n, cols, p, seed, off, udt, bdt, io = ctx.meta
        dg_c = dg.to(io).contiguous()
        du = torch.empty_like(u_c)
        dbias = torch.empty(cols, dtype=torch.float32, device=u_c.device) if bdt is not None else None
        with torch.cuda.device(u_c.device):
            check(_ew(lib, "hwgat_bias_gelu_dropout_bwd", io)(u_c.data_ptr(), _ptr(b_c), dg_c.data_ptr(), du.data_ptr(),
                                                              _ptr(dbias), n, cols, p, seed, off, _stream()),
                  "hwgat_bias_gelu_dropout_bwd")
        return du.to(udt), None if dbias is None else dbias.to(bdt), None, None


def bias_gelu_dropout(u0: torch.Tensor, bias: Optional[torch.Tensor], p: float, training: bool,
                      io: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """dropout(gelu(u0 + bias)), exact erf GELU: fc1 bias + ff.act + ff.drop (HWGATE.py:131-133)."""
    return _BiasGeluDropout.apply(u0, bias, p if training else 0.0, io)


FFN_FUSED = os.environ.get("HWGAT_FFN_FUSED", "1") != "0"     # A/B switch of the one-kernel inference FeedForward (K10f)


class _FeedForwardCore(torch.autograd.Function):
    """v0 = dropout(gelu(h W1^T + b1)) W2^T on the tcgen05 GEMMs of K10 (bias, GELU and dropout in the first GEMM's
    epilogue).  Saves h, the hidden activation and its local derivative (mask folded in); nothing is recomputed."""

    @staticmethod
    def forward(ctx, h, w1, b1, w2, p, grad_mode=True):
        lib = _lib.load()
        _need_cuda(h, w1, b1, w2)
        h_c = h.to(torch.bfloat16).contiguous()
        d = h_c.shape[-1]
        n = h_c.numel() // d
        hidden = w1.shape[0]
        if tuple(w1.shape) != (hidden, d) or tuple(w2.shape) != (d, hidden):
            raise ValueError(f"fc1 / fc2 weights {tuple(w1.shape)} / {tuple(w2.shape)} do not match (n, {d}) input")
        w1_c = cast_cached(w1, torch.bfloat16)
        w2_c = cast_cached(w2, torch.bfloat16)
        b1_c = cast_cached(b1, torch.float32) if b1 is not None else None
        # needs_input_grad is also set under torch.no_grad() (parameters still require grad) and grad mode is always
        # off inside forward: the caller passes the grad mode it saw.  Inference then takes the GELU-only epilogue and
        # writes no local derivative (1.6 GB per call at batch 256, T = 192)
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad[:4])
        seed, off = _philox_stream(h_c.device) if p > 0 else (0, 0)
        # inference: one kernel, the activation tile never leaves the SM (K10f), where the shape allows
        fused = (not need_grad and p == 0 and FFN_FUSED and bool(lib.hwgat_ffn_fused_supported(n, d, hidden)))
        act = None if fused else torch.empty((n, hidden), dtype=torch.bfloat16, device=h_c.device)
        gp = torch.empty_like(act) if need_grad else None
        v0 = torch.empty(h_c.shape, dtype=torch.bfloat16, device=h_c.device)
        with torch.cuda.device(h_c.device):
            check(lib.hwgat_ffn_fwd(h_c.data_ptr(), w1_c.data_ptr(), _ptr(b1_c), w2_c.data_ptr(), _ptr(act),
                                    _ptr(gp), v0.data_ptr(), n, d, hidden, float(p), seed, off, _stream()),
                  "hwgat_ffn_fwd")
        if need_grad:
            ctx.save_for_backward(h_c, act, gp, w1_c, w2_c)
        ctx.meta = (n, d, hidden, h.dtype, w1.dtype, None if b1 is None else b1.dtype, w2.dtype)
        return v0

    @staticmethod
    def backward(ctx, dv0):
        lib = _lib.load()
        h_c, act, gp, w1_c, w2_c = ctx.saved_tensors
        n, d, hidden, hdt, w1dt, b1dt, w2dt = ctx.meta
        dv = dv0.to(torch.bfloat16).contiguous()
        dev = h_c.device
        dh = torch.empty_like(h_c)
        dw1 = torch.empty((hidden, d), dtype=torch.float32, device=dev)
        db1 = torch.empty((hidden,), dtype=torch.float32, device=dev)
        dw2 = torch.empty((d, hidden), dtype=torch.float32, device=dev)
        ws_bytes = lib.hwgat_ffn_bwd_workspace_bytes(n, d, hidden)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.hwgat_ffn_bwd(dv.data_ptr(), h_c.data_ptr(), act.data_ptr(), gp.data_ptr(), w1_c.data_ptr(),
                                    w2_c.data_ptr(), dh.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(),
                                    ws.data_ptr(), ws.numel(), n, d, hidden, _stream()), "hwgat_ffn_bwd")
        return (dh.to(hdt), dw1.to(w1dt), None if b1dt is None else db1.to(b1dt), dw2.to(w2dt), None, None)


def feed_forward_core(h: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor, p: float,
                      training: bool) -> torch.Tensor:
    """dropout(gelu(h @ w1.T + b1)) @ w2.T as bf16: ff.fc1, ff.act, ff.drop and ff.fc2's matmul (HWGATE.py:130-134);
    fc2's bias, the second dropout and the residual add are K6's (bias_dropout_add_ln).
    h: (..., d) with prod(...) % 128 == 0, d % 128 == 0, hidden % 128 == 0, hidden <= 2048."""
    return _FeedForwardCore.apply(h, w1, b1, w2, p if training else 0.0, torch.is_grad_enabled())


def ffn_supported(n_tokens: int, d: int, hidden: int) -> bool:
    return n_tokens % 128 == 0 and d % 128 == 0 and hidden % 128 == 0 and hidden <= 2048


class _OutputProjection(torch.autograd.Function):
    """y = ctx W^T (bf16) on the library's tcgen05 GEMMs (K12): self.proj's matmul, HWGATE.py:115."""

    @staticmethod
    def forward(ctx_, x, w):
        lib = _lib.load()
        _need_cuda(x, w)
        x_c = x.to(torch.bfloat16).contiguous()
        d_in = x_c.shape[-1]
        n = x_c.numel() // d_in
        d_out = w.shape[0]
        if tuple(w.shape) != (d_out, d_in):
            raise ValueError(f"proj weight {tuple(w.shape)} does not match (n, {d_in}) input")
        w_c = cast_cached(w, torch.bfloat16)
        y = torch.empty(x_c.shape[:-1] + (d_out,), dtype=torch.bfloat16, device=x_c.device)
        with torch.cuda.device(x_c.device):
            check(lib.hwgat_proj_fwd(x_c.data_ptr(), w_c.data_ptr(), y.data_ptr(), n, d_in, d_out, _stream()),
                  "hwgat_proj_fwd")
        if any(ctx_.needs_input_grad[:2]):
            ctx_.save_for_backward(x_c, w_c)
        ctx_.meta = (n, d_in, d_out, x.dtype, w.dtype)
        return y

    @staticmethod
    def backward(ctx_, dy):
        lib = _lib.load()
        x_c, w_c = ctx_.saved_tensors
        n, d_in, d_out, xdt, wdt = ctx_.meta
        dy_c = dy.to(torch.bfloat16).contiguous()
        dev = x_c.device
        dx = torch.empty_like(x_c)
        dw = torch.empty((d_out, d_in), dtype=torch.float32, device=dev)
        ws = torch.empty(max(lib.hwgat_proj_bwd_workspace_bytes(d_in, d_out), 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.hwgat_proj_bwd(dy_c.data_ptr(), x_c.data_ptr(), w_c.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                     ws.data_ptr(), ws.numel(), n, d_in, d_out, _stream()), "hwgat_proj_bwd")
        return dx.to(xdt), dw.to(wdt)


def output_projection(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x @ w.T as bf16 (no bias: the bias, proj_drop and the shortcut add are K6's): MSA's self.proj matmul
    (HWGATE.py:115).  x: (..., d_in) with prod(...) % 128 == 0, d_in % 128 == 0, d_out % 128 == 0."""
    return _OutputProjection.apply(x, w)


def proj_supported(n_tokens: int, d_in: int, d_out: int) -> bool:
    return n_tokens % 128 == 0 and d_in % 128 == 0 and d_out % 128 == 0


# --------------------------------------------------------------------------
# K8 / K9: model head and tail
# --------------------------------------------------------------------------

def fourier_embed(x: torch.Tensor, fourier_b: torch.Tensor, pe: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """dropout([sin(2 pi x.B^T), cos(2 pi x.B^T)] + pe[frame]) as fp32 (B, T, K, E): the Fourier keypoint embedding
    and PositionalEncoding.forward (HWGATE.py:343-347, 25-28) in one pass.  Forward only: `B` is frozen and `pe` is
    a buffer, so no gradient flows upstream."""
    lib = _lib.load()
    _need_cuda(x, fourier_b, pe)
    if fourier_b.requires_grad or x.requires_grad:
        raise _lib.HwgatError("fourier_embed is forward-only; the reference's B matrix is frozen (HWGATE.py:299)")
    Bsz, T, K, C = x.shape
    E = 2 * fourier_b.shape[0]
    x_c = x.detach().float().contiguous()
    b_c = fourier_b.detach().float().contiguous()
    pe_c = pe.detach().float().reshape(-1, E)[:T].contiguous()
    out = torch.empty((Bsz, T, K, E), dtype=torch.float32, device=x.device)
    pp = float(p) if training else 0.0
    seed, off = _philox_stream(x.device) if pp > 0 else (0, 0)
    with torch.cuda.device(x.device):
        check(lib.hwgat_embed_fwd(x_c.data_ptr(), b_c.data_ptr(), pe_c.data_ptr(), out.data_ptr(), Bsz * T * K, C, E, K,
                                  T, pp, seed, off, _stream()), "hwgat_embed_fwd")
    return out


class _LayerNormWeightedPool(torch.autograd.Function):
    """K9 with learned token weights (GATE.py:205-207)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, tok_w, tok_b, eps, kp_real):
        lib = _lib.load()
        _need_cuda(x, gamma, beta, tok_w)
        if x.dtype != torch.float32 or x.dim() != 4:
            raise _lib.HwgatError("layer_norm_weighted_pool takes the fp32 residual stream as (B, F, K, d)")
        x_c = x.contiguous()
        Bsz, F, kp_st, d = x_c.shape
        kp_pad = kp_st if kp_real and kp_real != kp_st else 0
        kp_r = kp_real if kp_pad else 0
        tokens = F * (kp_real or kp_st)
        w_c = tok_w.detach().float().reshape(-1).contiguous()
        if w_c.numel() != tokens:
            raise ValueError(f"token weights: {w_c.numel()} for {tokens} tokens")
        g_c, b_c = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        wsum = w_c.sum()
        beta_eff = b_c * wsum + (tok_b.detach().float().reshape(()) if tok_b is not None else 0.0)
        pooled = torch.empty((Bsz, d), dtype=torch.float32, device=x_c.device)
        mean = torch.empty(Bsz * F * kp_st, dtype=torch.float32, device=x_c.device)
        rstd = torch.empty(Bsz * F * kp_st, dtype=torch.float32, device=x_c.device)
        sc_bytes = lib.hwgat_ln_pool_scratch_bytes(Bsz, tokens, d)
        scratch = torch.empty(sc_bytes, dtype=torch.uint8, device=x_c.device) if sc_bytes else None
        with torch.cuda.device(x_c.device):
            check(lib.hwgat_ln_wpool_fwd(x_c.data_ptr(), g_c.data_ptr(), beta_eff.data_ptr(), w_c.data_ptr(),
                                         pooled.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _ptr(scratch), sc_bytes,
                                         Bsz, tokens, d, float(eps), kp_r, kp_pad, _stream()), "hwgat_ln_wpool_fwd")
        ctx.save_for_backward(x_c, g_c, b_c, w_c, mean, rstd)
        ctx.meta = (Bsz, tokens, d, gamma.dtype, beta.dtype, kp_r, kp_pad, tok_w.shape, tok_w.dtype, wsum,
                    None if tok_b is None else (tok_b.shape, tok_b.dtype))
        return pooled

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x_c, g_c, b_c, w_c, mean, rstd = ctx.saved_tensors
        Bsz, tokens, d, gdt, bdt, kp_r, kp_pad, w_shape, w_dt, wsum, tb = ctx.meta
        gg = g.float().contiguous()
        dx = torch.zeros_like(x_c) if kp_pad else torch.empty_like(x_c)
        dgamma = torch.empty(d, dtype=torch.float32, device=x_c.device)
        d_w = torch.empty(tokens, dtype=torch.float32, device=x_c.device)
        part = torch.empty(max(Bsz * tokens, 1), dtype=torch.float32, device=x_c.device)
        with torch.cuda.device(x_c.device):
            check(lib.hwgat_ln_wpool_bwd(gg.data_ptr(), x_c.data_ptr(), mean.data_ptr(), rstd.data_ptr(), g_c.data_ptr(),
                                         w_c.data_ptr(), dx.data_ptr(), dgamma.data_ptr(), d_w.data_ptr(),
                                         part.data_ptr(), Bsz, tokens, d, kp_r, kp_pad, _stream()), "hwgat_ln_wpool_bwd")
        gsum = gg.sum(0)                                   # (d,) tiny
        d_w = d_w + (gsum * b_c).sum()                     # + sum_b g[b] . beta, the same for every token
        d_tb = None if tb is None else gsum.sum().reshape(tb[0]).to(tb[1])
        return dx, dgamma.to(gdt), (gsum * wsum).to(bdt), d_w.reshape(w_shape).to(w_dt), d_tb, None, None


def layer_norm_weighted_pool(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, tok_w: torch.Tensor,
                             tok_b: Optional[torch.Tensor], eps: float = 1e-5, kp_real: int = 0) -> torch.Tensor:
    """(B, F, K, d) -> (B, d): sum_t tok_w[t] * LayerNorm(x)[b, t] + tok_b - self.norm followed by GATE's learned
    `weightedAvg` Linear(F*K, 1) over the token axis (GATE.py:205-207) in one pass over x.  kp_real as in
    layer_norm_mean_pool (tok_w indexes the real tokens frame * kp_real + keypoint)."""
    return _LayerNormWeightedPool.apply(x, gamma, beta, tok_w, tok_b, eps, kp_real)


class _LayerNormPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, kp_real):
        lib = _lib.load()
        _need_cuda(x, gamma, beta)
        if x.dtype != torch.float32:
            raise _lib.HwgatError("layer_norm_mean_pool takes the fp32 residual stream")
        x_c = x.contiguous()
        Bsz, d = x_c.shape[0], x_c.shape[-1]
        stored = x_c.numel() // (Bsz * d) if Bsz else 1
        kp_pad = 0
        tokens = stored
        if kp_real:                      # padded keypoint axis (HGATE): x is (B, F, kp_pad, d), kp_real real keypoints
            kp_pad = x_c.shape[-2]
            if x_c.dim() != 4 or kp_pad < kp_real:
                raise ValueError("a padded keypoint axis needs x as (B, F, kp_pad, d) with kp_pad >= kp_real")
            tokens = stored // kp_pad * kp_real
            if kp_pad == kp_real:
                kp_real = kp_pad = 0
        g_c, b_c = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        pooled = torch.empty((Bsz, d), dtype=torch.float32, device=x_c.device)
        mean = torch.empty(Bsz * stored, dtype=torch.float32, device=x_c.device)
        rstd = torch.empty(Bsz * stored, dtype=torch.float32, device=x_c.device)
        sc_bytes = lib.hwgat_ln_pool_scratch_bytes(Bsz, tokens, d)
        scratch = torch.empty(sc_bytes, dtype=torch.uint8, device=x_c.device) if sc_bytes else None
        with torch.cuda.device(x_c.device):
            check(lib.hwgat_ln_pool_fwd(x_c.data_ptr(), g_c.data_ptr(), b_c.data_ptr(), pooled.data_ptr(),
                                        mean.data_ptr(), rstd.data_ptr(), _ptr(scratch), sc_bytes, Bsz, tokens, d,
                                        float(eps), kp_real, kp_pad, _stream()), "hwgat_ln_pool_fwd")
        ctx.save_for_backward(x_c, g_c, mean, rstd)
        ctx.meta = (Bsz, tokens, d, gamma.dtype, beta.dtype, kp_real, kp_pad)
        return pooled

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x_c, g_c, mean, rstd = ctx.saved_tensors
        Bsz, tokens, d, gdt, bdt, kp_real, kp_pad = ctx.meta
        gg = g.float().contiguous()
        dx = torch.zeros_like(x_c) if kp_pad else torch.empty_like(x_c)     # padded rows get no gradient
        dgamma = torch.empty(d, dtype=torch.float32, device=x_c.device)
        with torch.cuda.device(x_c.device):
            check(lib.hwgat_ln_pool_bwd(gg.data_ptr(), x_c.data_ptr(), mean.data_ptr(), rstd.data_ptr(), g_c.data_ptr(),
                                        dx.data_ptr(), dgamma.data_ptr(), Bsz, tokens, d, kp_real, kp_pad, _stream()),
                  "hwgat_ln_pool_bwd")
        return dx, dgamma.to(gdt), gg.sum(0).to(bdt), None, None


def layer_norm_mean_pool(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                         kp_real: int = 0) -> torch.Tensor:
    """(B, ..., d) -> (B, d): mean over all tokens of LayerNorm(x): self.norm + self.avgpool (HWGATE.py:353-354).
    kp_real > 0: x is (B, F, kp_pad, d) with only the first kp_real keypoints of every frame real (HGATE's 29
    keypoints stored as 32); the padded rows are neither pooled nor given a gradient."""
    return _LayerNormPool.apply(x, gamma, beta, eps, kp_real)


# --------------------------------------------------------------------------
# K13 / K14: classifier head and label-smoothed cross entropy
# --------------------------------------------------------------------------

class _LinearF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        lib = _lib.load()
        _need_cuda(x, w, b)
        x_c = x.float().contiguous()
        w_c = cast_cached(w, torch.float32)
        b_c = cast_cached(b, torch.float32) if b is not None else None
        d_in, d_out = x_c.shape[-1], w_c.shape[0]
        if tuple(w_c.shape) != (d_out, d_in):
            raise ValueError(f"head weight {tuple(w.shape)} does not match (n, {d_in}) input")
        n = x_c.numel() // d_in
        y = torch.empty(x_c.shape[:-1] + (d_out,), dtype=torch.float32, device=x_c.device)
        # x3 mode with a weight gradient to come: keep the bf16 planes of x (6 bytes per element) instead of x (4):
        # the weight-gradient GEMM contracts the same planes, so the backward does not split x again
        planes = None
        if ctx.needs_input_grad[1] and n > 0 and linear_x3_active(n, d_in, d_out):
            planes = torch.empty((3, n, d_in), dtype=torch.bfloat16, device=x_c.device)
        with torch.cuda.device(x_c.device):
            if planes is not None:
                check(lib.hwgat_linear_x3_fwd(x_c.data_ptr(), w_c.data_ptr(), _ptr(b_c), y.data_ptr(), planes.data_ptr(), n,
                                              d_in, d_out, _stream()), "hwgat_linear_x3_fwd")
            else:
                check(lib.hwgat_linear_f32_fwd(x_c.data_ptr(), w_c.data_ptr(), _ptr(b_c), y.data_ptr(), n, d_in, d_out,
                                               _stream()), "hwgat_linear_f32_fwd")
        ctx.save_for_backward(x_c if planes is None else None, w_c, planes)
        ctx.meta = (n, d_in, d_out, x.dtype, w.dtype, None if b is None else b.dtype, tuple(x_c.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x_c, w_c, planes = ctx.saved_tensors
        n, d_in, d_out, xdt, wdt, bdt, xshape = ctx.meta
        dy_c = dy.float().contiguous()
        dev = w_c.device
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], bdt is not None and ctx.needs_input_grad[2]
        dx = torch.empty(xshape, dtype=torch.float32, device=dev) if need_x else None
        dw = torch.empty((d_out, d_in), dtype=torch.float32, device=dev) if need_w else None
        db = torch.empty((d_out,), dtype=torch.float32, device=dev) if need_b else None
        with torch.cuda.device(dev):
            if planes is not None:
                check(lib.hwgat_linear_x3_bwd(dy_c.data_ptr(), planes.data_ptr(), w_c.data_ptr(), _ptr(dx), _ptr(dw),
                                              _ptr(db), n, d_in, d_out, _stream()), "hwgat_linear_x3_bwd")
            else:
                check(lib.hwgat_linear_f32_bwd(dy_c.data_ptr(), x_c.data_ptr(), w_c.data_ptr(), _ptr(dx), _ptr(dw),
                                               _ptr(db), n, d_in, d_out, _stream()), "hwgat_linear_f32_bwd")
        return (None if dx is None else dx.to(xdt), None if dw is None else dw.to(wdt),
                None if db is None else db.to(bdt))


def linear_f32(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """x @ w.T + b in fp32 on the library's FFMA GEMM (K13): the classifier head self.head (HWGATE.py:359)."""
    return _LinearF32.apply(x, w, b)


class _SmoothCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, smooth):
        lib = _lib.load()
        _need_cuda(logits, target)
        z = logits.float().contiguous()
        if z.dim() != 2 or target.shape != z.shape[:1]:
            raise ValueError("smooth_cross_entropy takes (rows, classes) logits and (rows,) class indices")
        t = target.to(torch.int64).contiguous()
        rows, classes = z.shape
        dev = z.device
        lse = torch.empty(rows, dtype=torch.float32, device=dev)
        row_loss = torch.empty(rows, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.hwgat_smooth_ce_fwd(z.data_ptr(), t.data_ptr(), lse.data_ptr(), row_loss.data_ptr(), loss.data_ptr(),
                                          rows, classes, float(smooth), _stream()), "hwgat_smooth_ce_fwd")
        ctx.save_for_backward(z, t, lse)
        ctx.meta = (rows, classes, float(smooth), logits.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        z, t, lse = ctx.saved_tensors
        rows, classes, smooth, dt = ctx.meta
        g_c = g.float().contiguous()
        dz = torch.empty_like(z)
        with torch.cuda.device(z.device):
            check(lib.hwgat_smooth_ce_bwd(z.data_ptr(), t.data_ptr(), lse.data_ptr(), g_c.data_ptr(), dz.data_ptr(), rows,
                                          classes, smooth, _stream()), "hwgat_smooth_ce_bwd")
        return dz.to(dt), None, None


def smooth_cross_entropy(logits: torch.Tensor, target: torch.Tensor, smooth: float = 0.01) -> torch.Tensor:
    """Label-smoothed cross entropy, mean over the batch (losses/SmoothCrossEntropy.py:35-39), as K14."""
    return _SmoothCE.apply(logits, target, smooth)
