"""CPU oracle for the HWGATE windowed-graph-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sl_hwgat_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the timed CPU baseline - never as the product path.

This is a functional restatement (plain torch on CPU, any float dtype, fp64 by
default in the tests) of what the reference computes on the path that
BASELINE.json's ``north_star`` names.  It was written from the behaviour of the
reference, not translated from it: the reference is a tree of ``nn.Module``s
that materialises rolled / partitioned copies and float masks; the oracle is a
set of pure functions over a flat ``state_dict`` that index windows directly
and treat the masks as booleans.  Every function cites the reference lines it
restates (paths relative to /root/reference/).

Parity pin: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference itself, run in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference/hwgat/models
unmodified) and committed under ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` replays them on every CPU test run.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

NEG_FILL = -10000.0  # hwgat/models/HWGATE.py:110


# --------------------------------------------------------------------------
# configuration (hwgat/models/model_params.py:243-259)
# --------------------------------------------------------------------------

# One 16-keypoint window: head(3) + arm(3) + hand(10); 25 undirected edges.
# Values restated from model_params.py:261-369 (the four lists there are equal).
_WINDOW_EDGES = (
    (0, 1), (0, 2), (0, 3), (3, 4), (4, 5), (5, 6), (6, 7), (6, 8), (8, 9),
    (8, 10), (6, 10), (10, 11), (10, 12), (6, 12), (12, 13), (12, 14),
    (14, 15), (6, 14), (7, 9), (9, 11), (11, 13), (13, 15), (7, 15), (7, 11),
    (7, 13),
)


@dataclass
class HWGATEConfig:
    """Hyper-parameters, defaults = HWGATEParams (model_params.py:245-259)."""
    kp_dim: int = 2
    num_kps: int = 64
    temporal_dim: int = 64
    num_classes: int = 262
    embed_dim: int = 128
    temporal_patch_size: int = 2
    pe: bool = True
    depths: Sequence[int] = (2, 2, 4)
    num_heads: Sequence[int] = (2, 4, 8)
    window_size: int = 16
    ff_ratio: float = 2.0
    edges: Sequence[Sequence[Sequence[int]]] = field(
        default_factory=lambda: [list(map(list, _WINDOW_EDGES))] * 4)

    @property
    def n_windows(self) -> int:
        return self.num_kps // self.window_size

    @property
    def tokens_per_window(self) -> int:
        return self.temporal_patch_size * self.window_size

    def level_dim(self, i: int) -> int:          # HWGATE.py:312
        return int(self.embed_dim * 2 ** i)

    def level_frames(self, i: int) -> int:       # HWGATE.py:314
        return self.temporal_dim // self.temporal_patch_size ** i


# --------------------------------------------------------------------------
# masks (rows a1-a3 of SURVEY.md section 8)
# --------------------------------------------------------------------------

def skeleton_adjacency(edges: Sequence[Sequence[int]], W: int) -> np.ndarray:
    """Symmetric WxW 0/1 adjacency with self loops (model_params.py:394-400)."""
    a = np.zeros((W, W), dtype=bool)
    a[np.arange(W), np.arange(W)] = True
    for i, j in edges:
        a[i, j] = True
        a[j, i] = True
    return a


def window_adjacency(edges_per_window, W: int, TP: int) -> np.ndarray:
    """(nW, TP*W, TP*W) bool.  Same-frame pairs follow the skeleton; a joint is
    linked to itself in the adjacent frame; frames further apart are not
    linked (model_params.py:373-392)."""
    nW = len(edges_per_window)
    N = TP * W
    out = np.zeros((nW, N, N), dtype=bool)
    tp = np.arange(N) // W
    kp = np.arange(N) % W
    dt = np.abs(tp[:, None] - tp[None, :])
    for w in range(nW):
        a = skeleton_adjacency(edges_per_window[w], W)
        same_frame = a[kp[:, None], kp[None, :]]
        same_joint = kp[:, None] == kp[None, :]
        out[w] = np.where(dt == 0, same_frame, np.where(dt == 1, same_joint, False))
    return out


def frame_group_ids(F: int, TP: int, shift: int) -> np.ndarray:
    """Group id per (rolled) frame used by the shifted-window mask: frames
    [0, F-TP) -> 0, [F-TP, F-shift) -> 1, [F-shift, F) -> 2 (HWGATE.py:170-178)."""
    g = np.zeros(F, dtype=np.int64)
    g[F - TP:F - shift] = 1
    g[F - shift:] = 2
    return g


def shift_window_mask(F: int, nW: int, W: int, TP: int, shift: int) -> Optional[np.ndarray]:
    """(f*nW, N, N) bool, True where query and key sit in the same frame group
    after the cyclic shift; None for unshifted blocks (HWGATE.py:169-187)."""
    if shift <= 0:
        return None
    f, N = F // TP, TP * W
    g = frame_group_ids(F, TP, shift)
    tok_frame = (np.arange(f)[:, None] * TP + (np.arange(N) // W)[None, :])  # (f, N)
    gid = g[tok_frame]                                                      # (f, N)
    m = gid[:, :, None] == gid[:, None, :]                                  # (f, N, N)
    return np.repeat(m, nW, axis=0)                                         # index fi*nW + w


def combined_mask(adj: np.ndarray, F: int, W: int, TP: int, shift: int) -> np.ndarray:
    """(f*nW, N, N) bool = replicated adjacency AND shift mask: the two
    multiplicative masks of HWGATE.py:102-108 with the replication of
    HWGATE.py:309 (window index = fi*nW + w)."""
    nW = adj.shape[0]
    f = F // TP
    m = np.tile(adj, (f, 1, 1))
    sm = shift_window_mask(F, nW, W, TP, shift)
    if sm is not None:
        m = m & sm
    return m


def pack_mask_bits(mask: np.ndarray) -> np.ndarray:
    """(nwin, N, N) bool -> (nwin, N, ceil(N/32)) uint32, bit (j%32) of word
    j//32 of row i = key j visible to query i (the packed layout of K1)."""
    nwin, N, _ = mask.shape
    words = (N + 31) // 32
    out = np.zeros((nwin, N, words), dtype=np.uint32)
    for j in range(N):
        out[:, :, j // 32] |= (mask[:, :, j].astype(np.uint32) << np.uint32(j % 32))
    return out


# --------------------------------------------------------------------------
# index maps (rows a4, a5, a13)
# --------------------------------------------------------------------------

def window_partition(x: torch.Tensor, W: int, TP: int) -> torch.Tensor:
    """(B,F,K,d) -> (B*f*nW, TP*W, d); window (b*f+fi)*nW+w, token tp*W+k
    (HWGATE.py:30-36)."""
    B, F, K, d = x.shape
    f, nW = F // TP, K // W
    return x.reshape(B, f, TP, nW, W, d).permute(0, 1, 3, 2, 4, 5).reshape(B * f * nW, TP * W, d)


def window_reverse(xw: torch.Tensor, W: int, TP: int, F: int, K: int) -> torch.Tensor:
    """Inverse of window_partition (HWGATE.py:39-47)."""
    f, nW = F // TP, K // W
    d = xw.shape[-1]
    B = xw.shape[0] // (f * nW)
    return xw.reshape(B, f, nW, TP, W, d).permute(0, 1, 3, 2, 4, 5).reshape(B, F, K, d)


def temporal_merge(x: torch.Tensor, TP: int) -> torch.Tensor:
    """out[b,fi,k,tp*d+e] = x[b,fi*TP+tp,k,e] (HWGATE.py:55-63)."""
    B, F, K, d = x.shape
    return x.reshape(B, F // TP, TP, K, d).permute(0, 1, 3, 2, 4).reshape(B, F // TP, K, TP * d)


def temporal_merge_backward(g: torch.Tensor, TP: int) -> torch.Tensor:
    """Adjoint of temporal_merge: gradient w.r.t. its input."""
    B, f, K, D = g.shape
    d = D // TP
    return g.reshape(B, f, K, TP, d).permute(0, 1, 3, 2, 4).reshape(B, f * TP, K, d)


# --------------------------------------------------------------------------
# rounding model of the bf16 kernels (not in the reference: the reference has
# no bf16 path; see DESIGN.md "bf16 definition")
# --------------------------------------------------------------------------

def _rb(t: torch.Tensor, on: bool) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype) if on else t


# --------------------------------------------------------------------------
# attention core = what K2/K3 compute (rows a5-a10 minus the output proj)
# --------------------------------------------------------------------------

def attention_core(xn: torch.Tensor, w_qkv: torch.Tensor, b_qkv: torch.Tensor, heads: int,
                   mask: np.ndarray, W: int, TP: int, shift: int,
                   threshold: Optional[float] = None, bf16_points: bool = False,
                   drop_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Windowed graph attention of one block, from the normalised residual
    stream ``xn`` (B,F,K,d) to the head-merged context (B,F,K,d) *before* the
    output projection, in the un-rolled, un-partitioned layout.

    Restates: roll -shift (HWGATE.py:197-200), window_partition (201), QKV
    (86-89), logits (91), the train-only threshold drop (94-100), the two
    multiplicative masks (102-108), fill and softmax (110-111), P.V and head
    merge (114), window_reverse (207), roll +shift (210-215).  ``mask`` is
    combined_mask(...): (f*nW, N, N) bool.  ``threshold`` None = eval mode.

    bf16_points=True rounds at the points where the CUDA bf16 kernels round
    (xn, weights, q, k, v, P, O), so it can be compared tightly with them.
    drop_mask: (B_, heads, N, N) keep / (1-p) factors of self.attn_drop
    (HWGATE.py:112), applied to the probabilities; None = p 0 / eval.
    """
    B, F, K, d = xn.shape
    hd = d // heads
    f, nW, N = F // TP, K // W, TP * W
    dt = xn.dtype
    xs = torch.roll(xn, shifts=-shift, dims=1) if shift > 0 else xn
    xw = window_partition(_rb(xs, bf16_points), W, TP)                  # (B_, N, d)
    qkv = xw @ _rb(w_qkv.to(dt), bf16_points).t() + b_qkv.to(dt)         # (B_, N, 3d)
    qkv = qkv.reshape(-1, N, 3, heads, hd).permute(2, 0, 3, 1, 4)        # (3, B_, h, N, hd)
    q = _rb(qkv[0] * (hd ** -0.5), bf16_points)
    k = _rb(qkv[1], bf16_points)
    v = _rb(qkv[2], bf16_points)
    s = q @ k.transpose(-1, -2)                                         # (B_, h, N, N)
    m = torch.from_numpy(mask).to(device=s.device, dtype=torch.bool)    # (f*nW, N, N)
    m = m.unsqueeze(0).unsqueeze(2).expand(B, f * nW, heads, N, N).reshape(B * f * nW, heads, N, N)
    if threshold is not None:
        p0 = torch.softmax(s.detach(), dim=-1)
        m = m & ~(p0 > threshold)
    live = m & (s != 0)
    s = torch.where(live, s, torch.full_like(s, NEG_FILL))
    p = torch.softmax(s, dim=-1)
    if drop_mask is not None:
        p = p * drop_mask.to(p.dtype)                                    # HWGATE.py:112
    o = _rb(p, bf16_points) @ v                                          # (B_, h, N, hd)
    o = _rb(o.transpose(1, 2).reshape(-1, N, d), bf16_points)
    o = window_reverse(o, W, TP, F, K)
    return torch.roll(o, shifts=shift, dims=1) if shift > 0 else o


def attention_core_backward(xn, w_qkv, b_qkv, heads, mask, W, TP, shift, threshold, d_out):
    """Closed-form gradients of attention_core w.r.t. xn, w_qkv, b_qkv
    (what K3 computes; SURVEY.md section 8 row a9:
    dS = live ? P*(dP - sum_j P*dP) : 0).  Independent of autograd so the two
    can be checked against each other in tests/test_oracle_golden.py."""
    B, F, K, d = xn.shape
    hd = d // heads
    f, nW, N = F // TP, K // W, TP * W
    scale = hd ** -0.5
    xs = torch.roll(xn, shifts=-shift, dims=1) if shift > 0 else xn
    gs = torch.roll(d_out, shifts=-shift, dims=1) if shift > 0 else d_out
    xw = window_partition(xs, W, TP)
    gw = window_partition(gs, W, TP).reshape(-1, N, heads, hd).transpose(1, 2)   # (B_, h, N, hd)
    qkv = (xw @ w_qkv.t() + b_qkv).reshape(-1, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    s = q @ k.transpose(-1, -2)
    m = torch.from_numpy(mask).to(device=s.device, dtype=torch.bool)
    m = m.unsqueeze(0).unsqueeze(2).expand(B, f * nW, heads, N, N).reshape(B * f * nW, heads, N, N)
    if threshold is not None:
        m = m & ~(torch.softmax(s, dim=-1) > threshold)
    live = m & (s != 0)
    p = torch.softmax(torch.where(live, s, torch.full_like(s, NEG_FILL)), dim=-1)
    dv = p.transpose(-1, -2) @ gw
    dp = gw @ v.transpose(-1, -2)
    ds = p * (dp - (p * dp).sum(-1, keepdim=True))
    ds = torch.where(live, ds, torch.zeros_like(ds))
    dq = (ds @ k) * scale
    dk = ds.transpose(-1, -2) @ q
    dqkv = torch.stack([dq, dk, dv], 0).permute(1, 3, 0, 2, 4).reshape(-1, N, 3 * d)  # (B_, N, 3d)
    d_w = dqkv.reshape(-1, 3 * d).t() @ xw.reshape(-1, d)
    d_b = dqkv.reshape(-1, 3 * d).sum(0)
    d_xw = dqkv @ w_qkv
    d_x = window_reverse(d_xw, W, TP, F, K)
    d_x = torch.roll(d_x, shifts=shift, dims=1) if shift > 0 else d_x
    return d_x, d_w, d_b


# --------------------------------------------------------------------------
# the rest of the block and the model (rows a11-a14), needed so the end-to-end
# fwd+bwd metric has a reference on CPU
# --------------------------------------------------------------------------

def _layer_norm(x, w, b, eps=1e-5):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w.to(x.dtype), b.to(x.dtype), eps)


def _dropout(x, p, training):
    return torch.nn.functional.dropout(x, p, training) if (training and p > 0) else x


def sinusoid_table(T: int, d: int, dtype=torch.float32) -> torch.Tensor:
    """(1,T,1,d) sin/cos table over frames (HWGATE.py:15-23)."""
    pos = torch.arange(T, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(T, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.reshape(1, T, 1, d).to(dtype)


def block_forward(x, sd: Dict[str, torch.Tensor], prefix: str, heads: int, mask: np.ndarray,
                  W: int, TP: int, shift: int, threshold, drop: float, training: bool,
                  bf16_points: bool = False):
    """x + proj(attention(LN1 x)) then x + FFN(LN2 x)  (HWGATE.py:189-221)."""
    g = lambda n: sd[prefix + n].to(x.dtype)
    xn = _layer_norm(x, g("norm1.weight"), g("norm1.bias"))
    ctx = attention_core(xn, g("attn.qkv.weight"), g("attn.qkv.bias"), heads, mask, W, TP,
                         shift, threshold, bf16_points)
    a = ctx @ g("attn.proj.weight").t() + g("attn.proj.bias")            # HWGATE.py:115
    x = x + _dropout(a, drop, training)                                  # 116, 217
    h = _layer_norm(x, g("norm2.weight"), g("norm2.bias"))
    h = torch.nn.functional.gelu(h @ g("ff.fc1.weight").t() + g("ff.fc1.bias"))   # 131-132
    h = _dropout(h, drop, training)
    h = h @ g("ff.fc2.weight").t() + g("ff.fc2.bias")
    return x + _dropout(h, drop, training)                               # 219


def block_plan(cfg: HWGATEConfig):
    """[(prefix, level, heads, frames, shift)] in execution order
    (HWGATE.py:236-246, 306-325)."""
    plan = []
    for i, depth in enumerate(cfg.depths):
        for j in range(depth):
            shift = 0 if j % 2 == 0 else cfg.temporal_patch_size // 2
            plan.append((f"layers.{i}.blocks.{j}.", i, cfg.num_heads[i], cfg.level_frames(i), shift))
    return plan


def model_forward(x, sd: Dict[str, torch.Tensor], cfg: HWGATEConfig,
                  thresholds: Optional[List[float]] = None, drop: float = 0.0,
                  bf16_points: bool = False):
    """(B,T,64,C) keypoints -> (B,num_classes) logits (HWGATE.py:342-360).
    ``thresholds``: one scalar per block in execution order = training mode
    (the reference draws them from the CPU RNG, HWGATE.py:96); None = eval."""
    training = thresholds is not None
    W, TP = cfg.window_size, cfg.temporal_patch_size
    adj = window_adjacency(cfg.edges, W, TP)
    dt = x.dtype
    xp = (2.0 * math.pi * x) @ sd["B"].to(dt).t()                        # 343
    h = torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)                # 344
    if cfg.pe:
        h = _dropout(h + sd["pos_encoder.pe"].to(dt)[:, :h.shape[1]], drop, training)   # 25-28
    plan = block_plan(cfg)
    for bi, (prefix, lvl, heads, frames, shift) in enumerate(plan):
        m = combined_mask(adj, frames, W, TP, shift)
        thr = thresholds[bi] if training else None
        h = block_forward(h, sd, prefix, heads, m, W, TP, shift, thr, drop, training, bf16_points)
        last_in_level = bi + 1 == len(plan) or plan[bi + 1][1] != lvl
        if last_in_level and lvl < len(cfg.depths) - 1:
            h = temporal_merge(h, TP)                                    # 256-257
    h = _layer_norm(h, sd["norm.weight"], sd["norm.bias"])               # 353
    h = h.mean(dim=(1, 2))                                               # 354 (avg over f*K tokens)
    return h @ sd["head.weight"].to(dt).t() + sd["head.bias"].to(dt)     # 359


def smoothed_cross_entropy(logits, target, smooth: float = 0.01):
    """hwgat/losses/SmoothCrossEntropy.py:35-39."""
    lp = torch.log_softmax(logits, dim=-1)
    nll = -lp.gather(-1, target.unsqueeze(1)).squeeze(1)
    return ((1.0 - smooth) * nll + smooth * (-lp.mean(-1))).mean()


# --------------------------------------------------------------------------
# deterministic parameters and inputs (numpy PCG64, so fixtures are a seed,
# not a tensor dump)
# --------------------------------------------------------------------------

def make_state_dict(cfg: HWGATEConfig, seed: int, weight_std: float = 0.02,
                    dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """A state_dict with exactly the reference's names and shapes (SURVEY.md
    section 5 'Checkpoint').  Linear weights ~ N(0, weight_std); biases and LN
    offsets are small non-zero values so that they are exercised."""
    rng = np.random.default_rng(seed)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    nrm = lambda *s, std=weight_std: T(rng.standard_normal(s) * std)
    sd: Dict[str, torch.Tensor] = {}
    sd["B"] = T(rng.standard_normal((cfg.embed_dim // 2, cfg.kp_dim)) * 10.0)    # HWGATE.py:297-299
    if cfg.pe:
        sd["pos_encoder.pe"] = sinusoid_table(cfg.temporal_dim, cfg.embed_dim, dtype)
    W, TP = cfg.window_size, cfg.temporal_patch_size
    for prefix, lvl, heads, frames, shift in block_plan(cfg):
        d = cfg.level_dim(lvl)
        hid = int(d * cfg.ff_ratio)
        if shift > 0:                                                    # buffer, HWGATE.py:187
            sm = shift_window_mask(frames, cfg.n_windows, W, TP, shift)
            sd[prefix + "attn_mask"] = T(sm.astype(np.float32))
        sd[prefix + "norm1.weight"] = 1.0 + nrm(d, std=0.1)
        sd[prefix + "norm1.bias"] = nrm(d, std=0.05)
        sd[prefix + "attn.qkv.weight"] = nrm(3 * d, d)
        sd[prefix + "attn.qkv.bias"] = nrm(3 * d, std=0.05)
        sd[prefix + "attn.proj.weight"] = nrm(d, d)
        sd[prefix + "attn.proj.bias"] = nrm(d, std=0.05)
        sd[prefix + "norm2.weight"] = 1.0 + nrm(d, std=0.1)
        sd[prefix + "norm2.bias"] = nrm(d, std=0.05)
        sd[prefix + "ff.fc1.weight"] = nrm(hid, d)
        sd[prefix + "ff.fc1.bias"] = nrm(hid, std=0.05)
        sd[prefix + "ff.fc2.weight"] = nrm(d, hid)
        sd[prefix + "ff.fc2.bias"] = nrm(d, std=0.05)
    dl = cfg.level_dim(len(cfg.depths) - 1)
    sd["norm.weight"] = 1.0 + nrm(dl, std=0.1)
    sd["norm.bias"] = nrm(dl, std=0.05)
    sd["head.weight"] = nrm(cfg.num_classes, dl)
    sd["head.bias"] = nrm(cfg.num_classes, std=0.05)
    return sd


# WindowCreate table, hwgat/dataTransform.py:428-441 (29 keypoints -> 4 x 16)
_HEAD, _LARM, _RARM = [0, 1, 2], [3, 5, 7], [4, 6, 8]
_LHAND, _RHAND = list(range(9, 19)), list(range(19, 29))
WINDOW_GATHER = np.array(_HEAD + _LARM + _LHAND + _HEAD + _RARM + _RHAND +
                         _HEAD + _LARM + _RHAND + _HEAD + _RARM + _LHAND, dtype=np.int64)


def synthetic_keypoints(B: int, T: int, C: int = 2, seed: int = 1001):
    """U(0,1) raw keypoints (B,T,29,C) gathered to the 64-slot window layout;
    SURVEY.md section 8(d) 'Synthetic input'.  Returns float32 (B,T,64,C)."""
    rng = np.random.default_rng(seed)
    raw = rng.random((B, T, 29, C), dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(raw[:, :, WINDOW_GATHER, :]))


def synthetic_labels(B: int, classes: int, seed: int = 1001):
    rng = np.random.default_rng(seed + 7)
    return torch.from_numpy(rng.integers(0, classes, size=(B,), dtype=np.int64))
