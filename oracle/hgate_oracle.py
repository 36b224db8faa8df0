"""CPU oracle of the sibling model HGATE (hwgat/models/HGATE.py) - TEST INFRASTRUCTURE ONLY.

Same contract as oracle/hwgate_oracle.py: a functional restatement of the reference's algorithm over a flat
state_dict, every function citing the reference lines it follows; pinned to outputs of the unmodified reference by
tests/golden/hgate.npz (tests/golden/make_golden.py section 7, checked in tests/test_oracle_golden.py).  Nothing under
sl_hwgat_b200/ may import it.

HGATE is HWGATE without keypoint windows: a block is ALL 29 keypoints of TP = 2 consecutive frames (58 tokens,
HGATE.py:30-36), the attention mask is one (58, 58) skeleton adjacency (model_params.py:459-481) times a per-temporal-
group shift mask (HGATE.py:155-171), both multiplicative with the -10000 fill (HGATE.py:93-102), and there is no
training-time threshold drop.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from oracle.hwgate_oracle import (NEG_FILL, _dropout, _layer_norm, _rb, frame_group_ids, sinusoid_table,
                                  temporal_merge)


def _hand(base: int):
    """12 edges of one 10-keypoint hand whose wrist is keypoint `base` (model_params.py:431-443 / 445-456)."""
    fingers = [base + 2 * i for i in range(1, 5)]
    return ([(base, base + 1)] + [(base, f) for f in fingers] + [(f, f + 1) for f in fingers] +
            [(fingers[i], fingers[i + 1]) for i in range(3)])


# 29 mediapipe keypoints: 0-2 head, 3-8 shoulders / elbows / wrists, 9-18 left hand, 19-28 right hand; 34 edges
# (values restated from model_params.py:422-457)
HGATE_EDGES = ([(2, 0), (1, 0), (0, 3), (0, 4), (3, 5), (4, 6), (5, 7), (6, 8), (7, 9)] + _hand(9) + [(8, 19)] + _hand(19))


@dataclass
class HGATEConfig:
    """Hyper-parameters, defaults = HGATEParams (model_params.py:405-420)."""
    kp_dim: int = 2
    num_kps: int = 29
    temporal_dim: int = 64
    num_classes: int = 262
    embed_dim: int = 128
    temporal_patch_size: int = 2
    pe: bool = True
    depths: Sequence[int] = (2, 2, 4)
    num_heads: Sequence[int] = (2, 4, 8)
    ff_ratio: float = 2.0
    edges: Sequence[Sequence[int]] = field(default_factory=lambda: [list(e) for e in HGATE_EDGES])

    def level_dim(self, i: int) -> int:
        return int(self.embed_dim * 2 ** i)

    def level_frames(self, i: int) -> int:
        return self.temporal_dim // self.temporal_patch_size ** i


def block_adjacency(edges, K: int, TP: int) -> np.ndarray:
    """(TP*K, TP*K) bool: skeleton with self loops inside a frame, identity between adjacent frames, nothing further
    apart (model_params.py:459-481)."""
    a = np.eye(K, dtype=bool)
    for i, j in edges:
        a[i, j] = True
        a[j, i] = True
    N = TP * K
    tp, kp = np.arange(N) // K, np.arange(N) % K
    dt = np.abs(tp[:, None] - tp[None, :])
    return np.where(dt == 0, a[kp[:, None], kp[None, :]], np.where(dt == 1, kp[:, None] == kp[None, :], False))


def block_shift_mask(F: int, K: int, TP: int, shift: int) -> Optional[np.ndarray]:
    """(f, TP*K, TP*K) bool, True where query and key frames share a group id after the cyclic shift; None for
    unshifted blocks (HGATE.py:155-171; group ids as HWGATE's, HGATE.py:159-166)."""
    if shift <= 0:
        return None
    f, N = F // TP, TP * K
    g = frame_group_ids(F, TP, shift)
    gid = g[np.arange(f)[:, None] * TP + (np.arange(N) // K)[None, :]]          # (f, N)
    return gid[:, :, None] == gid[:, None, :]


def block_mask(adj: np.ndarray, F: int, K: int, TP: int, shift: int) -> np.ndarray:
    """(f, N, N) bool: the product of the two multiplicative masks of MSA.forward (HGATE.py:93-99)."""
    f = F // TP
    m = np.tile(adj[None], (f, 1, 1))
    sm = block_shift_mask(F, K, TP, shift)
    return m & sm if sm is not None else m


def attention_core(xn: torch.Tensor, w_qkv: torch.Tensor, b_qkv: torch.Tensor, heads: int, mask: np.ndarray, TP: int,
                   shift: int, bf16_points: bool = False) -> torch.Tensor:
    """Block attention of one GraphAttentionBlock from the normalised stream xn (B,F,K,d) to the head-merged context
    (B,F,K,d) before the output projection: roll (HGATE.py:189-192), block_partition (30-36, 194), QKV (85-89), logits
    (91), the two masks (93-99), fill + softmax (101-102), P.V (105), block_reverse (39-47, 200), roll back (203-206).
    norm1 is per token, so applying it before the partition (as the caller does) equals HGATE.py:196."""
    B, F, K, d = xn.shape
    hd, f, N = d // heads, F // TP, TP * K
    dt = xn.dtype
    xs = torch.roll(xn, shifts=-shift, dims=1) if shift > 0 else xn
    xb = _rb(xs, bf16_points).reshape(B * f, N, d)
    qkv = xb @ _rb(w_qkv.to(dt), bf16_points).t() + b_qkv.to(dt)
    qkv = qkv.reshape(-1, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q = _rb(qkv[0] * (hd ** -0.5), bf16_points)
    k = _rb(qkv[1], bf16_points)
    v = _rb(qkv[2], bf16_points)
    s = q @ k.transpose(-1, -2)                                             # (B*f, h, N, N)
    m = torch.from_numpy(mask).to(device=s.device, dtype=torch.bool)       # (f, N, N)
    m = m.unsqueeze(0).unsqueeze(2).expand(B, f, heads, N, N).reshape(B * f, heads, N, N)
    live = m & (s != 0)
    p = torch.softmax(torch.where(live, s, torch.full_like(s, NEG_FILL)), dim=-1)
    o = _rb(p, bf16_points) @ v
    o = _rb(o.transpose(1, 2).reshape(B, F, K, d), bf16_points)
    return torch.roll(o, shifts=shift, dims=1) if shift > 0 else o


def block_plan(cfg: HGATEConfig):
    plan = []
    for i, depth in enumerate(cfg.depths):
        for j in range(depth):
            shift = 0 if j % 2 == 0 else cfg.temporal_patch_size // 2
            plan.append((f"layers.{i}.blocks.{j}.", i, cfg.num_heads[i], cfg.level_frames(i), shift))
    return plan


def model_forward(x, sd: Dict[str, torch.Tensor], cfg: HGATEConfig, drop: float = 0.0, training: bool = False,
                  bf16_points: bool = False):
    """(B,T,29,C) keypoints -> (B,num_classes) logits (HGATE.py:327-346)."""
    TP, K = cfg.temporal_patch_size, cfg.num_kps
    adj = block_adjacency(cfg.edges, K, TP)
    dt = x.dtype
    xp = (2.0 * math.pi * x) @ sd["B"].to(dt).t()
    h = torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)
    if cfg.pe:
        h = _dropout(h + sd["pos_encoder.pe"].to(dt)[:, :h.shape[1]], drop, training)
    plan = block_plan(cfg)
    for bi, (prefix, lvl, heads, frames, shift) in enumerate(plan):
        g = lambda n: sd[prefix + n].to(h.dtype)
        m = block_mask(adj, frames, K, TP, shift)
        xn = _layer_norm(h, g("norm1.weight"), g("norm1.bias"))
        ctx = attention_core(xn, g("attn.qkv.weight"), g("attn.qkv.bias"), heads, m, TP, shift, bf16_points)
        a = ctx @ g("attn.proj.weight").t() + g("attn.proj.bias")
        h = h + _dropout(a, drop, training)
        u = _layer_norm(h, g("norm2.weight"), g("norm2.bias"))
        u = torch.nn.functional.gelu(u @ g("ff.fc1.weight").t() + g("ff.fc1.bias"))
        u = _dropout(u, drop, training)
        u = u @ g("ff.fc2.weight").t() + g("ff.fc2.bias")
        h = h + _dropout(u, drop, training)
        last_in_level = bi + 1 == len(plan) or plan[bi + 1][1] != lvl
        if last_in_level and lvl < len(cfg.depths) - 1:
            h = temporal_merge(h, TP)
    h = _layer_norm(h, sd["norm.weight"], sd["norm.bias"])
    h = h.mean(dim=(1, 2))
    return h @ sd["head.weight"].to(dt).t() + sd["head.bias"].to(dt)


def make_state_dict(cfg: HGATEConfig, seed: int, weight_std: float = 0.02, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """A state_dict with the reference HGATE's names and shapes (HGATE.py:241-313)."""
    rng = np.random.default_rng(seed)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    nrm = lambda *s, std=weight_std: T(rng.standard_normal(s) * std)
    sd: Dict[str, torch.Tensor] = {"B": T(rng.standard_normal((cfg.embed_dim // 2, cfg.kp_dim)) * 10.0)}
    if cfg.pe:
        sd["pos_encoder.pe"] = sinusoid_table(cfg.temporal_dim, cfg.embed_dim, dtype)
    TP, K = cfg.temporal_patch_size, cfg.num_kps
    for prefix, lvl, heads, frames, shift in block_plan(cfg):
        d = cfg.level_dim(lvl)
        hid = int(d * cfg.ff_ratio)
        sd[prefix + "norm1.weight"] = 1.0 + nrm(d, std=0.1)
        sd[prefix + "norm1.bias"] = nrm(d, std=0.05)
        sd[prefix + "norm2.weight"] = 1.0 + nrm(d, std=0.1)
        sd[prefix + "norm2.bias"] = nrm(d, std=0.05)
        sd[prefix + "ff.fc1.weight"] = nrm(hid, d)
        sd[prefix + "ff.fc1.bias"] = nrm(hid, std=0.05)
        sd[prefix + "ff.fc2.weight"] = nrm(d, hid)
        sd[prefix + "ff.fc2.bias"] = nrm(d, std=0.05)
        if shift > 0:                                                    # buffer, HGATE.py:173
            sd[prefix + "attn_mask"] = T(block_shift_mask(frames, K, TP, shift).astype(np.float32))
        sd[prefix + "attn.qkv.weight"] = nrm(3 * d, d)
        sd[prefix + "attn.qkv.bias"] = nrm(3 * d, std=0.05)
        sd[prefix + "attn.proj.weight"] = nrm(d, d)
        sd[prefix + "attn.proj.bias"] = nrm(d, std=0.05)
    dl = cfg.level_dim(len(cfg.depths) - 1)
    sd["norm.weight"] = 1.0 + nrm(dl, std=0.1)
    sd["norm.bias"] = nrm(dl, std=0.05)
    sd["head.weight"] = nrm(cfg.num_classes, dl)
    sd["head.bias"] = nrm(cfg.num_classes, std=0.05)
    return sd


def synthetic_keypoints(B: int, T: int, K: int = 29, C: int = 2, seed: int = 1001):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.random((B, T, K, C), dtype=np.float32))
