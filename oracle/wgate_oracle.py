"""CPU oracle of the sibling models WGATE (hwgat/models/WGATE.py) and GATE (hwgat/models/GATE.py) - TEST
INFRASTRUCTURE ONLY.

Same contract as oracle/hwgate_oracle.py: a functional restatement of the reference's algorithm over a flat state_dict,
every function citing the reference lines it follows; pinned to outputs of the unmodified reference by
tests/golden/wgate_gate.npz (tests/golden/make_golden.py section 8, checked in tests/test_oracle_golden.py).  Nothing
under sl_hwgat_b200/ may import it.

Both models attend DENSELY over all frames with the graph as an additive mask, and so does this oracle: WGATE inside
each window of 16 keypoints x all frames (N = F*16, WGATE.py:32-45, 87-106), GATE over all 29 keypoints x all frames
(N = F*29, GATE.py:49-66).  The product kernels evaluate only the graph's 3-frame band; that the two agree is exactly
what the parity tests establish.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Sequence

import numpy as np
import torch

from oracle.hgate_oracle import HGATE_EDGES
from oracle.hwgate_oracle import NEG_FILL, _WINDOW_EDGES, _dropout, _layer_norm, _rb, sinusoid_table


def frame_band(same: np.ndarray, frames: int) -> np.ndarray:
    """(F*k, F*k): `same` on the diagonal frame blocks, identity on the two adjacent block diagonals, zero elsewhere
    (the nested loops of model_params.py:208-224 / 62-72)."""
    k = same.shape[0]
    fr = np.arange(frames)
    dt = np.abs(fr[:, None] - fr[None, :])
    blocks = np.where((dt == 0)[:, None, :, None], same[None, :, None, :],
                      np.where((dt == 1)[:, None, :, None], np.eye(k)[None, :, None, :], 0.0))
    return blocks.reshape(frames * k, frames * k)


def wgate_adjacency(edges, frames: int, W: int) -> np.ndarray:
    """(nW, F*W, F*W) float (model_params.py:204-239): eye(W) + the window's edges inside a frame."""
    out = []
    for win in edges:
        a = np.eye(W)
        for i, j in win:
            a[i, j] = 1
            a[j, i] = 1
        out.append(frame_band(a, frames))
    return np.stack(out)


def gate_adjacency(edges, frames: int, K: int) -> np.ndarray:
    """(F*K, F*K) float (model_params.py:59-74): the edges inside a frame WITHOUT self loops + temporal links."""
    a = np.zeros((K, K))
    for i, j in edges:
        a[i, j] = 1
        a[j, i] = 1
    return frame_band(a, frames)


def additive_mask(adj: np.ndarray) -> np.ndarray:
    """0 where adj == 1, -10000 where adj == 0 (WGATE.py:190, GATE.py:142)."""
    return np.where(adj == 0, NEG_FILL, np.where(adj == 1, 0.0, adj))


@dataclass
class WGATEConfig:
    """Hyper-parameters, defaults = WGATEParams (model_params.py:80-97)."""
    kp_dim: int = 2
    num_kps: int = 64
    temporal_dim: int = 64
    num_classes: int = 262
    embed_dim: int = 128
    pe: bool = True
    depths: int = 8
    num_heads: int = 8
    window_size: int = 16
    ff_ratio: float = 2.0
    edges: Sequence = field(default_factory=lambda: [[list(e) for e in _WINDOW_EDGES] for _ in range(4)])


@dataclass
class GATEConfig:
    """Hyper-parameters, defaults = GATEParams (model_params.py:5-20)."""
    kp_dim: int = 2
    num_kps: int = 29
    temporal_dim: int = 64
    num_classes: int = 262
    embed_dim: int = 128
    pe: bool = True
    depths: int = 8
    num_heads: int = 8
    ff_ratio: float = 2.0
    edges: Sequence = field(default_factory=lambda: [list(e) for e in HGATE_EDGES])


def dense_attention(x: torch.Tensor, w_qkv, b_qkv, heads: int, mask: torch.Tensor, bf16_points: bool = False):
    """MSA.forward up to the head merge (WGATE.py:87-106 / GATE.py:49-66): x (G, N, d) normalised tokens of G
    sequences, mask broadcastable to (G, heads, N, N), additive."""
    G, N, d = x.shape
    hd = d // heads
    dt = x.dtype
    qkv = _rb(x, bf16_points) @ _rb(w_qkv.to(dt), bf16_points).t() + b_qkv.to(dt)
    qkv = _rb(qkv, bf16_points).reshape(G, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q = _rb(qkv[0] * (hd ** -0.5), bf16_points)
    s = q @ qkv[1].transpose(-1, -2) + mask.to(dt)
    p = torch.softmax(s, dim=-1)
    o = _rb(p, bf16_points) @ qkv[2]
    return _rb(o.transpose(1, 2).reshape(G, N, d), bf16_points)


def wgate_attention_core(xn: torch.Tensor, w_qkv, b_qkv, heads: int, mask: np.ndarray, W: int,
                         bf16_points: bool = False) -> torch.Tensor:
    """(B,F,K,d) normalised stream -> head-merged context (B,F,K,d): window_partition (WGATE.py:32-45), MSA with
    the (nW, N, N) additive mask shared by the batch (WGATE.py:102-104), window_reverse (WGATE.py:49-66).  norm1 is
    per token, so applying it before the partition (as the caller does) equals WGATE.py:156."""
    B, F, K, d = xn.shape
    nW = K // W
    xw = xn.reshape(B, F, nW, W, d).transpose(1, 2).reshape(B * nW, F * W, d)
    m = torch.from_numpy(np.asarray(mask)).to(xn.device)                       # (nW, N, N)
    m = m[None, :, None].expand(B, nW, 1, F * W, F * W).reshape(B * nW, 1, F * W, F * W)
    o = dense_attention(xw, w_qkv, b_qkv, heads, m, bf16_points)
    return o.reshape(B, nW, F, W, d).transpose(1, 2).reshape(B, F, K, d)


def gate_attention_core(xn: torch.Tensor, w_qkv, b_qkv, heads: int, mask: np.ndarray, bf16_points: bool = False):
    """(B,F,K,d) normalised stream -> context (B,F,K,d): GATE flattens to (B, F*K, d) (GATE.py:198) and attends over
    everything with the (1,1,N,N) additive mask (GATE.py:60-62)."""
    B, F, K, d = xn.shape
    m = torch.from_numpy(np.asarray(mask)).to(xn.device).reshape(1, 1, F * K, F * K)
    return dense_attention(xn.reshape(B, F * K, d), w_qkv, b_qkv, heads, m, bf16_points).reshape(B, F, K, d)


def _embed(x, sd, pe: bool, drop: float, training: bool):
    dt = x.dtype
    xp = (2.0 * math.pi * x) @ sd["B"].to(dt).t()
    h = torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)
    if pe:
        h = _dropout(h + sd["pos_encoder.pe"].to(dt)[:, :h.shape[1]], drop, training)
    return h


def _block(h, sd, prefix, core, drop, training):
    g = lambda n: sd[prefix + n].to(h.dtype)
    xn = _layer_norm(h, g("norm1.weight"), g("norm1.bias"))
    a = core(xn, g("attn.qkv.weight"), g("attn.qkv.bias")) @ g("attn.proj.weight").t() + g("attn.proj.bias")
    h = h + _dropout(a, drop, training)
    u = _layer_norm(h, g("norm2.weight"), g("norm2.bias"))
    u = torch.nn.functional.gelu(u @ g("ff.fc1.weight").t() + g("ff.fc1.bias"))
    u = _dropout(u, drop, training)
    u = u @ g("ff.fc2.weight").t() + g("ff.fc2.bias")
    return h + _dropout(u, drop, training)


def wgate_forward(x, sd: Dict[str, torch.Tensor], cfg: WGATEConfig, drop: float = 0.0, training: bool = False,
                  bf16_points: bool = False):
    """(B,T,64,C) keypoints -> (B,num_classes) logits (WGATE.py:243-263)."""
    mask = sd["adj_mask"].cpu().numpy()
    h = _embed(x, sd, cfg.pe, drop, training)
    for i in range(cfg.depths):
        core = lambda xn, w, b: wgate_attention_core(xn, w, b, cfg.num_heads, mask, cfg.window_size, bf16_points)
        h = _block(h, sd, f"layers.{i}.", core, drop, training)
    h = _layer_norm(h, sd["norm.weight"].to(h.dtype), sd["norm.bias"].to(h.dtype))
    h = h.mean(dim=(1, 2))                                                   # AvgPool1d over all F*K tokens
    return h @ sd["head.weight"].to(h.dtype).t() + sd["head.bias"].to(h.dtype)


def gate_forward(x, sd: Dict[str, torch.Tensor], cfg: GATEConfig, drop: float = 0.0, training: bool = False,
                 bf16_points: bool = False):
    """(B,T,29,C) keypoints -> (B,num_classes) logits (GATE.py:188-216): the pool is the learned `weightedAvg`
    Linear(F*K, 1) over the token axis (GATE.py:207)."""
    mask = sd["adj_mask"].cpu().numpy()
    h = _embed(x, sd, cfg.pe, drop, training)
    for i in range(cfg.depths):
        core = lambda xn, w, b: gate_attention_core(xn, w, b, cfg.num_heads, mask, bf16_points)
        h = _block(h, sd, f"layers.{i}.", core, drop, training)
    h = _layer_norm(h, sd["norm.weight"].to(h.dtype), sd["norm.bias"].to(h.dtype))
    B, F, K, d = h.shape
    wa = sd["weightedAvg.weight"].to(h.dtype).reshape(F * K)
    h = torch.einsum("btd,t->bd", h.reshape(B, F * K, d), wa) + sd["weightedAvg.bias"].to(h.dtype)
    return h @ sd["head.weight"].to(h.dtype).t() + sd["head.bias"].to(h.dtype)


def make_state_dict(cfg, seed: int, weight_std: float = 0.02, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """A state_dict with the reference's names and shapes (WGATE.py:162-241 / GATE.py:118-186)."""
    rng = np.random.default_rng(seed)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    nrm = lambda *s, std=weight_std: T(rng.standard_normal(s) * std)
    gate = isinstance(cfg, GATEConfig)
    d, F, K = cfg.embed_dim, cfg.temporal_dim, cfg.num_kps
    if gate:
        sd = {"adj_mask": T(additive_mask(gate_adjacency(cfg.edges, F, K)))[None, None]}
    else:
        sd = {"adj_mask": T(additive_mask(wgate_adjacency(cfg.edges, F, cfg.window_size)))}
    sd["B"] = T(rng.standard_normal((d // 2, cfg.kp_dim)) * 10.0)
    if cfg.pe:
        sd["pos_encoder.pe"] = sinusoid_table(F, d, dtype)
    hid = int(d * cfg.ff_ratio)
    for i in range(cfg.depths):
        prefix = f"layers.{i}."
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
    sd["norm.weight"] = 1.0 + nrm(d, std=0.1)
    sd["norm.bias"] = nrm(d, std=0.05)
    if gate:
        sd["weightedAvg.weight"] = T(np.full((1, F * K), 1.0 / (F * K)) + rng.standard_normal((1, F * K)) * 0.2 / (F * K))
        sd["weightedAvg.bias"] = nrm(1, std=0.05)
    sd["head.weight"] = nrm(cfg.num_classes, d)
    sd["head.bias"] = nrm(cfg.num_classes, std=0.05)
    return sd


def synthetic_keypoints(B: int, T: int, K: int, C: int = 2, seed: int = 1001):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.random((B, T, K, C), dtype=np.float32))
