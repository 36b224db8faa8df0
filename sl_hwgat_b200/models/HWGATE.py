"""B200-native HWGATE: drop-in for hwgat/models/HWGATE.py.

Same class names, constructor signatures, forward signatures and state_dict
keys as the reference, so `utils.load_model` (utils.py:55-59), checkpoints
(utils.py:185-237) and `inference.py:95` work unchanged.  What differs is what
runs: the roll / window_partition / QKV / masked attention / window_reverse /
roll-back chain of every block is ONE call into the hand-written sm_100a
kernels (sl_hwgat_b200.ops.window_graph_attention: K2 forward, K3 backward with
recompute), the masks are packed bitmasks built on the GPU (K1), and
TemporalMerging is kernel K4.  Nothing here falls back to PyTorch attention: on
a non-CUDA tensor the ops raise.

Precision: fp32 tensors take the true-fp32 kernels (1e-5 parity mode).  Inside
`torch.autocast("cuda", dtype=torch.bfloat16)` the attention runs the bf16
tensor-core kernels (activations/weights bf16, accumulation fp32), matching
what autocast does to the reference's qkv / matmul ops.
"""
import math
import os

import torch
import torch.nn as nn

from sl_hwgat_b200 import ops
from sl_hwgat_b200.ops import LAYOUT_BFKD, LAYOUT_WINDOWS


# The reference's loop calls model(x) without autocast (utils.py:102, 128; inference.py:95), which lands on the true-fp32
# parity kernels (1e-5, ~15x slower).  HWGAT_AUTOCAST=bf16 in the environment (or models.HWGATE.AUTOCAST = "bf16") makes
# Model.forward open the bf16 autocast region itself, so `python main.py ...` gets the tcgen05 path with no source change.
# Off by default: it changes the numerics from fp32 to the reference's autocast band.
AUTOCAST = os.environ.get("HWGAT_AUTOCAST", "")


def _trunc_normal_(tensor, std):
    # timm's trunc_normal_(std=.02) as used at HWGATE.py:335 == torch's, cut at +-2
    return nn.init.trunc_normal_(tensor, mean=0.0, std=std, a=-2.0, b=2.0)


def _attn_dtype(x):
    """bf16 kernels under bf16 autocast or for bf16 inputs, fp32 kernels otherwise."""
    if x.dtype == torch.bfloat16:
        return torch.bfloat16
    if torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        return torch.bfloat16
    return torch.float32


def chain_enabled(x):
    """the fused chain can run on x: fp32 CUDA tensor under bf16 autocast, or without autocast in the x3 fp32 mode"""
    return x.is_cuda and x.dtype == torch.float32 and (
        _attn_dtype(x) == torch.bfloat16 or (not torch.is_autocast_enabled() and ops.fp32_mode() == "x3"))


def linear(mod, x):
    """`mod(x)` for an nn.Linear of a block.  On the fp32 path (no autocast: the reference's own loop) in x3 mode
    (ops.set_fp32_mode / HWGAT_FP32) it runs on the library's tcgen05 fp32 GEMM (gemm_x3.cu) instead of cuBLAS' FFMA
    sgemm; bf16 autocast never comes here (K10 / K12)."""
    if (type(mod) is nn.Linear and x.is_cuda and x.dtype == torch.float32 and mod.weight.dtype == torch.float32
            and not torch.is_autocast_enabled()
            and ops.linear_x3_active(x.numel() // x.shape[-1], x.shape[-1], mod.weight.shape[0])):
        return ops.linear_f32(x, mod.weight, mod.bias)
    return mod(x)


def embed_generic(model, x):
    """Fourier embedding + positional encoding with PyTorch ops (HWGATE.py:343-347): the fp32 mode, and inputs that
    need a gradient.  The embedding stays fp32 even under autocast (bf16 would round the argument 2*pi*x.B, values up
    to ~300, by up to 2 rad)."""
    with torch.autocast(device_type=x.device.type, enabled=False):
        proj = (2. * math.pi * x.float()) @ model.B.float().t()
        x = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)
    if model.pe:
        x = model.pos_encoder(x)
    return x


class PositionalEncoding(nn.Module):
    """Sinusoid over the frame axis + dropout (HWGATE.py:8-28); buffer `pe` is (1, max_len, 1, d)."""

    def __init__(self, d_model, dropout, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pos = torch.arange(max_len, dtype=torch.float32)[:, None]
        freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
        table = torch.empty(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer('pe', table[None, :, None, :])

    def forward(self, x):
        return self.dropout(x + self.pe[:, :x.size(1)])


def window_partition(x, window_size=16, temporal_patch_size=4):
    """(B,F,K,d) -> (B*f*nW, TP*W, d) (HWGATE.py:30-36).  Kept for callers that
    want the partitioned view; the blocks below never materialise it."""
    B, F, K, d = x.shape
    f, nW = F // temporal_patch_size, K // window_size
    x = x.reshape(B, f, temporal_patch_size, nW, window_size, d).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B * f * nW, temporal_patch_size * window_size, d)


def window_reverse(x, window_size=16, temporal_patch_size=4, temporal_dim=128, num_kp=64):
    """Inverse of window_partition (HWGATE.py:39-47)."""
    f, nW = temporal_dim // temporal_patch_size, num_kp // window_size
    d = x.shape[-1]
    B = x.shape[0] // (f * nW)
    x = x.reshape(B, f, nW, temporal_patch_size, window_size, d).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, temporal_dim, num_kp, d)


class TemporalMerging(nn.Module):
    """(B,F,K,d) -> (B,F/TP,K,TP*d), no parameters (HWGATE.py:49-63): kernel K4."""

    def __init__(self, dim, temporal_patch_size):
        super().__init__()
        self.dim = dim
        self.temporal_patch_size = temporal_patch_size

    def forward(self, x):
        if self.temporal_patch_size != ops.TEMPORAL_PATCH:
            raise NotImplementedError("the merge kernel is built for temporal_patch_size == 2")
        return ops.temporal_merge(x)


class MSA(nn.Module):
    """Windowed multi-head graph attention (HWGATE.py:65-118)."""

    def __init__(self, dim, num_heads, adj_mat=None, attn_drop=0., proj_drop=0.) -> None:
        super().__init__()
        self.window_size = ops.WINDOW      # set by PartAttentionBlock (the reference's MSA infers N from its input)
        self.dim = dim
        self.num_heads = num_heads
        assert dim % num_heads == 0, 'dim and number of heads are incompatible'
        head_dim = dim // num_heads
        if head_dim != ops.HEAD_DIM:
            raise NotImplementedError("the attention kernels are built for head_dim == 64")
        # attn_drop != 0 (the reference default is 0.0, model_params.py:256) runs on K2b / K3b, the kernels that
        # have the dropout of the probabilities built in (bf16 only)
        self.scale = head_dim ** -0.5
        self.adj_mat = adj_mat
        self.qkv = nn.Linear(dim, dim * 3)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)
        self._bits = {}  # packed masks, keyed by how they were derived
        self._thr_gen = None  # see sl_hwgat_b200.parallel.sync_threshold_rng

    def _draw_threshold(self):
        # one CPU-generator scalar per call, exactly where the reference draws it (HWGATE.py:96).  Data-parallel
        # training installs a dedicated, identically seeded generator on every rank (parallel.sync_threshold_rng):
        # the global CPU generator is also consumed by data loading at rank-dependent rates.
        if not self.training:
            return None
        if self._thr_gen is not None:
            return torch.rand(1, generator=self._thr_gen).item()
        return torch.rand(1).item()

    def _project(self, ctx):
        return self.proj_drop(linear(self.proj, ctx))

    # -- fast path used by PartAttentionBlock: x is the un-partitioned (B,F,K,d) tensor
    def _attn_p(self):
        return self.attn_drop.p if self.training else 0.0

    def attend(self, xn, shift, block_bits):
        dt = _attn_dtype(xn)
        ctx = ops.window_graph_attention(xn.to(dt), self.qkv.weight, self.qkv.bias, block_bits, self.num_heads,
                                         shift=shift, threshold=self._draw_threshold(), layout=LAYOUT_BFKD,
                                         window=self.window_size, attn_drop=self._attn_p())
        return self._project(ctx)

    # -- reference signature: x is (B*f*nW, TP*W, d), already rolled and partitioned
    def forward(self, x, B, f, nW, mask=None):
        n_win, N, d = x.shape
        window = N // ops.TEMPORAL_PATCH
        if n_win != B * f * nW or N != window * ops.TEMPORAL_PATCH or window not in ops.WINDOWS:
            raise ValueError("x must be (B*f*nW, 2*window_size, d) with window_size in (16, 32, 64)")
        adj = self.adj_mat
        key = (None if mask is None else (mask.data_ptr(), mask._version),
               None if adj is None else (adj.data_ptr(), adj._version), f, nW, x.device)
        bits = self._bits.get(key)
        if bits is None:
            self._bits.clear()
            dev = x.device
            bits = ops.mask_pack(None if adj is None else adj.to(dev), None if mask is None else mask.to(dev),
                                 f * nW, N, dev)
            self._bits[key] = bits
        dt = _attn_dtype(x)
        ctx = ops.window_graph_attention(x.to(dt), self.qkv.weight, self.qkv.bias, bits, self.num_heads, shift=0,
                                         threshold=self._draw_threshold(), layout=LAYOUT_WINDOWS,
                                         frames=f * ops.TEMPORAL_PATCH, kps=nW * window, window=window,
                                         attn_drop=self._attn_p())
        return self._project(ctx)


class FeedForward(nn.Module):
    """fc1 -> act -> drop -> fc2 -> drop (HWGATE.py:120-136)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(linear(self.fc2, self.drop(self.act(linear(self.fc1, x)))))


class PartAttentionBlock(nn.Module):
    """x + MSA(LN x) over (shifted) windows, then x + FFN(LN x) (HWGATE.py:138-221)."""

    def __init__(self, dim, num_kps=64, num_heads=4, window_size=16, temporal_patch_size=4, temporal_dim=128,
                 shift_size=0, adj_mat=None, drop=0., attn_drop=0., ff_ratio=4., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm):
        super().__init__()
        if window_size not in ops.WINDOWS or temporal_patch_size != ops.TEMPORAL_PATCH:
            raise NotImplementedError("the attention kernels are built for window_size 16 / 32 / 64 and "
                                      "temporal_patch_size 2 (the multi-level reference only runs with TP = 2)")
        self.dim = dim
        self.num_kps = num_kps
        self.num_heads = num_heads
        self.window_size = window_size
        self.temporal_patch_size = temporal_patch_size
        self.temporal_dim = temporal_dim
        self.shift_size = shift_size
        self.drop = drop
        self.attn_drop = attn_drop
        self.ff_dim = dim * ff_ratio
        self.act_layer = act_layer

        self.norm1 = norm_layer(dim)
        self.attn = MSA(dim, num_heads=num_heads, adj_mat=adj_mat, attn_drop=attn_drop, proj_drop=drop)
        self.attn.window_size = window_size
        self.norm2 = norm_layer(dim)
        self.ff = FeedForward(in_features=dim, hidden_features=int(self.ff_dim), act_layer=act_layer, drop=drop)

        # `attn_mask` stays a float buffer with the reference's name and shape (f*nW, N, N)
        # so that checkpoints round-trip (HWGATE.py:169-187); the kernels use the packed
        # form built by K1 in _block_bits().
        if self.shift_size > 0:
            f, nW = temporal_dim // temporal_patch_size, num_kps // window_size
            frame = torch.arange(temporal_dim)
            group = (frame >= temporal_dim - temporal_patch_size).long() + (frame >= temporal_dim - shift_size).long()
            tok_group = group.reshape(f, temporal_patch_size, 1).expand(f, temporal_patch_size, window_size)
            tok_group = tok_group.reshape(f, 1, -1).expand(f, nW, -1).reshape(f * nW, -1)
            attn_mask = (tok_group[:, :, None] == tok_group[:, None, :]).float()
        else:
            attn_mask = None
        self.register_buffer("attn_mask", attn_mask)
        self._bits, self._bits_key = None, None

    def _block_bits(self, device):
        """Packed graph + shift mask of this block (K1), rebuilt when `attn.adj_mat` is replaced or modified in
        place (the reference multiplies by whatever tensor is there on every call, HWGATE.py:106-108)."""
        adj = self.attn.adj_mat
        key = (device, None if adj is None else (adj.data_ptr(), adj._version, tuple(adj.shape)))
        if self._bits is None or self._bits_key != key:
            nW = self.num_kps // self.window_size
            f = self.temporal_dim // self.temporal_patch_size
            N = self.window_size * self.temporal_patch_size
            if adj is None:
                adj_d = torch.ones(nW, N, N, device=device)
            else:
                adj_d = adj.to(device)
            if adj_d.shape[0] == nW or (adj_d.shape[0] == f * nW and
                                        bool((adj_d.reshape(f, nW, N, N) == adj_d[:nW]).all().item())):
                # the layer hands in the adjacency replicated over temporal groups (HWGATE.py:309);
                # K1b replicates by index, so only the first nW windows are needed
                self._bits = ops.mask_build(adj_d[:nW], self.temporal_dim, self.shift_size,
                                            self.window_size, self.temporal_patch_size)
            elif adj_d.shape[0] == f * nW:
                # a per-temporal-group adjacency: pack the float tensors as MSA.forward would see them (K1c)
                am = self.attn_mask.to(device) if self.attn_mask is not None else None
                self._bits = ops.mask_pack(adj_d, am, f * nW, N, device)
            else:
                raise ValueError(f"adj_mat has {adj_d.shape[0]} windows; expected {nW} or {f * nW}")
            self._bits_key = key
        return self._bits

    def _chain_io(self, x):
        """dtype of the activations between the kernels of the fused chain, or None when the block runs module by
        module.  bf16 under bf16 autocast.  float32 without autocast in the x3 mode of the fp32 path: the same kernels K5 - K7 with fp32 activations around the x3 tcgen05 GEMMs (gemm_x3.cu), so the
        reference's unmodified fp32 loop runs no PyTorch elementwise op inside the blocks either."""
        if x.dtype != torch.float32 or not x.is_cuda or self.dim not in (128, 256, 512):
            return None
        if not (type(self.norm1) is nn.LayerNorm and type(self.norm2) is nn.LayerNorm and isinstance(self.ff.act, nn.GELU)
                and getattr(self.ff.act, "approximate", "none") == "none"):
            return None
        n, hidden = x.numel() // self.dim, self.ff.fc1.weight.shape[0]
        if _attn_dtype(x) == torch.bfloat16:
            ok = ops.proj_supported(n, self.dim, self.dim) and ops.ffn_supported(n, self.dim, hidden)
            return torch.bfloat16 if ok else None
        if (not torch.is_autocast_enabled() and hidden in (256, 512, 1024)
                and all(type(m) is nn.Linear and m.weight.dtype == torch.float32
                        for m in (self.attn.proj, self.ff.fc1, self.ff.fc2))
                and ops.linear_x3_active(n, self.dim, self.dim) and ops.linear_x3_active(n, self.dim, hidden)
                and ops.linear_x3_active(n, hidden, self.dim)):
            return torch.float32
        return None

    def _fusable(self, x):
        """the elementwise chains around the GEMMs run as the fused kernels K5-K7 instead of PyTorch ops"""
        return self._chain_io(x) is not None

    def _check_shape(self, x):
        B, F, K, d = x.shape
        if F != self.temporal_dim or K != self.num_kps:
            raise ValueError(f"expected (B,{self.temporal_dim},{self.num_kps},d), got {tuple(x.shape)}")

    def forward(self, x):
        self._check_shape(x)
        if not self._fusable(x):
            x = x + self.attn.attend(self.norm1(x), self.shift_size, self._block_bits(x.device))
            return x + self.ff(self.norm2(x))
        x, xn = ops.layer_norm_residual(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, io=self._chain_io(x))
        return self.forward_chain(x, xn, None)[0]

    def _context(self, x, xn, bits=None):
        """the block's attention up to the head-merged context (before self.attn.proj)"""
        attn = self.attn
        return ops.window_graph_attention(xn, attn.qkv.weight, attn.qkv.bias, self._block_bits(x.device),
                                          attn.num_heads, shift=self.shift_size, threshold=attn._draw_threshold(),
                                          layout=LAYOUT_BFKD, window=self.window_size, attn_drop=attn._attn_p())

    def forward_chain(self, x, xn, next_norm, merge=False, bits=None):
        """Fused bf16 path.  x: fp32 residual stream, xn = norm1(x) in bf16 (made by the previous kernel of the
        chain).  Returns (x_out, next_norm(x_out) in bf16 or None): the LayerNorm that consumes the block's output
        is computed by the same kernel that forms the output (K6).  merge=True (last block of a level, next_norm =
        the next level's first norm1): x_out is stored directly in TemporalMerging's layout (B, F/2, K, 2d) and
        next_norm runs over the merged 2d-wide rows - K4 and its adjoint are folded into K6 / K5'."""
        attn, ff = self.attn, self.ff
        io = xn.dtype                           # bf16 (autocast) or float32 (the fp32 path in x3 mode)
        ctx = self._context(x, xn, bits)       # (the sibling models WGATE / GATE plug their banded attention in here)
        if io == torch.float32:
            # the same chain with fp32 activations: the three Linears on the x3 tcgen05 GEMM (their biases stay with
            # K6 / K7, as in the bf16 chain), fc1's bias + GELU + dropout as K7
            a0 = ops.linear_f32(ctx, attn.proj.weight, None)
            x, h = ops.bias_dropout_add_ln(x, a0, attn.proj.bias, self.norm2, attn.proj_drop.p, self.training, io=io)
            act = ops.bias_gelu_dropout(ops.linear_f32(h, ff.fc1.weight, None), ff.fc1.bias, ff.drop.p, self.training,
                                        io=io)
            v0 = ops.linear_f32(act, ff.fc2.weight, None)
        else:
            a0 = ops.output_projection(ctx, attn.proj.weight)          # K12; bias, dropout, shortcut and norm2: K6
            x, h = ops.bias_dropout_add_ln(x, a0, attn.proj.bias, self.norm2, attn.proj_drop.p, self.training)
            # K10: fc1 + bias + GELU + dropout + fc2's matmul on the tcgen05 GEMMs with fused epilogues
            v0 = ops.feed_forward_core(h, ff.fc1.weight, ff.fc1.bias, ff.fc2.weight, ff.drop.p, self.training)
        # fc2's bias, dropout, residual (and the next norm1): K6
        if merge:
            return ops.bias_dropout_add_merge_ln(x, v0, ff.fc2.bias, next_norm, ff.drop.p, self.training, io=io)
        return ops.bias_dropout_add_ln(x, v0, ff.fc2.bias, next_norm, ff.drop.p, self.training, io=io)


class PartAttentionLayer(nn.Module):
    """`depth` blocks, odd ones shifted by TP//2 frames, then the merge (HWGATE.py:223-258)."""

    def __init__(self, dim, temporal_patch_size, temporal_dim, num_kps, depth, num_heads, window_size, adj_mat,
                 drop=0., attn_drop=0., ff_ratio=4., norm_layer=nn.LayerNorm, downsample=None, i_layer=0,
                 device=None):
        super().__init__()
        self.dim = dim
        self.depth = depth
        self.num_heads = num_heads
        self.window_size = window_size
        self.adj_mat = adj_mat.to(device) if adj_mat is not None else None
        self.i_layer = i_layer
        self.blocks = nn.ModuleList([
            PartAttentionBlock(dim=dim, num_kps=num_kps, num_heads=num_heads, window_size=window_size,
                               temporal_patch_size=temporal_patch_size, temporal_dim=temporal_dim,
                               shift_size=0 if (i % 2 == 0) else temporal_patch_size // 2,
                               adj_mat=self.adj_mat, drop=drop, attn_drop=attn_drop, ff_ratio=ff_ratio,
                               norm_layer=norm_layer)
            for i in range(depth)])
        self.downsample = downsample(dim, temporal_patch_size) if downsample is not None else None

    def fusable(self, x):
        blocks = list(self.blocks)
        return bool(blocks) and all(b._fusable(x) for b in blocks)

    def forward_fused(self, x, xn=None, next_level_norm=None):
        """The level on the fused bf16 kernels.  xn: norm1 of the first block already applied to x (bf16), or None.
        next_level_norm: the first norm1 of the NEXT level; when given (and this level ends in a TemporalMerging the
        fold supports) the last K6 stores the merged layout and computes that LayerNorm, and (x_merged, xn_next) is
        returned; otherwise (x_merged or x, None)."""
        blocks = list(self.blocks)
        if xn is None:
            first = blocks[0]
            x, xn = ops.layer_norm_residual(x, first.norm1.weight, first.norm1.bias, first.norm1.eps,
                                            io=first._chain_io(x))
        fold = (next_level_norm is not None and type(self.downsample) is TemporalMerging
                and self.downsample.temporal_patch_size == ops.TEMPORAL_PATCH and type(next_level_norm) is nn.LayerNorm
                and ops.merge_fold_supported(self.dim, x.shape[1]))
        for i, blk in enumerate(blocks):
            last = i + 1 == len(blocks)
            if last and fold:
                return blk.forward_chain(x, xn, next_level_norm, merge=True)
            x, xn = blk.forward_chain(x, xn, None if last else blocks[i + 1].norm1)
        if self.downsample is not None:
            x = self.downsample(x)
        return x, None

    def forward(self, x):
        if self.fusable(x):
            return self.forward_fused(x)[0]
        for blk in self.blocks:
            x = blk(x)
        if self.downsample is not None:
            x = self.downsample(x)
        return x


class Model(nn.Module):
    """HWGATE classifier: (B,T,K,C) keypoints -> (B,num_classes) (HWGATE.py:260-360)."""

    def __init__(self, kp_dim=26, num_kps=64, temporal_dim=256, num_classes=1000, embed_dim=64,
                 temporal_patch_size=4, pe=False, depths=[2, 2, 6, 2], num_heads=[2, 4, 8, 16], window_size=16,
                 adj_mat=None, drop_rate=0., attn_drop_rate=0., ff_ratio=4., norm_layer=nn.LayerNorm,
                 device=None) -> None:
        super().__init__()
        self.kp_dim = kp_dim
        self.num_kps = num_kps
        self.temporal_dim = temporal_dim
        self.window_size = window_size
        self.num_classes = num_classes
        self.num_layers = len(depths)
        self.pe = pe
        self.adj_mat = adj_mat
        self.embed_dim = embed_dim
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.temporal_out_dim = temporal_dim // temporal_patch_size ** (self.num_layers - 1)

        assert self.num_kps % window_size == 0, "window size and number of kps are incompatible"
        assert self.temporal_dim % temporal_patch_size == 0, \
            "temporal dimension and temporal patch size are incompatible"

        # frozen Gaussian Fourier features, std 10 (HWGATE.py:293-299)
        self.B = nn.Parameter(torch.randn(embed_dim // 2, self.kp_dim) * 10.0, requires_grad=False)
        if self.pe:
            self.pos_encoder = PositionalEncoding(embed_dim, drop_rate, temporal_dim)

        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            frames = temporal_dim // temporal_patch_size ** i
            # the reference replicates the adjacency over the temporal groups of the level
            # (HWGATE.py:309); kept so that layer.adj_mat has the reference's shape
            adj_t = torch.cat([adj_mat] * (frames // temporal_patch_size)) if adj_mat is not None else None
            self.layers.append(PartAttentionLayer(
                dim=int(embed_dim * 2 ** i), temporal_patch_size=temporal_patch_size, temporal_dim=frames,
                num_kps=num_kps, depth=depths[i], num_heads=num_heads[i], window_size=window_size, adj_mat=adj_t,
                drop=drop_rate, attn_drop=attn_drop_rate, ff_ratio=ff_ratio, norm_layer=norm_layer,
                downsample=TemporalMerging if i < self.num_layers - 1 else None, i_layer=i, device=device))

        self.norm = norm_layer(self.num_features)
        self.avgpool = nn.AvgPool1d(self.temporal_out_dim * self.num_kps)
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # (overridden by the sibling model HGATE, which stores its 29 keypoints as 32)
    _kp_real = 0

    def _embed_fused(self, x):
        """K8: Fourier embedding + positional encoding + its dropout in one pass, fp32 (bf16 would round the
        argument 2*pi*x.B, values up to ~300, by up to 2 rad: SURVEY.md 7.2)."""
        return ops.fourier_embed(x, self.B, self.pos_encoder.pe, self.pos_encoder.dropout.p, self.training)

    def forward_features(self, x):
        # the fused chain: bf16 autocast, or fp32 without autocast in the x3 mode (PartAttentionBlock._chain_io)
        fused = chain_enabled(x)
        if fused and self.pe and not self.B.requires_grad and not x.requires_grad:
            x = self._embed_fused(x)
        else:
            x = embed_generic(self, x)
        layers = list(self.layers)
        if fused and all(isinstance(l, PartAttentionLayer) and l.fusable(x) for l in layers[:1]):
            # levels chained on the fused kernels: each level's last residual add stores TemporalMerging's layout and
            # applies the next level's first LayerNorm in the same pass (K4 folded away)
            xn = None
            for i, layer in enumerate(layers):
                if not layer.fusable(x):
                    x, xn = layer(x), None
                    continue
                nxt = layers[i + 1].blocks[0].norm1 if i + 1 < len(layers) and len(layers[i + 1].blocks) else None
                x, xn = layer.forward_fused(x, xn, nxt)
        else:
            for layer in layers:
                x = layer(x)
        B, f, K, d = x.shape
        if fused and type(self.norm) is nn.LayerNorm and d in (128, 256, 512):
            # K9: final LayerNorm + mean over the f*K tokens (self.avgpool, HWGATE.py:354) in one pass
            return ops.layer_norm_mean_pool(x, self.norm.weight, self.norm.bias, self.norm.eps, kp_real=self._kp_real)
        x = self.norm(x)
        # self.avgpool (AvgPool1d over all f*K tokens, HWGATE.py:354) is a mean over tokens; mean() has the
        # same value and a broadcast backward instead of avg_pool2d_backward (3.5 ms per step at B=512)
        return x.reshape(B, f * K, d).mean(dim=1)

    def forward(self, x):
        if AUTOCAST == "bf16" and x.is_cuda and not torch.is_autocast_enabled():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.forward(x)
        feats = self.forward_features(x)
        if feats.is_cuda and type(self.head) is nn.Linear:
            # K13: the classifier head in fp32 on the library's FFMA GEMM (also under autocast: its output feeds
            # the log-softmax of the loss)
            return ops.linear_f32(feats, self.head.weight, self.head.bias)
        return self.head(feats)
