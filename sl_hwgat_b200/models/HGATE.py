"""B200-native HGATE: drop-in for hwgat/models/HGATE.py (the first sibling model).

HGATE is HWGATE without keypoint windows: a block is ALL 29 keypoints of TP = 2 consecutive frames (58 tokens,
HGATE.py:30-36), masked by one (58, 58) skeleton adjacency (model_params.py:459-481) times a shifted-block mask
(HGATE.py:155-171), both multiplicative with the -10000 fill (HGATE.py:93-102); there is no training-time threshold
drop.  Same class names, constructor / forward signatures and state_dict keys as the reference.

How it runs here: the keypoint axis is stored padded to 32, so a block is one window of N = 2 x 32 = 64 tokens of the
general-window tcgen05 attention (ops.window_graph_attention, window = 32: K2b / K3b) with the six padded tokens of
every block masked out as keys; every other kernel of the HWGATE block (K5, K6, K10, K12, the folded temporal merge)
runs unchanged on the padded rows, and the final LayerNorm + mean pool (K9) skips them.  Padded rows never reach the
loss, so they get zero gradient and contribute nothing to any parameter gradient.

Precision: inside `torch.autocast("cuda", dtype=torch.bfloat16)` the fused bf16 chain runs (the timed path).  Without
autocast (the reference's unmodified loop, utils.py:102) the attention runs the true-fp32 general-window kernels
(attn_win_f32.cu, 1e-5 against the reference's fp64 outputs) between PyTorch LayerNorm / Linear / GELU, the keypoint
axis padded around every layer: correct, not fast.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F_

from sl_hwgat_b200 import _lib, ops
from sl_hwgat_b200.models import HWGATE as _hw
from sl_hwgat_b200.models.HWGATE import FeedForward, PositionalEncoding, TemporalMerging  # noqa: F401 (reference names)
from sl_hwgat_b200.ops import LAYOUT_WINDOWS

KP_PAD = 32          # stored keypoints per frame (29 real ones; 3 padded)


def block_partition(x, temporal_patch_size=4):
    """(B,F,K,d) -> (B*f, TP*K, d) (HGATE.py:30-36)."""
    B, F, K, d = x.shape
    return x.reshape(B * (F // temporal_patch_size), temporal_patch_size * K, d)


def block_reverse(x, temporal_patch_size=4, temporal_dim=128, num_kp=64):
    """Inverse of block_partition (HGATE.py:39-47)."""
    f = temporal_dim // temporal_patch_size
    return x.reshape(x.shape[0] // f, temporal_dim, num_kp, x.shape[-1])


def _pad_kp(x, dim):
    """zero-pad the keypoint axis `dim` to KP_PAD"""
    k = x.shape[dim]
    if k == KP_PAD:
        return x
    pad = [0, 0] * (x.dim() - 1 - dim) + [0, KP_PAD - k]
    return F_.pad(x, pad)


def _pad_mask(m, K, TP):
    """(..., TP*K, TP*K) float mask over block tokens tp*K + k -> (..., TP*32, TP*32) over tp*32 + k, zeros on pads"""
    lead = m.shape[:-2]
    m = m.reshape(*lead, TP, K, TP, K)
    out = m.new_zeros(*lead, TP, KP_PAD, TP, KP_PAD)
    out[..., :, :K, :, :K] = m
    return out.reshape(*lead, TP * KP_PAD, TP * KP_PAD)


class MSA(_hw.MSA):
    """Block multi-head graph attention (HGATE.py:65-109): no threshold drop, adj_mat is ONE (TP*K, TP*K) matrix."""

    def _draw_threshold(self):
        return None                      # HGATE.py:91-102 has no training-time drop

    # -- reference signature: x is (B*f, TP*K, d), already rolled and partitioned (and normalised)
    def forward(self, x, B, f, attn_mask=None):
        B_f, TP_K, d = x.shape
        TP = ops.TEMPORAL_PATCH
        K = TP_K // TP
        if B_f != B * f or K * TP != TP_K or K > KP_PAD:
            raise ValueError("x must be (B*f, 2*num_kps, d) with num_kps <= 32")
        if (B * f) % 2 and _hw._attn_dtype(x) == torch.bfloat16:
            raise _lib.HwgatError("the padded block layout needs an even number of blocks (B*f) on the bf16 kernels")
        adj = self.adj_mat
        key = (None if attn_mask is None else (attn_mask.data_ptr(), attn_mask._version),
               None if adj is None else (adj.data_ptr(), adj._version), f, x.device)
        bits = self._bits.get(key)
        if bits is None:
            self._bits.clear()
            dev = x.device
            a = _pad_mask((adj if adj is not None else torch.ones(TP_K, TP_K)).to(dev).float(), K, TP)[None]
            m = _pad_mask(attn_mask.to(dev).float(), K, TP) if attn_mask is not None else None
            bits = ops.mask_pack(a, m, f, TP * KP_PAD, dev)
            self._bits[key] = bits
        xp = _pad_kp(x.reshape(B_f, TP, K, d), 2).reshape(B_f, TP * KP_PAD, d)
        ctx = ops.window_graph_attention(xp.to(_hw._attn_dtype(x)), self.qkv.weight, self.qkv.bias, bits, self.num_heads,
                                         shift=0, threshold=None, layout=LAYOUT_WINDOWS, frames=f * TP, kps=KP_PAD,
                                         window=KP_PAD, attn_drop=self._attn_p())
        ctx = ctx.reshape(B_f, TP, KP_PAD, d)[:, :, :K].reshape(B_f, TP_K, d)
        return self._project(ctx)


class GraphAttentionBlock(_hw.PartAttentionBlock):
    """x + MSA(LN x) over (shifted) blocks of TP frames, then x + FFN(LN x) (HGATE.py:124-213).  Works on the
    padded (B, F, 32, d) stream; a (B, F, 29, d) input is padded and the result cut back."""

    def __init__(self, dim, num_kps=64, num_heads=4, temporal_patch_size=4, temporal_dim=128, shift_size=0,
                 adj_mat=None, drop=0., attn_drop=0., ff_ratio=4., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        if num_kps > KP_PAD:
            raise NotImplementedError("the block-attention kernels hold one block as a 64-token window: num_kps <= 32")
        super().__init__(dim, num_kps=KP_PAD, num_heads=num_heads, window_size=KP_PAD,
                         temporal_patch_size=temporal_patch_size, temporal_dim=temporal_dim, shift_size=shift_size,
                         adj_mat=adj_mat, drop=drop, attn_drop=attn_drop, ff_ratio=ff_ratio, act_layer=act_layer,
                         norm_layer=norm_layer)
        self.real_kps = num_kps
        # the reference's MSA class / signature, sharing the parameters the parent created
        attn = MSA(dim, num_heads=num_heads, adj_mat=adj_mat, attn_drop=attn_drop, proj_drop=drop)
        attn.qkv, attn.proj, attn.window_size = self.attn.qkv, self.attn.proj, KP_PAD
        self.attn = attn
        # `attn_mask` keeps the reference's name and shape (f, TP*K, TP*K) (HGATE.py:155-173)
        if shift_size > 0:
            TP, K = temporal_patch_size, num_kps
            frame = torch.arange(temporal_dim)
            group = (frame >= temporal_dim - TP).long() + (frame >= temporal_dim - shift_size).long()
            tok = group.reshape(temporal_dim // TP, TP, 1).expand(-1, -1, K).reshape(temporal_dim // TP, TP * K)
            self.attn_mask = (tok[:, :, None] == tok[:, None, :]).float()

    def _block_bits(self, device):
        adj = self.attn.adj_mat
        key = (device, None if adj is None else (adj.data_ptr(), adj._version, tuple(adj.shape)))
        if self._bits is None or self._bits_key != key:
            TP, K = self.temporal_patch_size, self.real_kps
            a = adj.to(device).float() if adj is not None else torch.ones(TP * K, TP * K, device=device)
            # one keypoint "window" of 32 slots; K1b adds the shift-group condition by frame index
            self._bits = ops.mask_build(_pad_mask(a, K, TP)[None].contiguous(), self.temporal_dim, self.shift_size,
                                        KP_PAD, TP)
            self._bits_key = key
        return self._bits

    def _check_shape(self, x):
        if x.shape[1] != self.temporal_dim or x.shape[2] not in (self.real_kps, KP_PAD):
            raise ValueError(f"expected (B,{self.temporal_dim},{self.real_kps} or {KP_PAD},d), got {tuple(x.shape)}")

    def forward(self, x):
        self._check_shape(x)
        if not x.is_cuda:
            raise _lib.HwgatError("HGATE runs on the sm_100a kernels only (no CPU fallback)")
        K = x.shape[2]
        y = super().forward(_pad_kp(x, 2))       # fused bf16 chain under autocast, the fp32 parity kernels without
        return y[:, :, :K] if K != KP_PAD else y


class BlockAttentionLayer(_hw.PartAttentionLayer):
    """`depth` blocks, odd ones shifted by TP//2 frames, then the merge (HGATE.py:215-255)."""

    def __init__(self, dim, temporal_patch_size, temporal_dim, num_kps, depth, num_heads, adj_mat, drop=0.,
                 attn_drop=0., ff_ratio=4., norm_layer=nn.LayerNorm, downsample=None, i_layer=0, device=None):
        nn.Module.__init__(self)
        self.dim = dim
        self.depth = depth
        self.num_heads = num_heads
        self.window_size = KP_PAD
        self.adj_mat = adj_mat.to(device) if adj_mat is not None else None
        self.i_layer = i_layer
        self.blocks = nn.ModuleList([
            GraphAttentionBlock(dim=dim, num_kps=num_kps, num_heads=num_heads, temporal_patch_size=temporal_patch_size,
                                temporal_dim=temporal_dim, shift_size=0 if (i % 2 == 0) else temporal_patch_size // 2,
                                adj_mat=self.adj_mat, drop=drop, attn_drop=attn_drop, ff_ratio=ff_ratio,
                                norm_layer=norm_layer)
            for i in range(depth)])
        self.downsample = downsample(dim, temporal_patch_size) if downsample is not None else None

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.HwgatError("HGATE runs on the sm_100a kernels only (no CPU fallback)")
        K = x.shape[2]
        y = super().forward(_pad_kp(x, 2))       # fused when the shapes / autocast allow it, block by block otherwise
        return y[:, :, :K] if K != KP_PAD else y


class Model(_hw.Model):
    """HGATE classifier: (B,T,29,C) keypoints -> (B,num_classes) (HGATE.py:257-346)."""

    def __init__(self, kp_dim=26, num_kps=64, temporal_dim=256, num_classes=1000, embed_dim=64,
                 temporal_patch_size=4, pe=False, depths=[2, 2, 6, 2], num_heads=[2, 4, 8, 16], adj_mat=None,
                 drop_rate=0., attn_drop_rate=0., ff_ratio=4., norm_layer=nn.LayerNorm, device=None) -> None:
        nn.Module.__init__(self)
        self.kp_dim = kp_dim
        self.num_kps = num_kps
        self.temporal_dim = temporal_dim
        self.num_classes = num_classes
        self.num_layers = len(depths)
        self.pe = pe
        self.adj_mat = adj_mat
        self.embed_dim = embed_dim
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.temporal_out_dim = temporal_dim // temporal_patch_size ** (self.num_layers - 1)
        assert self.temporal_dim % temporal_patch_size == 0, "temporal dimension and temporal patch size are incompatible"
        self._kp_real = num_kps
        self.B = nn.Parameter(torch.randn(embed_dim // 2, self.kp_dim) * 10.0, requires_grad=False)
        if self.pe:
            self.pos_encoder = PositionalEncoding(embed_dim, drop_rate, temporal_dim)
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BlockAttentionLayer(
                dim=int(embed_dim * 2 ** i), temporal_patch_size=temporal_patch_size,
                temporal_dim=temporal_dim // temporal_patch_size ** i, num_kps=num_kps, depth=depths[i],
                num_heads=num_heads[i], adj_mat=adj_mat, drop=drop_rate, attn_drop=attn_drop_rate, ff_ratio=ff_ratio,
                norm_layer=norm_layer, downsample=TemporalMerging if i < self.num_layers - 1 else None, i_layer=i,
                device=device))
        self.norm = norm_layer(self.num_features)
        self.avgpool = nn.AvgPool1d(self.temporal_out_dim * self.num_kps)
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

    def _embed_fused(self, x):
        # the three padded keypoints embed to finite values (sin 0, cos 0 + pe); they are masked as keys, skipped by
        # the pool, and carry no gradient
        return super()._embed_fused(_pad_kp(x, 2))

    def forward_features(self, x):
        if not x.is_cuda:
            raise _lib.HwgatError("HGATE runs on the sm_100a kernels only (no CPU fallback)")
        return super().forward_features(x)
