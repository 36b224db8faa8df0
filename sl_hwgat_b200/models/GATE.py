"""B200-native GATE: drop-in for hwgat/models/GATE.py (the third sibling model).

GATE is the plain ablation: `depths` identical blocks, each attending over ALL F * 29 tokens of a sample (N = 1856 at
T = 64, GATE.py:49-66) with the graph as an ADDITIVE mask - 0 on edges, -10000 elsewhere (GATE.py:142, 60-62) - and a
learned weighted average over the tokens instead of the mean pool (`weightedAvg`, GATE.py:185, 207).  Same class
names, constructor / forward signatures and state_dict keys as the reference.

How it runs here: GATEParams' graph links a token to keypoints of its own frame and to itself in the two adjacent
frames (model_params.py:59-74), so the dense 1856^2 softmax is exactly a softmax over a 3-frame band; the keypoint axis
is stored padded to 32 and K15 / K16 (ops.band_graph_attention, one "window" of 32 keypoints) evaluate that band on
the (B, F, 32, d) stream; the three padded keypoints have no edges (zero context, never a key), are skipped by the
weighted pool and carry no gradient.  Everything else runs on the HWGATE kernels (K5, K6, K10, K12; K8 embedding, K9
with token weights, K13 head).  ops.band_mask_pack proves the band property of whatever `adj_mask` holds and refuses
anything else: there is no dense fallback.  Precision as WGATE: bf16 autocast = the fused chain; no autocast = the
true-fp32 band kernels (1e-5) between PyTorch LayerNorm / Linear / GELU.
"""
import torch
import torch.nn as nn

from sl_hwgat_b200 import _lib, ops
from sl_hwgat_b200.models import HWGATE as _hw
from sl_hwgat_b200.models.HGATE import KP_PAD, _pad_kp
from sl_hwgat_b200.models.HWGATE import FeedForward, PositionalEncoding  # noqa: F401 (reference names)
from sl_hwgat_b200.models.WGATE import _BandBits, _NEED_BF16


class MSA(nn.Module):
    """Multi-head graph attention over all tokens of a sample with an additive mask (GATE.py:30-69)."""

    def __init__(self, num_heads, dim, adj_mask=None, attn_drop=0., proj_drop=0.) -> None:
        super().__init__()
        assert dim % num_heads == 0, 'dim and number of heads are incompatible'
        self.dim = dim
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.adj_mask = adj_mask
        self.softmax = nn.Softmax(dim=-1)

    def context(self, xn, bits, window=KP_PAD):
        if self.training and self.attn_drop.p > 0:
            raise _lib.HwgatError("band attention has no attention dropout (the reference's default is 0; no fallback)")
        return ops.band_graph_attention(xn.to(_hw._attn_dtype(xn)), self.qkv.weight, self.qkv.bias, bits,
                                        self.num_heads, window)

    # -- reference signature: x is (B, F*K, d), normalised
    def forward(self, x, parent):
        B, F_K, d = x.shape
        if self.adj_mask is None:
            raise _lib.HwgatError("GATE without a graph is dense attention over all tokens: not built (no fallback)")
        K, F = parent.num_kps, F_K // parent.num_kps
        bits = parent._bits.get(getattr(parent, self.adj_mask), F, KP_PAD, x.device)
        xb = _pad_kp(x.reshape(B, F, K, d), 2)
        ctx = self.context(xb, bits)[:, :, :K].reshape(B, F_K, d)
        return self.proj_drop(_hw.linear(self.proj, ctx))


class AttentionBlock(nn.Module):
    """x + MSA(LN x) over all tokens, then x + FFN(LN x) (GATE.py:88-116).  The fused chain works on the padded
    (B, F, 32, d) stream; the reference's (B, F*K, d) input is accepted, padded and cut back."""

    _chain_io = _hw.PartAttentionBlock._chain_io          # bf16 under autocast, float32 in the x3 fp32 mode
    _fusable = _hw.PartAttentionBlock._fusable
    forward_chain = _hw.PartAttentionBlock.forward_chain

    def __init__(self, dim, num_kps=27, num_heads=8, ff_ratio=4., temporal_dim=128, adj_mask=None, drop=0.,
                 attn_drop=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if num_kps > KP_PAD:
            raise NotImplementedError("the band-attention kernels hold one frame's keypoints as a 32-slot window")
        self.dim = dim
        self.num_kps = num_kps
        self.temporal_dim = temporal_dim
        self.drop = drop
        self.ff_dim = self.dim * ff_ratio
        self.attn_drop = attn_drop
        self.norm1 = norm_layer(dim)
        self.attn = MSA(num_heads, dim, adj_mask=adj_mask, attn_drop=attn_drop, proj_drop=drop)
        self.norm2 = norm_layer(dim)
        self.ff = FeedForward(in_features=dim, hidden_features=int(self.ff_dim), act_layer=act_layer, drop=drop)

    def _context(self, x, xn, bits=None):
        return self.attn.context(xn, bits)

    def band_bits(self, x, parent):
        if self.attn.adj_mask is None:
            raise _lib.HwgatError("GATE without a graph is dense attention over all tokens: not built (no fallback)")
        return parent._bits.get(getattr(parent, self.attn.adj_mask), x.shape[1], KP_PAD, x.device)

    def supported(self, x):
        B, F, K, d = x.shape
        return (K == KP_PAD and self._fusable(x)
                and ops.band_attention_supported(B, F, K, d, self.attn.num_heads, KP_PAD))

    def forward_generic(self, x, bits):
        """the block on the padded (B, F, 32, d) stream with PyTorch LayerNorm / Linear / GELU around the band
        attention: the fp32 mode (GATE.py:111-116)"""
        a = self.attn
        x = x + a.proj_drop(_hw.linear(a.proj, a.context(self.norm1(x), bits)))
        return x + self.ff(self.norm2(x))

    def forward(self, x, parent):
        B, F_K, d = x.shape
        K = parent.num_kps
        if not x.is_cuda:
            raise _lib.HwgatError(_NEED_BF16.format("GATE"))
        xp = _pad_kp(x.reshape(B, F_K // K, K, d), 2)
        bits = self.band_bits(xp, parent)
        if self.supported(xp):
            xp, xn = ops.layer_norm_residual(xp, self.norm1.weight, self.norm1.bias, self.norm1.eps, io=self._chain_io(xp))
            y = self.forward_chain(xp, xn, None, bits=bits)[0]
        else:
            y = self.forward_generic(xp, bits)
        return y[:, :, :K].reshape(B, F_K, d)


class Model(nn.Module):
    """GATE classifier: (B,T,29,C) keypoints -> (B,num_classes) (GATE.py:118-216)."""

    def __init__(self, kp_dim=26, num_kps=29, temporal_dim=256, num_classes=1000, embed_dim=64, pe=False, depths=16,
                 num_heads=8, ff_ratio=4., adj_mat=None, drop_rate=0., attn_drop_rate=0., norm_layer=nn.LayerNorm,
                 device=None) -> None:
        super().__init__()
        self.kp_dim = kp_dim
        self.num_kps = num_kps
        self.temporal_dim = temporal_dim
        self.num_classes = num_classes
        self.pe = pe
        self.num_heads = num_heads
        self.ff_ratio = ff_ratio
        self.embed_dim = embed_dim
        self.drop_rate = drop_rate
        self.attn_drop_rate = attn_drop_rate
        self.norm_layer = norm_layer
        # the additive mask, a persistent (1, 1, N, N) buffer with the reference's name and values (GATE.py:142-152)
        self.adj_mask_name = 'adj_mask'
        adj_mask = adj_mat.masked_fill(adj_mat == 0, float(-10000)).masked_fill(adj_mat == 1, float(0))
        self.register_buffer(self.adj_mask_name, adj_mask.unsqueeze(0).unsqueeze(0).to(device))
        self._bits = _BandBits()

        self.B = nn.Parameter(torch.randn(embed_dim // 2, self.kp_dim) * 10.0, requires_grad=False)
        if self.pe:
            self.pos_encoder = PositionalEncoding(self.embed_dim, self.drop_rate, self.temporal_dim)
        self.layers = nn.ModuleList([
            AttentionBlock(dim=self.embed_dim, num_kps=self.num_kps, num_heads=self.num_heads, ff_ratio=self.ff_ratio,
                           adj_mask=self.adj_mask_name, drop=self.drop_rate, attn_drop=self.attn_drop_rate,
                           norm_layer=self.norm_layer)
            for _ in range(depths)])
        self.norm = norm_layer(self.embed_dim)
        self.weightedAvg = nn.Linear(self.temporal_dim * self.num_kps, 1)
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

    _init_weights = _hw.Model._init_weights

    def forward_features(self, x):
        if not (x.is_cuda and x.shape[2] == self.num_kps <= KP_PAD):
            raise _lib.HwgatError(_NEED_BF16.format("GATE"))
        blocks = list(self.layers)
        fast = (_hw.chain_enabled(x) and self.pe and not x.requires_grad
                and type(self.norm) is nn.LayerNorm and self.embed_dim in (128, 256, 512))
        # the keypoint axis is padded to 32: the padded keypoints embed to finite values, have no edges, are never
        # pooled and carry no gradient
        if fast:
            x = ops.fourier_embed(_pad_kp(x, 2), self.B, self.pos_encoder.pe, self.pos_encoder.dropout.p, self.training)
        else:
            x = _hw.embed_generic(self, _pad_kp(x, 2))
        bits = blocks[0].band_bits(x, self) if blocks else None
        if fast and blocks and blocks[0].supported(x):
            first = blocks[0].norm1
            x, xn = ops.layer_norm_residual(x, first.weight, first.bias, first.eps, io=blocks[0]._chain_io(x))    # K5
            for i, blk in enumerate(blocks):
                nxt = blocks[i + 1].norm1 if i + 1 < len(blocks) else None
                x, xn = blk.forward_chain(x, xn, nxt, bits=bits)
        else:
            for blk in blocks:
                x = blk.forward_generic(x, bits)
        if fast:
            # K9 with token weights: final LayerNorm + weightedAvg over the F*29 real tokens (GATE.py:205-207)
            return ops.layer_norm_weighted_pool(x, self.norm.weight, self.norm.bias, self.weightedAvg.weight,
                                                self.weightedAvg.bias, self.norm.eps, kp_real=self.num_kps)
        B, F, _, d = x.shape
        K = self.num_kps
        x = self.norm(x[:, :, :K]).reshape(B, F * K, d)
        return self.weightedAvg(x.transpose(1, 2)).squeeze(-1)

    def forward(self, x):
        if _hw.AUTOCAST == "bf16" and x.is_cuda and not torch.is_autocast_enabled():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.forward(x)
        feats = self.forward_features(x)
        if type(self.head) is nn.Linear:
            return ops.linear_f32(feats, self.head.weight, self.head.bias)                                   # K13
        return self.head(feats)
