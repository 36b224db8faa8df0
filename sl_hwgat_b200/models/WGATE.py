"""B200-native WGATE: drop-in for hwgat/models/WGATE.py (the second sibling model).

WGATE is the non-hierarchical ablation: `depths` identical blocks at one width, each attending inside windows of
`window_size` keypoints over ALL frames (N = F * 16 tokens per window, WGATE.py:32-45), with the graph as an ADDITIVE
mask - 0 on edges, -10000 elsewhere (WGATE.py:190, 102-106) - no shift, no threshold drop, no merging.  Same class names,
constructor / forward signatures and state_dict keys as the reference (`adj_mask` stays a buffer of the reference's
shape and values).

How it runs here: the graph WGATEParams builds links a token to keypoints of its own frame and to itself in the two
adjacent frames (model_params.py:204-229), so the dense (F*16)^2 softmax the reference evaluates is exactly a softmax
over a 3-frame band; K15 / K16 (ops.band_graph_attention) evaluate only that band on the (B, F, K, d) stream - the
partition / reverse copies disappear into index arithmetic - and every other stage of the block runs on the HWGATE
kernels (K5, K6, K10, K12; K8 embedding, K9 pool, K13 head).  ops.band_mask_pack proves the band property of whatever
`adj_mask` holds and refuses a mask that is not banded: there is no dense fallback.

Precision: inside `torch.autocast("cuda", dtype=torch.bfloat16)` the fused bf16 chain runs (the timed path).  Without
autocast (the reference's unmodified loop, utils.py:102) the attention runs the true-fp32 band kernels (1e-5 against
the reference's fp64 outputs) between PyTorch LayerNorm / Linear / GELU: correct, not fast.
"""
import torch
import torch.nn as nn

from sl_hwgat_b200 import _lib, ops
from sl_hwgat_b200.models import HWGATE as _hw
from sl_hwgat_b200.models.HWGATE import FeedForward, PositionalEncoding  # noqa: F401 (reference names)

_NEED_BF16 = ("{} runs on the sm_100a kernels only: call it on an fp32 CUDA tensor, under "
              "torch.autocast('cuda', dtype=torch.bfloat16) for the fast path (no CPU fallback)")


def window_partition(x, window_size=16):
    """(B,F,K,d) -> (B*nW, F*W, d) (WGATE.py:32-45)."""
    B, F, K, d = x.shape
    nW = K // window_size
    return x.reshape(B, F, nW, window_size, d).transpose(1, 2).reshape(B * nW, F * window_size, d)


def window_reverse(x, window_size=16, temporal_dim=128, num_kp=64):
    """Inverse of window_partition (WGATE.py:49-66)."""
    nW = num_kp // window_size
    B = x.shape[0] // nW
    return x.reshape(B, nW, temporal_dim, window_size, x.shape[-1]).transpose(1, 2).reshape(B, temporal_dim, num_kp, -1)


class _BandBits:
    """Packed band words of the parent model's `adj_mask` buffer, rebuilt when the buffer is replaced, modified in
    place or moved (the reference adds whatever the buffer holds on every call, WGATE.py:102-104)."""

    def __init__(self):
        self._key, self._bits = None, None

    def get(self, mask, frames, window, device):
        key = (mask.data_ptr(), mask._version, tuple(mask.shape), frames, window, device)
        if key != self._key:
            self._bits = ops.band_mask_pack(mask.to(device), frames, window)
            self._key = key
        return self._bits


class MSA(nn.Module):
    """Window multi-head graph attention with an additive mask (WGATE.py:68-108)."""

    def __init__(self, num_heads, dim, adj_mask=None, attn_drop=0., proj_drop=0.) -> None:
        super().__init__()
        assert dim % num_heads == 0, 'dim and number of heads are incompatible'
        self.dim = dim
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.adj_mask = adj_mask
        self.qkv = nn.Linear(dim, dim * 3)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)

    def _check_drop(self):
        if self.training and self.attn_drop.p > 0:
            raise _lib.HwgatError("band attention has no attention dropout (the reference's default is 0; no fallback)")

    def context(self, xn, bits, window):
        self._check_drop()
        return ops.band_graph_attention(xn.to(_hw._attn_dtype(xn)), self.qkv.weight, self.qkv.bias, bits,
                                        self.num_heads, window)

    # -- reference signature: x is (B*nW, F*W, d), already partitioned and normalised
    def forward(self, x, B, nW, parent):
        B_nW, F_W, d = x.shape
        if self.adj_mask is None:
            raise _lib.HwgatError("WGATE without a graph is dense attention over all frames: not built (no fallback)")
        mask = getattr(parent, self.adj_mask)
        W = parent.window_size
        F = F_W // W
        bits = parent._bits.get(mask, F, W, x.device)
        xb = window_reverse(x, W, F, nW * W)
        ctx = window_partition(self.context(xb, bits, W), W)
        return self.proj_drop(_hw.linear(self.proj, ctx))


class PartAttentionBlock(nn.Module):
    """x + MSA(LN x) inside keypoint windows over all frames, then x + FFN(LN x) (WGATE.py:126-160)."""

    _chain_io = _hw.PartAttentionBlock._chain_io          # bf16 under autocast, float32 in the x3 fp32 mode
    _fusable = _hw.PartAttentionBlock._fusable
    forward_chain = _hw.PartAttentionBlock.forward_chain

    def __init__(self, dim, num_kps=64, num_heads=4, window_size=16, ff_ratio=4., temporal_dim=128, adj_mask=None,
                 drop=0., attn_drop=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.num_kps = num_kps
        self.num_heads = num_heads
        self.window_size = window_size
        self.temporal_dim = temporal_dim
        self.drop = drop
        self.ff_dim = self.dim * ff_ratio
        self.attn_drop = attn_drop
        self.norm1 = norm_layer(dim)
        self.attn = MSA(num_heads, dim, adj_mask=adj_mask, attn_drop=attn_drop, proj_drop=drop)
        self.norm2 = norm_layer(dim)
        self.ff = FeedForward(in_features=dim, hidden_features=int(self.ff_dim), act_layer=act_layer, drop=drop)

    def _context(self, x, xn, bits=None):
        return self.attn.context(xn, bits, self.window_size)

    def band_bits(self, x, parent):
        if self.attn.adj_mask is None:
            raise _lib.HwgatError("WGATE without a graph is dense attention over all frames: not built (no fallback)")
        return parent._bits.get(getattr(parent, self.attn.adj_mask), x.shape[1], self.window_size, x.device)

    def supported(self, x):
        B, F, K, d = x.shape
        return self._fusable(x) and ops.band_attention_supported(B, F, K, d, self.num_heads, self.window_size)

    def forward_generic(self, x, bits):
        """the block with PyTorch LayerNorm / Linear / GELU around the band attention: the fp32 mode (WGATE.py:150-160)"""
        a = self.attn
        x = x + a.proj_drop(_hw.linear(a.proj, a.context(self.norm1(x), bits, self.window_size)))
        return x + self.ff(self.norm2(x))

    def forward(self, x, parent):
        if not x.is_cuda:
            raise _lib.HwgatError(_NEED_BF16.format("WGATE"))
        bits = self.band_bits(x, parent)
        if not self.supported(x):
            return self.forward_generic(x, bits)
        x, xn = ops.layer_norm_residual(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, io=self._chain_io(x))
        return self.forward_chain(x, xn, None, bits=bits)[0]


class Model(nn.Module):
    """WGATE classifier: (B,T,64,C) keypoints -> (B,num_classes) (WGATE.py:162-263)."""

    def __init__(self, kp_dim=26, num_kps=64, temporal_dim=256, num_classes=1000, embed_dim=64, pe=False, depths=16,
                 num_heads=8, window_size=16, ff_ratio=4., adj_mat=None, drop_rate=0., attn_drop_rate=0.,
                 norm_layer=nn.LayerNorm, device=None) -> None:
        super().__init__()
        self.kp_dim = kp_dim
        self.num_kps = num_kps
        self.temporal_dim = temporal_dim
        self.window_size = window_size
        self.num_classes = num_classes
        self.pe = pe
        self.num_heads = num_heads
        self.ff_ratio = ff_ratio
        self.embed_dim = embed_dim
        self.drop_rate = drop_rate
        self.attn_drop_rate = attn_drop_rate
        self.norm_layer = norm_layer
        assert self.num_kps % window_size == 0, "window size and number of kps are incompatible"
        # the additive mask, a persistent buffer with the reference's name, shape and values (WGATE.py:190-197)
        self.adj_mask_name = 'adj_mask'
        adj_mask = adj_mat.masked_fill(adj_mat == 0, float(-10000)).masked_fill(adj_mat == 1, float(0))
        self.register_buffer(self.adj_mask_name, adj_mask.to(device))
        self._bits = _BandBits()

        self.B = nn.Parameter(torch.randn(embed_dim // 2, self.kp_dim) * 10.0, requires_grad=False)
        if self.pe:
            self.pos_encoder = PositionalEncoding(self.embed_dim, self.drop_rate, self.temporal_dim)
        self.layers = nn.ModuleList([
            PartAttentionBlock(dim=self.embed_dim, num_kps=self.num_kps, num_heads=self.num_heads,
                               window_size=self.window_size, ff_ratio=self.ff_ratio, adj_mask=self.adj_mask_name,
                               drop=self.drop_rate, attn_drop=self.attn_drop_rate, norm_layer=self.norm_layer)
            for _ in range(depths)])
        self.norm = norm_layer(self.embed_dim)
        self.avgpool = nn.AvgPool1d(self.temporal_dim * self.num_kps)
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

    _init_weights = _hw.Model._init_weights

    def forward_features(self, x):
        if not x.is_cuda:
            raise _lib.HwgatError(_NEED_BF16.format("WGATE"))
        blocks = list(self.layers)
        fast = (_hw.chain_enabled(x) and self.pe and not x.requires_grad
                and type(self.norm) is nn.LayerNorm and self.embed_dim in (128, 256, 512))
        if fast:
            x = ops.fourier_embed(x, self.B, self.pos_encoder.pe, self.pos_encoder.dropout.p, self.training)   # K8
        else:
            x = _hw.embed_generic(self, x)
        bits = blocks[0].band_bits(x, self) if blocks else None
        if fast and blocks and blocks[0].supported(x):
            first = blocks[0].norm1
            x, xn = ops.layer_norm_residual(x, first.weight, first.bias, first.eps, io=blocks[0]._chain_io(x))    # K5
            for i, blk in enumerate(blocks):
                nxt = blocks[i + 1].norm1 if i + 1 < len(blocks) else None
                x, xn = blk.forward_chain(x, xn, nxt, bits=bits)
        else:
            # shapes the chain does not take (or the FFMA fp32 mode): the band kernels between PyTorch LayerNorm / Linear / GELU
            for blk in blocks:
                x = blk.forward_generic(x, bits)
        if fast:
            # K9: final LayerNorm + mean over all F*K tokens (self.avgpool, WGATE.py:256)
            return ops.layer_norm_mean_pool(x, self.norm.weight, self.norm.bias, self.norm.eps)
        B, F, K, d = x.shape
        return self.norm(x).reshape(B, F * K, d).mean(dim=1)

    def forward(self, x):
        if _hw.AUTOCAST == "bf16" and x.is_cuda and not torch.is_autocast_enabled():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.forward(x)
        feats = self.forward_features(x)
        if type(self.head) is nn.Linear:
            return ops.linear_f32(feats, self.head.weight, self.head.bias)                                   # K13
        return self.head(feats)
