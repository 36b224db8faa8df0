"""K2b / K3b (attn_core_tc2.cu): windowed graph attention with every product on tcgen05, for windows of
W = 16, 32, 64 keypoints x 2 frames (N = 32, 64, 128 tokens).  Parity against
 (a) outputs of the unmodified reference run with window_size 32 / 64 (tests/golden/larger_windows.npz), and
 (b) the fp64 oracle on seeded inputs (the oracle itself is pinned to (a) by tests/test_oracle_golden.py),
through the C ABI.  bf16: 2e-2 relative (north_star); masks bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import hwgate_oracle as O
from tests._util import rel_inf, rel_l2

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
EDGES = O.HWGATEConfig().edges


def adj_w(W):
    return O.window_adjacency(EDGES[:64 // W], W, 2)


def dev_bits(W, F, shift):
    from sl_hwgat_b200 import ops
    return ops.mask_build(torch.from_numpy(adj_w(W).astype(np.float32)).cuda(), F, shift, W, 2)


def seeded(d, shift, W, B=1, F=4, std=None):
    std = (0.2 if d == 128 else 0.1) if std is None else std          # as tests/golden/make_golden.py section 6
    rng = np.random.default_rng(3000 + d + 10 * shift + W)
    xn = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
    return xn, w, b, g


def run_tc2(xn, w, b, g, heads, shift, thr, W, layout=0, save_qkv=True, impl="tc2"):
    from sl_hwgat_b200 import ops
    B, F = xn.shape[0], xn.shape[1]
    bits = dev_bits(W, F, shift)
    x_ = xn.to("cuda", torch.bfloat16).requires_grad_(True)
    w_ = w.to(torch.bfloat16).float().cuda().requires_grad_(True)
    b_ = b.float().cuda().requires_grad_(True)
    y = ops.window_graph_attention(x_, w_, b_, bits, heads, shift=shift, threshold=thr, layout=layout, window=W,
                                   impl=impl, save_qkv=save_qkv)
    y.backward(g.to("cuda", torch.bfloat16))
    torch.cuda.synchronize()
    return y.detach(), x_.grad, w_.grad, b_.grad


def oracle_points(xn, w, b, g, heads, F, shift, thr, W, bf16_points=True):
    mask = O.combined_mask(adj_w(W), F, W, 2, shift)
    x_, w_, b_ = (t.clone().requires_grad_(True) for t in (xn, w, b))
    y = O.attention_core(x_, w_, b_, heads, mask, W, 2, shift, thr, bf16_points=bf16_points)
    (y * g).sum().backward()
    return y.detach(), x_.grad, w_.grad, b_.grad


def rounded(xn, w, b, g):
    return xn.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), b.float().double(), g.to(torch.bfloat16).double()


# ------------------------------------------------------------------ K1 for larger windows
@pytest.mark.parametrize("W", [32, 64])
def test_larger_window_masks_bit_exact(golden_dir, W):
    from sl_hwgat_b200 import ops
    G = np.load(os.path.join(golden_dir, "larger_windows.npz"))
    adj = ops.adjacency_build(EDGES[:64 // W], W, 2, "cuda").cpu().numpy()
    assert np.array_equal(adj, G[f"adj_W{W}"].astype(np.float32))            # reference get_adj_mat() at this W
    for F in (8, 4):
        for shift in (0, 1):
            bits = dev_bits(W, F, shift).cpu().numpy().view(np.uint32)
            assert np.array_equal(bits, G[f"bits_W{W}_F{F}_s{shift}"])       # reference adj * attn_mask, packed


# ------------------------------------------------------------------ attention core, every window size
@pytest.mark.parametrize("W", [16, 32, 64])
@pytest.mark.parametrize("d,h", [(128, 2), (256, 4), (512, 8)])
@pytest.mark.parametrize("shift", [0, 1])
@pytest.mark.parametrize("thr", [None, 0.04])
def test_tc2_attention_vs_oracle(W, d, h, shift, thr):
    B, F = 2, 8
    xn, w, b, g = rounded(*seeded(d, shift, W, B=B, F=F, std=0.05))
    y, dx, dw, db = run_tc2(xn, w, b, g, h, shift, thr, W)
    by, bdx, bdw, bdb = oracle_points(xn, w, b, g, h, F, shift, thr, W)
    errs = dict(y=rel_l2(y, by), dx=rel_l2(dx, bdx), dw=rel_l2(dw, bdw), db=rel_l2(db, bdb))
    assert all(e < BF16_TOL for e in errs.values()), errs
    if thr is None:
        # eval mode is continuous: also within the tolerance of the UNROUNDED fp64 oracle
        uy, udx, udw, udb = oracle_points(xn, w, b, g, h, F, shift, None, W, bf16_points=False)
        assert rel_l2(y, uy) < BF16_TOL and rel_l2(dx, udx) < BF16_TOL and rel_l2(dw, udw) < BF16_TOL


@pytest.mark.parametrize("W", [32, 64])
@pytest.mark.parametrize("d,h", [(128, 2), (256, 4)])
@pytest.mark.parametrize("shift", [0, 1])
def test_tc2_attention_vs_reference_golden(golden_dir, W, d, h, shift):
    """Against the unmodified reference (fp64) run with window_size W, eval mode, on the golden's seeded inputs."""
    G = np.load(os.path.join(golden_dir, "larger_windows.npz"))
    key = f"W{W}_d{d}_s{shift}_thrNone"
    xn, w, b, g = seeded(d, shift, W)
    y, dx, dw, db = run_tc2(xn, w, b, g, h, shift, None, W)

    def chk(t, name, stride):
        a = t.detach().double().cpu().reshape(-1).numpy()[::stride]
        ref = G[key + "_" + name]
        return float(np.linalg.norm(a - ref) / np.linalg.norm(ref))
    # inputs here are NOT bf16-representable (the golden's are fp64 draws): the kernel sees their bf16 rounding, so
    # the bound is the bf16 tolerance
    assert chk(y, "y", 61) < BF16_TOL and chk(dx, "dx", 61) < BF16_TOL and chk(dw, "dw", 251) < BF16_TOL
    ref_db = G[key + "_db"]
    assert np.linalg.norm(db.double().cpu().numpy() - ref_db) / np.linalg.norm(ref_db) < BF16_TOL


def test_tc2_fully_masked_rows_uniform():
    """threshold below 1/N drops every logit: softmax of N fills is uniform over the window's N keys."""
    for W in (16, 64):
        xn, w, b, g = rounded(*seeded(128, 0, W, std=0.02))
        y, dx, dw, db = run_tc2(xn, w, b, g, 2, 0, 0.001, W)
        by, bdx, bdw, bdb = oracle_points(xn, w, b, g, 2, 4, 0, 0.001, W)
        assert rel_l2(y, by) < BF16_TOL
        assert float(dx.abs().max()) == 0.0 or rel_l2(dx, bdx) < BF16_TOL    # dead rows: zero logit gradient


@pytest.mark.parametrize("W,d,h,F", [(16, 128, 2, 64), (32, 256, 4, 32), (64, 512, 8, 16), (16, 512, 8, 16)])
def test_tc2_multi_item_replication(W, d, h, F):
    """Many more (tile, head) items than CTAs, and an odd count: every sample is the same sequence, so every sample's
    output / input gradient is bit-identical to the one-sample case (oracle-checked above) and dW is B times it."""
    xn, w, b, g = rounded(*seeded(d, 1, W, B=1, F=F, std=0.05))
    ys, dxs, dws, dbs = run_tc2(xn, w, b, g, h, 1, 0.05, W)
    by, bdx, bdw, bdb = oracle_points(xn, w, b, g, h, F, 1, 0.05, W)
    assert rel_l2(ys, by) < BF16_TOL and rel_l2(dxs, bdx) < BF16_TOL and rel_l2(dws, bdw) < BF16_TOL
    for Bbig in (64, 37):
        yb, dxb, dwb, dbb = run_tc2(xn.expand(Bbig, -1, -1, -1).contiguous(), w, b,
                                    g.expand(Bbig, -1, -1, -1).contiguous(), h, 1, 0.05, W)
        assert torch.equal(yb, ys.expand(Bbig, -1, -1, -1))
        assert torch.equal(dxb, dxs.expand(Bbig, -1, -1, -1))
        assert rel_l2(dwb, dws * Bbig) < 1e-3 and rel_l2(dbb, dbs * Bbig) < 1e-3


@pytest.mark.parametrize("W", [16, 32, 64])
def test_tc2_windows_layout_and_reprojection(W):
    """MSA.forward's pre-partitioned input (LAYOUT_WINDOWS) and the backward that re-projects q, k, v instead of
    keeping them give the same numbers as the default path."""
    from sl_hwgat_b200 import ops
    d, h, B, F = 256, 4, 2, 8
    xn, w, b, g = rounded(*seeded(d, 0, W, B=B, F=F, std=0.05))
    y0, dx0, dw0, db0 = run_tc2(xn, w, b, g, h, 0, 0.05, W)
    y1, dx1, dw1, db1 = run_tc2(xn, w, b, g, h, 0, 0.05, W, save_qkv=False)
    assert torch.equal(y0, y1) and torch.equal(dx0, dx1) and rel_l2(dw1, dw0) < 1e-5
    bits = dev_bits(W, F, 0)
    xw = O.window_partition(xn.to("cuda", torch.bfloat16), W, 2).contiguous().requires_grad_(True)
    yw = ops.window_graph_attention(xw, w.float().cuda(), b.float().cuda(), bits, h, layout=1, frames=F, kps=64,
                                    threshold=0.05, window=W, impl="tc2")
    assert torch.equal(O.window_reverse(yw.detach(), W, 2, F, 64), y0)
    yw.backward(O.window_partition(g.to("cuda", torch.bfloat16), W, 2).contiguous())
    assert torch.equal(O.window_reverse(xw.grad, W, 2, F, 64), dx0)


def test_tc2_matches_fused_kernels_at_reference_window():
    """W = 16: K2b / K3b against K2 / K3 (both bf16, different summation orders): within bf16 rounding of each other."""
    d, h, B, F = 256, 4, 3, 16
    xn, w, b, g = rounded(*seeded(d, 1, 16, B=B, F=F, std=0.05))
    a = run_tc2(xn, w, b, g, h, 1, None, 16, impl="tc2")
    f = run_tc2(xn, w, b, g, h, 1, None, 16, impl="fused")
    for x, y in zip(a, f):
        assert rel_l2(x, y) < 1e-2


# ------------------------------------------------------------------ the drop-in model with larger windows
def build_w(W, T=16, classes=10, drop=0.0):
    from sl_hwgat_b200.models import HWGATE, model_params
    p = model_params.HWGATEParams({"num_class": classes, "src_len": T}, 2, "cuda")
    p.drop_rate = drop
    p.set_window_size(W)
    torch.manual_seed(0)
    m = HWGATE.Model(*p.get_model_params())
    cfg = O.HWGATEConfig(temporal_dim=T, num_classes=classes, window_size=W, edges=EDGES[:64 // W])
    sd = O.make_state_dict(cfg, seed=1001, weight_std=0.05)
    m.load_state_dict(sd, strict=True)            # same names and shapes as the reference at this window size
    return m.cuda(), cfg, sd


@pytest.mark.parametrize("W", [32, 64])
def test_model_larger_windows_vs_reference_golden(golden_dir, W):
    G = np.load(os.path.join(golden_dir, "larger_windows.npz"))
    m, cfg, sd = build_w(W)
    x = O.synthetic_keypoints(2, 16, 2, seed=1001).cuda()
    m.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ev = m(x).float()
    assert rel_l2(ev, torch.from_numpy(G[f"model_W{W}_eval_logits"])) < BF16_TOL
    m.train()
    thr = [0.03, 0.05, 0.031, 0.2, 0.033, 0.04, 0.0312, 0.1]
    it = iter(thr)
    real = torch.rand
    torch.rand = lambda *a, **k: torch.tensor([next(it)])
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            tr = m(x).float()
    finally:
        torch.rand = real
    assert rel_l2(tr, torch.from_numpy(G[f"model_W{W}_train_logits"])) < 3e-2   # train: threshold flips, see DESIGN 2


@pytest.mark.parametrize("W", [32, 64])
def test_model_larger_windows_gradients_vs_oracle(W):
    m, cfg, sd = build_w(W)
    m.train()
    x = O.synthetic_keypoints(2, 16, 2, seed=1001).cuda()
    y = O.synthetic_labels(2, 10, seed=1001).cuda()
    thr = [0.03, 0.05, 0.031, 0.2, 0.033, 0.04, 0.0312, 0.1]
    it = iter(thr)
    real = torch.rand
    torch.rand = lambda *a, **k: torch.tensor([next(it)])
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = m(x)
            loss = O.smoothed_cross_entropy(logits.float(), y)
    finally:
        torch.rand = real
    loss.backward()
    frozen = ("B", "pos_encoder.pe")
    sd64 = {k: v.cuda().double().requires_grad_(k not in frozen and not k.endswith("attn_mask")) for k, v in sd.items()}
    ref = O.model_forward(x.double(), sd64, cfg, thresholds=thr)
    O.smoothed_cross_entropy(ref, y).backward()
    assert rel_l2(logits.float(), ref.detach()) < 3e-2
    errs = {n: rel_l2(p.grad, sd64[n].grad) for n, p in m.named_parameters() if p.grad is not None}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print(f"W={W} worst gradient rel_l2:", worst)
    assert worst[0][1] < 8e-2, worst


def test_unsupported_window_is_refused():
    """windows other than 16 / 32 / 64 raise instead of falling back"""
    from sl_hwgat_b200 import _lib, ops
    xn, w, b, g = seeded(128, 0, 32)
    with pytest.raises(_lib.HwgatError):
        ops.window_graph_attention(xn.to(torch.bfloat16).cuda(), w.float().cuda(), b.float().cuda(), dev_bits(32, 4, 0),
                                   2, window=8)


# ------------------------------------------------------------------ fp32 parity mode for window_size 32 / 64
def run_f32(xn, w, b, g, heads, shift, thr, W, layout=0):
    from sl_hwgat_b200 import ops
    F = xn.shape[1]
    bits = dev_bits(W, F, shift)
    x_ = xn.float().cuda().requires_grad_(True)
    w_ = w.float().cuda().requires_grad_(True)
    b_ = b.float().cuda().requires_grad_(True)
    y = ops.window_graph_attention(x_, w_, b_, bits, heads, shift=shift, threshold=thr, layout=layout, window=W)
    assert y.dtype == torch.float32
    y.backward(g.float().cuda())
    return y.detach(), x_.grad, w_.grad, b_.grad


@pytest.mark.parametrize("W", [32, 64])
@pytest.mark.parametrize("d,h", [(128, 2), (256, 4)])
@pytest.mark.parametrize("shift", [0, 1])
@pytest.mark.parametrize("thr", [None, 0.04])
def test_f32_larger_window_attention_vs_reference_golden(golden_dir, W, d, h, shift, thr):
    """attn_win_f32.cu against the unmodified reference (fp64) run with window_size W, eval and train (threshold drop):
    1e-5, the north_star's fp32 tolerance"""
    G = np.load(os.path.join(golden_dir, "larger_windows.npz"))
    key = f"W{W}_d{d}_s{shift}_thr{thr}"
    xn, w, b, g = seeded(d, shift, W)
    y, dx, dw, db = run_f32(xn, w, b, g, h, shift, thr, W)

    def chk(t, name, stride):
        a = t.detach().double().cpu().reshape(-1).numpy()[::stride]
        ref = G[key + "_" + name]
        return float(np.abs(a - ref).max() / np.abs(ref).max())
    errs = (chk(y, "y", 61), chk(dx, "dx", 61), chk(dw, "dw", 251), chk(db, "db", 1))
    print(key, "fp32 max-rel y/dx/dw/db:", errs)
    assert max(errs) < 1e-5, errs


@pytest.mark.parametrize("W", [32, 64])
def test_f32_larger_window_fully_masked_rows_and_windows_layout(W):
    """a threshold below 1/N drops every logit (uniform rows); LAYOUT_WINDOWS equals LAYOUT_BFKD on the partitioned
    tensor"""
    from sl_hwgat_b200 import ops
    xn, w, b, g = seeded(128, 0, W, std=0.02)
    y, dx, dw, db = run_f32(xn, w, b, g, 2, 0, 0.001, W)
    by, bdx, bdw, bdb = oracle_points(xn, w, b, g, 2, 4, 0, 0.001, W, bf16_points=False)
    # dead rows: no logit gradient (dQ = dK = 0); the value path (uniform P) still carries one
    assert rel_inf(y, by) < 1e-5 and rel_inf(dx, bdx) < 1e-5 and rel_inf(dw, bdw) < 1e-5
    xn, w, b, g = seeded(128, 0, W, B=2)
    y0, dx0, dw0, db0 = run_f32(xn, w, b, g, 2, 0, 0.04, W)
    xw, gw = O.window_partition(xn, W, 2), O.window_partition(g, W, 2)
    F = xn.shape[1]
    bits = dev_bits(W, F, 0)
    x_ = xw.float().cuda().requires_grad_(True)
    yw = ops.window_graph_attention(x_, w.float().cuda(), b.float().cuda(), bits, 2, threshold=0.04,
                                    layout=ops.LAYOUT_WINDOWS, frames=F, kps=64, window=W)
    yw.backward(gw.float().cuda())
    assert torch.equal(O.window_partition(y0.cpu().double(), W, 2).float(), yw.detach().cpu())
    assert torch.equal(O.window_partition(dx0.cpu().double(), W, 2).float(), x_.grad.cpu())


@pytest.mark.parametrize("W", [32, 64])
def test_f32_model_larger_windows_vs_reference_golden(golden_dir, W):
    """the drop-in model with window_size W called without autocast: the fp32 parity kernels end to end"""
    G = np.load(os.path.join(golden_dir, "larger_windows.npz"))
    m, cfg, sd = build_w(W)
    x = O.synthetic_keypoints(2, 16, 2, seed=1001).cuda()
    m.eval()
    with torch.no_grad():
        ev = m(x)
    assert ev.dtype == torch.float32
    assert rel_inf(ev, torch.from_numpy(G[f"model_W{W}_eval_logits"])) < 1e-5
    m.train()
    thr = [0.03, 0.05, 0.031, 0.2, 0.033, 0.04, 0.0312, 0.1]
    it = iter(thr)
    real = torch.rand
    torch.rand = lambda *a, **k: torch.tensor([next(it)])
    try:
        tr = m(x)
    finally:
        torch.rand = real
    assert rel_inf(tr, torch.from_numpy(G[f"model_W{W}_train_logits"])) < 1e-5


# ------------------------------------------------------------------ attention dropout (self.attn_drop, HWGATE.py:112)
def _revealing_inputs(W, d, h, B, F, seed=0):
    """Inputs whose attention OUTPUT is the probability matrix: token j of a window carries a one-hot of its position
    in the window (xn), the value projection copies that one-hot (v_j = e_j), so O[i, head, :N'] = P'[i, :N']."""
    N = 2 * W
    assert N <= 64 and d >= N
    rng = np.random.default_rng(seed)
    xn = torch.zeros(B, F, 64, d, dtype=torch.float64)
    for fr in range(F):
        for k in range(64):
            xn[:, fr, k, (fr % 2) * W + k % W] = 1.0
    xn[..., N:] = torch.from_numpy(rng.standard_normal((B, F, 64, d - N))) * 0.5      # makes the logits differ
    w = torch.zeros(3 * d, d, dtype=torch.float64)
    w[:2 * d, N:] = torch.from_numpy(rng.standard_normal((2 * d, d - N))) * 0.2        # q, k from the random part
    for hh in range(h):
        for e in range(N):
            w[2 * d + hh * 64 + e, e] = 1.0                                            # v_j = one-hot(position of j)
    b = torch.zeros(3 * d, dtype=torch.float64)
    return xn, w, b


@pytest.mark.parametrize("W", [16, 32])
def test_attention_dropout_mask_rate_and_gradients(W):
    """attn_drop > 0 (K2b / K3b): the dropped probabilities are exactly 0, the kept ones P / (1-p); the rate is p;
    the mask is reproducible from the seed and differs between rows and heads; and the backward, checked against
    oracle autograd with the recovered mask, regenerates the same mask."""
    from sl_hwgat_b200 import ops
    d, h, B, F, N, p_drop = 128, 2, 4, 8, 2 * W, 0.3
    xn, w, b = _revealing_inputs(W, d, h, B, F)
    bits = dev_bits(W, F, 0)
    g = torch.from_numpy(np.random.default_rng(1).standard_normal((B, F, 64, d))).to(torch.bfloat16)

    def run(p):
        x_ = xn.to("cuda", torch.bfloat16).requires_grad_(True)
        w_ = w.float().cuda().requires_grad_(True)
        b_ = b.float().cuda().requires_grad_(True)
        torch.manual_seed(7); torch.cuda.manual_seed(7)
        y = ops.window_graph_attention(x_, w_, b_, bits, h, shift=0, threshold=None, window=W, attn_drop=p)
        y.backward(g.cuda())
        return y.detach().float().cpu(), x_.grad.float().cpu(), w_.grad.cpu()

    y0, _, _ = run(0.0)
    y1, dx1, dw1 = run(p_drop)
    y1b, _, _ = run(p_drop)
    assert torch.equal(y1, y1b)                                        # same seed, same mask
    # O -> P: windows order, per head the first N channels
    def probs(y):
        yw = O.window_partition(y, W, 2)                               # (B_, N, d)
        return torch.stack([yw[:, :, hh * 64: hh * 64 + N] for hh in range(h)], 1)     # (B_, h, N, N)
    P0, P1 = probs(y0), probs(y1)
    live = P0 > 1e-3
    kept = P1 != 0
    rate = 1.0 - (kept & live).sum().item() / live.sum().item()
    assert abs(rate - p_drop) < 0.01, rate
    scale = 1.0 / (1.0 - p_drop)
    err = ((P1 - P0 * scale).abs()[kept & live] / (P0 * scale)[kept & live]).max().item()
    assert err < 1.5e-2, err                                           # two bf16 roundings
    assert not torch.equal(kept[:, 0], kept[:, 1]) and not torch.equal(kept[0], kept[1])
    # backward against oracle autograd with the recovered mask
    mask = O.combined_mask(adj_w(W), F, W, 2, 0)
    drop = kept.double() * scale
    x_, w_, b_ = (t.clone().requires_grad_(True) for t in rounded(xn, w, b, g.double())[:3])
    yo = O.attention_core(x_, w_, b_, h, mask, W, 2, 0, None, bf16_points=True, drop_mask=drop)
    (yo * g.double()).sum().backward()
    assert rel_l2(y1, yo.detach()) < BF16_TOL
    assert rel_l2(dx1, x_.grad) < BF16_TOL and rel_l2(dw1, w_.grad) < BF16_TOL


def test_model_with_attention_dropout_trains():
    """attn_drop_rate != 0 through the drop-in model (the reference passes it at model_params.py:256): runs on
    K2b / K3b in train mode, is the identity in eval mode."""
    from sl_hwgat_b200.models import HWGATE, model_params
    p = model_params.HWGATEParams({"num_class": 10, "src_len": 16}, 2, "cuda")
    p.drop_rate, p.attn_drop_rate = 0.0, 0.2
    torch.manual_seed(0)
    m = HWGATE.Model(*p.get_model_params()).cuda()
    p.attn_drop_rate = 0.0
    torch.manual_seed(0)
    m0 = HWGATE.Model(*p.get_model_params()).cuda()
    m0.load_state_dict(m.state_dict())
    x = O.synthetic_keypoints(2, 16, 2, seed=3).cuda()
    m.eval(); m0.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.equal(m(x), m0(x))
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x)
    out.float().sum().backward()
    assert torch.isfinite(out).all()
    assert all(torch.isfinite(q.grad).all() for q in m.parameters() if q.grad is not None)


@pytest.mark.parametrize("W", [16, 64])
def test_tc2_128_keypoints(W):
    """K = 128 keypoints: a temporal group spans two tiles (W = 16: 8 windows) or two whole-tile windows (W = 64)."""
    from sl_hwgat_b200 import ops
    d, h, B, F, K = 128, 2, 2, 8, 128
    edges = [EDGES[0]] * (K // W)
    adj = O.window_adjacency(edges, W, 2)
    rng = np.random.default_rng(77 + W)
    xn = torch.from_numpy(rng.standard_normal((B, F, K, d))).to(torch.bfloat16).double()
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * 0.05).to(torch.bfloat16).double()
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1).float().double()
    g = torch.from_numpy(rng.standard_normal((B, F, K, d))).to(torch.bfloat16).double()
    bits = ops.mask_build(torch.from_numpy(adj.astype(np.float32)).cuda(), F, 1, W, 2)
    x_ = xn.to("cuda", torch.bfloat16).requires_grad_(True)
    w_ = w.float().cuda().requires_grad_(True)
    b_ = b.float().cuda().requires_grad_(True)
    y = ops.window_graph_attention(x_, w_, b_, bits, h, shift=1, threshold=0.04, window=W, impl="tc2")
    y.backward(g.to("cuda", torch.bfloat16))
    mask = O.combined_mask(adj, F, W, 2, 1)
    xo, wo, bo = (t.clone().requires_grad_(True) for t in (xn, w, b))
    yo = O.attention_core(xo, wo, bo, h, mask, W, 2, 1, 0.04, bf16_points=True)
    (yo * g).sum().backward()
    assert rel_l2(y, yo.detach()) < BF16_TOL and rel_l2(x_.grad, xo.grad) < BF16_TOL
    assert rel_l2(w_.grad, wo.grad) < BF16_TOL and rel_l2(b_.grad, bo.grad) < BF16_TOL


def test_tc2_empty_batch():
    from sl_hwgat_b200 import ops
    bits = dev_bits(32, 4, 0)
    w = torch.zeros(384, 128, device="cuda", requires_grad=True)
    b = torch.zeros(384, device="cuda", requires_grad=True)
    x = torch.empty(0, 4, 64, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = ops.window_graph_attention(x, w, b, bits, 2, window=32)
    assert y.shape == (0, 4, 64, 128)
    y.sum().backward()
    assert float(w.grad.abs().sum()) == 0.0 and float(b.grad.abs().sum()) == 0.0


# ------------------------------------------------------------------ hybrid: K2 keeps q, k, v; backward on K3b
@pytest.mark.parametrize("d,h", [(128, 2), (256, 4), (512, 8)])
@pytest.mark.parametrize("shift", [0, 1])
@pytest.mark.parametrize("thr", [None, 0.04])
def test_hybrid_attention_vs_oracle_and_fused(d, h, shift, thr):
    """impl="hybrid": the fused forward (K2) also stores the q, k, v rows it formed (q, k column-permuted inside each
    head), and the backward is K3b's tcgen05 core on them.  Forward bit-identical to the fused kernels, gradients
    within the bf16 tolerance of the oracle."""
    B, F = 3, 8
    xn, w, b, g = rounded(*seeded(d, shift, 16, B=B, F=F, std=0.05))
    yh, dxh, dwh, dbh = run_tc2(xn, w, b, g, h, shift, thr, 16, impl="hybrid")
    yf, dxf, dwf, dbf = run_tc2(xn, w, b, g, h, shift, thr, 16, impl="fused")
    assert torch.equal(yh, yf)
    by, bdx, bdw, bdb = oracle_points(xn, w, b, g, h, F, shift, thr, 16)
    errs = dict(dx=rel_l2(dxh, bdx), dw=rel_l2(dwh, bdw), db=rel_l2(dbh, bdb))
    assert all(e < BF16_TOL for e in errs.values()), errs
    assert rel_l2(dxh, dxf) < 1e-2 and rel_l2(dwh, dwf) < 1e-2


def test_hybrid_multi_tile_replication():
    d, h, F = 512, 8, 16
    xn, w, b, g = rounded(*seeded(d, 1, 16, B=1, F=F, std=0.05))
    ys, dxs, dws, dbs = run_tc2(xn, w, b, g, h, 1, 0.05, 16, impl="hybrid")
    for Bbig in (64, 37):
        yb, dxb, dwb, dbb = run_tc2(xn.expand(Bbig, -1, -1, -1).contiguous(), w, b,
                                    g.expand(Bbig, -1, -1, -1).contiguous(), h, 1, 0.05, 16, impl="hybrid")
        assert torch.equal(yb, ys.expand(Bbig, -1, -1, -1)) and torch.equal(dxb, dxs.expand(Bbig, -1, -1, -1))
        assert rel_l2(dwb, dws * Bbig) < 1e-3 and rel_l2(dbb, dbs * Bbig) < 1e-3
