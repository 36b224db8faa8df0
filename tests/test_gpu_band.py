"""The sibling models WGATE and GATE (hwgat/models/WGATE.py, GATE.py) on the frame-banded
attention kernels K15 / K16.  The reference (and the oracle, oracle/wgate_oracle.py, pinned to it on the CPU) attends
DENSELY over all frames with an additive -10000 mask; the kernels evaluate the graph's 3-frame band only.  Parity
against outputs of the unmodified reference (tests/golden/wgate_gate.npz) and the fp64 dense oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import hwgate_oracle as O
from oracle import wgate_oracle as WG
from tests._util import rel_inf, rel_l2

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _params(name, T, classes, drop=0.0, depths=None):
    from sl_hwgat_b200.models import model_params
    cls = model_params.WGATEParams if name == "wgate" else model_params.GATEParams
    p = cls({"num_class": classes, "src_len": T}, 2, "cuda")
    p.drop_rate = drop
    if depths is not None:
        p.depths = depths
    return p


def build(name, T, classes=10, drop=0.0, depths=3):
    from sl_hwgat_b200.models import GATE, WGATE
    p = _params(name, T, classes, drop, depths)
    torch.manual_seed(0)
    m = (WGATE if name == "wgate" else GATE).Model(*p.get_model_params())
    cfg = (WG.WGATEConfig if name == "wgate" else WG.GATEConfig)(temporal_dim=T, num_classes=classes, depths=depths)
    sd = WG.make_state_dict(cfg, seed=1001, weight_std=0.05)
    m.load_state_dict(sd, strict=True)          # the reference's names and shapes (incl. the adj_mask buffer)
    return m.cuda(), cfg, sd, p


def core_inputs(name, d, h, B=2):
    """Same seeded inputs as tests/golden/make_golden.py section 8."""
    F, K = (8, 64) if name == "wgate" else (6, 29)
    std = 0.2 if d == 128 else 0.1
    rng = np.random.default_rng(6000 + d + h + (0 if name == "wgate" else 1))
    xn = torch.from_numpy(rng.standard_normal((B, F, K, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, K, d)))
    return xn, w, b, g


def dense_mask(name, F):
    if name == "wgate":
        return WG.additive_mask(WG.wgate_adjacency(WG.WGATEConfig().edges, F, 16))
    return WG.additive_mask(WG.gate_adjacency(WG.GATEConfig().edges, F, 29))


def cuda_core(name, xn, w, b, g, heads):
    """the product op on the (B, F, K, d) stream (GATE: keypoints padded to 32), gradients included"""
    from sl_hwgat_b200 import ops
    from sl_hwgat_b200.models.HGATE import _pad_kp
    B, F, K, d = xn.shape
    mask = torch.from_numpy(dense_mask(name, F)).float().cuda()
    W = 16 if name == "wgate" else 32
    bits = ops.band_mask_pack(mask, F, W)
    x_ = xn.float().cuda().requires_grad_(True)
    w_ = w.float().cuda().requires_grad_(True)
    b_ = b.float().cuda().requires_grad_(True)
    xp = _pad_kp(x_, 2) if name == "gate" else x_
    y = ops.band_graph_attention(xp.to(torch.bfloat16), w_, b_, bits, heads, W)[:, :, :K]
    (y.float() * g.float().cuda()).sum().backward()
    return y.detach().double().cpu(), x_.grad.double().cpu(), w_.grad.double().cpu(), b_.grad.double().cpu()


@pytest.mark.parametrize("name", ["wgate", "gate"])
@pytest.mark.parametrize("d,h", [(128, 8), (128, 2), (256, 8)])
def test_band_attention_vs_reference_golden_and_dense_oracle(golden_dir, name, d, h):
    """head dims 16, 64 and 32; forward and all three gradients"""
    G = np.load(os.path.join(golden_dir, "wgate_gate.npz"))
    key = f"{name}_d{d}_h{h}"
    xn, w, b, g = core_inputs(name, d, h)
    y, dx, dw, db = cuda_core(name, xn, w, b, g, h)

    def chk(t, nm, stride):
        a = t.reshape(-1).numpy()[::stride]
        ref = G[key + "_" + nm]
        return float(np.linalg.norm(a - ref) / np.linalg.norm(ref))
    assert chk(y, "y", 53) < BF16_TOL and chk(dx, "dx", 53) < BF16_TOL and chk(dw, "dw", 251) < BF16_TOL
    assert float(np.linalg.norm(db.numpy() - G[key + "_db"]) / np.linalg.norm(G[key + "_db"])) < BF16_TOL
    # the whole tensors against the fp64 dense oracle on the kernel's own bf16-rounded inputs
    rb = lambda t: t.float().bfloat16().double()
    x_, w_, b_ = rb(xn).requires_grad_(True), rb(w).requires_grad_(True), b.float().double().requires_grad_(True)
    mask = dense_mask(name, xn.shape[1])
    ref = (WG.wgate_attention_core(x_, w_, b_, h, mask, 16) if name == "wgate"
           else WG.gate_attention_core(x_, w_, b_, h, mask))
    (ref * g.float().double()).sum().backward()
    errs = (rel_l2(y, ref.detach()), rel_l2(dx, x_.grad), rel_l2(dw, w_.grad), rel_l2(db, b_.grad))
    print(key, "rel_l2 y/dx/dw/db:", errs)
    assert max(errs) < BF16_TOL, errs
    # per-token: no row may be wrong (a mis-indexed neighbour frame would hide in an aggregate)
    row_err = (y - ref.detach()).norm(dim=-1) / ref.detach().norm(dim=-1).clamp_min(1e-6)
    assert float(row_err.max()) < 8e-2, float(row_err.max())


@pytest.mark.parametrize("name,F,B", [("wgate", 1, 2), ("wgate", 2, 1), ("wgate", 13, 2), ("gate", 1, 4), ("gate", 7, 4),
                                      ("wgate", 64, 3), ("gate", 64, 2)])
def test_band_attention_frame_edges_and_ragged_chunks(name, F, B):
    """F = 1 (no neighbour frames), F = 2, F not a multiple of the 8 / 4 frames a CTA owns, and the full T = 64"""
    d, h = 128, 8
    K = 64 if name == "wgate" else 29
    rng = np.random.default_rng(40 + F)
    xn = torch.from_numpy(rng.standard_normal((B, F, K, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * 0.2)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, K, d)))
    y, dx, dw, db = cuda_core(name, xn, w, b, g, h)
    rb = lambda t: t.float().bfloat16().double()
    dev = "cuda" if F >= 64 else "cpu"        # the dense oracle at T = 64 (N = 1024 / 1856) runs in fp64 on the box's GPU
    x_, w_ = rb(xn).to(dev).requires_grad_(True), rb(w).to(dev).requires_grad_(True)
    b_ = b.float().double().to(dev).requires_grad_(True)
    mask = dense_mask(name, F)
    ref = (WG.wgate_attention_core(x_, w_, b_, h, mask, 16) if name == "wgate"
           else WG.gate_attention_core(x_, w_, b_, h, mask))
    (ref * g.float().double().to(dev)).sum().backward()
    errs = (rel_l2(y, ref.detach().cpu()), rel_l2(dx, x_.grad.cpu()), rel_l2(dw, w_.grad.cpu()), rel_l2(db, b_.grad.cpu()))
    assert max(errs) < BF16_TOL, errs


@pytest.mark.parametrize("W,K,hd", [(16, 64, 16), (32, 32, 16), (16, 32, 64), (32, 64, 32)])
def test_band_attention_general_band_and_forced_general_path(W, K, hd):
    """(1) An adjacency whose blocks between adjacent frames are NOT the identity (a shifted identity plus random
    links) - the kernels' general path - against the dense additive-mask oracle.  (2) A graph with the identity
    between frames: the diagonal fast path gives the same result as the general path forced on it."""
    from sl_hwgat_b200 import ops
    F, B, heads = 5, 4, 128 // hd
    d = heads * hd
    rng = np.random.default_rng(77 + W + hd)
    same = np.eye(W)
    for i, j in [(0, 1), (0, 2), (1, 5), (3, 4), (4, 9), (7, W - 1), (W - 2, W - 1)]:
        same[i, j] = same[j, i] = 1
    nxt = np.roll(np.eye(W), 1, axis=1) + (rng.random((W, W)) < 0.1)            # to the next frame: not the identity
    prv = np.eye(W) * (np.arange(W) % 3 != 0)[:, None] + (rng.random((W, W)) < 0.1)
    fr = np.arange(F)
    dt = fr[None, :] - fr[:, None]                                               # key frame - query frame

    def band(nx, pv):
        blocks = (np.where((dt == 0)[:, None, :, None], same[None, :, None, :], 0) +
                  np.where((dt == 1)[:, None, :, None], nx[None, :, None, :], 0) +
                  np.where((dt == -1)[:, None, :, None], pv[None, :, None, :], 0))
        return (blocks.reshape(F * W, F * W) != 0).astype(np.float64)
    nW = K // W
    mask = WG.additive_mask(np.stack([band(nxt, prv)] * nW))
    bits = ops.band_mask_pack(torch.from_numpy(mask).float().cuda(), F, W)
    assert bits.band_diag is False
    xn = torch.from_numpy(rng.standard_normal((B, F, K, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * 0.2)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, K, d)))

    def run(bits_, diag=None):
        x_ = xn.float().cuda().requires_grad_(True)
        w_ = w.float().cuda().requires_grad_(True)
        b_ = b.float().cuda().requires_grad_(True)
        y = ops.band_graph_attention(x_.to(torch.bfloat16), w_, b_, bits_, heads, W, diag)
        (y.float() * g.float().cuda()).sum().backward()
        return [t.detach().double().cpu() for t in (y, x_.grad, w_.grad, b_.grad)]

    got = run(bits)
    rb = lambda t: t.float().bfloat16().double()
    x_, w_, b_ = rb(xn).requires_grad_(True), rb(w).requires_grad_(True), b.float().double().requires_grad_(True)
    ref = WG.wgate_attention_core(x_, w_, b_, heads, mask, W)
    (ref * g.float().double()).sum().backward()
    errs = [rel_l2(a, r) for a, r in zip(got, (ref.detach(), x_.grad, w_.grad, b_.grad))]
    assert max(errs) < BF16_TOL, errs
    # (2) identity between frames: fast path == general path (the same fp32 values on the live entries)
    bits_i = ops.band_mask_pack(torch.from_numpy(WG.additive_mask(np.stack([band(np.eye(W), np.eye(W))] * nW))).float().cuda(), F, W)
    assert bits_i.band_diag is True
    fast, general = run(bits_i), run(bits_i, diag=False)
    for a, c in zip(fast, general):
        assert rel_l2(a, c) < 2e-3          # identical up to the summation order inside the MMAs / bf16 rounding


def test_band_diag_promise_is_checked_on_the_device():
    """diag = 1 with words that are not diagonal must not silently compute something else: the CTAs trap."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import torch, numpy as np, sys
        sys.path.insert(0, %r)
        from sl_hwgat_b200 import ops
        from oracle import wgate_oracle as WG
        F, W = 4, 16
        adj = WG.wgate_adjacency(WG.WGATEConfig().edges, F, W)
        for f in range(1, F):                     # keypoint 3 also sees keypoint 5 of the previous frame, in every frame
            adj[:, f * W + 3, (f - 1) * W + 5] = 1
        bits = ops.band_mask_pack(torch.from_numpy(WG.additive_mask(adj)).float().cuda(), F, W)
        assert bits.band_diag is False
        x = torch.randn(2, F, 64, 128, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(384, 128, device="cuda") * 0.1
        b = torch.zeros(384, device="cuda")
        try:
            ops.band_graph_attention(x, w, b, bits, 8, W, diag=True)
            torch.cuda.synchronize()
        except Exception as e:
            print("TRAPPED", type(e).__name__)
            sys.exit(0)
        print("NO TRAP")
    """ % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "TRAPPED" in r.stdout, (r.stdout, r.stderr[-500:])


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_band_model_params_and_state_dict_match_reference(golden_dir, name):
    G = np.load(os.path.join(golden_dir, "wgate_gate.npz"))
    F = 8 if name == "wgate" else 6
    m, cfg, sd, p = build(name, F)
    assert np.array_equal(p.adj_mat.numpy(), G[name + "_adj"].astype(np.float32))       # reference get_adj_mat / get_adj
    assert len(p.get_model_params()) == (15 if name == "wgate" else 14)
    assert sorted(m.state_dict().keys()) == sorted(G[name + "_state_dict_names"])
    shapes = dict(zip(G[name + "_state_dict_names"], G[name + "_state_dict_shapes"]))
    assert all(str(tuple(v.shape)) == shapes[k] for k, v in m.state_dict().items())


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_band_model_vs_reference_golden_and_oracle_gradients(golden_dir, name):
    G = np.load(os.path.join(golden_dir, "wgate_gate.npz"))
    F, K = (8, 64) if name == "wgate" else (6, 29)
    m, cfg, sd, p = build(name, F)
    m.train()                                    # drop 0: neither model has a threshold path, train == eval numerics
    x = WG.synthetic_keypoints(2, F, K, seed=1001).cuda()
    y = torch.tensor([3, 7]).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(x)
        loss = O.smoothed_cross_entropy(logits.float(), y)
    loss.backward()
    assert rel_l2(logits.float(), torch.from_numpy(G[name + "_model_logits"])) < BF16_TOL
    assert abs(loss.item() - float(G[name + "_model_loss"])) < 1e-2 * abs(float(G[name + "_model_loss"]))
    sd64 = {k: v.double().requires_grad_(k not in ("B", "pos_encoder.pe", "adj_mask")) for k, v in sd.items()}
    ref = (WG.wgate_forward if name == "wgate" else WG.gate_forward)(x.cpu().double(), sd64, cfg)
    O.smoothed_cross_entropy(ref, y.cpu()).backward()
    errs = {n: rel_l2(q.grad.cpu(), sd64[n].grad) for n, q in m.named_parameters() if q.grad is not None}
    assert set(errs) == {k for k, v in sd64.items() if v.grad is not None}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print(name, "worst gradient rel_l2:", worst)
    assert worst[0][1] < 5e-2, worst


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_band_model_reference_signatures_and_train_step(name):
    """Block.forward(x, parent) / MSA.forward(...) with the reference's argument conventions; a train step with dropout
    on the default 8-block model at T = 64; eval at batch 1 (inference.py:95) against the dense oracle."""
    from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss
    from sl_hwgat_b200.models import GATE, WGATE
    from sl_hwgat_b200.optim import AdamW
    F, K = (8, 64) if name == "wgate" else (8, 29)
    m, cfg, sd, p = build(name, F)
    m.eval()
    x = WG.synthetic_keypoints(4, F, K, seed=9).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        full = m(x).float()
        h = torch.randn(4, F, K, 128, device="cuda")
        blk = m.layers[0]
        if name == "wgate":
            y_blk = blk(h, m)                                                                # WGATE.py:251
            y_msa = blk.attn(WGATE.window_partition(blk.norm1(h), 16), 4, 4, m)               # WGATE.py:155-156
            a = WGATE.window_reverse(y_msa, 16, F, K)
        else:
            y_blk = blk(h.reshape(4, F * K, 128), m).reshape(4, F, K, 128)                    # GATE.py:200
            a = blk.attn(blk.norm1(h).reshape(4, F * K, 128), m).reshape(4, F, K, 128)        # GATE.py:112
        h2 = h + a
        y_ref = h2 + blk.ff(blk.norm2(h2))
    assert rel_l2(y_blk.float(), y_ref.float()) < BF16_TOL
    ref = (WG.wgate_forward if name == "wgate" else WG.gate_forward)(
        x.cpu().double(), {k: v.double() for k, v in sd.items()}, cfg)
    assert rel_l2(full, ref) < BF16_TOL
    # the default model (8 blocks, dropout 0.1), one optimiser step: loss finite and decreasing on a fixed batch
    m2, cfg2, sd2, _ = build(name, 64, classes=262, drop=0.1, depths=8)
    m2.train()
    opt = AdamW(m2.parameters(), lr=1e-3)
    crit = SmoothedCrossEntropyLoss(0.1)
    xb = WG.synthetic_keypoints(8, 64, K, seed=3).cuda()
    yb = torch.arange(8, device="cuda") % 262
    losses = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = crit(m2(xb), yb)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and min(losses[1:]) < losses[0], losses
    m2.eval()
    x1 = WG.synthetic_keypoints(4, 64, K, seed=5).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out = m2(x1).float()
    sd_now = {k: v.detach().double() for k, v in m2.state_dict().items()}
    ref1 = (WG.wgate_forward if name == "wgate" else WG.gate_forward)(x1.double(), sd_now, cfg2)   # fp64 on the GPU
    assert rel_l2(out, ref1) < BF16_TOL


def test_band_mask_pack_refuses_what_the_kernels_cannot_do():
    from sl_hwgat_b200 import _lib, ops
    F = 6
    good = torch.from_numpy(dense_mask("wgate", F)).float().cuda()
    bits = ops.band_mask_pack(good, F, 16)
    assert tuple(bits.shape) == (4, 16, 3) and bits.dtype == torch.int32
    # bit-exact against the dense adjacency: word (w, i, r) = row i of frame block (f, f-1+r)
    e = (good == 0).reshape(4, F, 16, F, 16).cpu().numpy()
    words = bits.cpu().numpy().astype(np.int64) & 0xffffffff
    for r in range(3):
        blk = e[:, 2, :, 1 + r, :]                                    # query frame 2 sees key frames 1, 2, 3
        got = (words[:, :, r, None] >> np.arange(16)) & 1
        assert np.array_equal(got.astype(bool), blk)
    far = good.clone(); far[0, 0, 3 * 16] = 0                            # an edge three frames away
    with pytest.raises(_lib.HwgatError, match="more than one frame"):
        ops.band_mask_pack(far, F, 16)
    var = good.clone(); var[1, 2 * 16 + 1, 2 * 16 + 5] = 0                # frame 2's graph differs from the others
    with pytest.raises(_lib.HwgatError, match="frame to frame"):
        ops.band_mask_pack(var, F, 16)
    val = good.clone(); val[0, 0, 1] = -5.0                              # not 0 / -10000
    with pytest.raises(_lib.HwgatError, match="only 0"):
        ops.band_mask_pack(val, F, 16)
    empty = torch.full_like(good, -10000.0)
    with pytest.raises(_lib.HwgatError, match="without any edge"):
        ops.band_mask_pack(empty, F, 16)


def test_band_models_refuse_cpu_and_attention_dropout():
    from sl_hwgat_b200 import _lib
    m, cfg, sd, p = build("wgate", 8)
    m.eval()
    with pytest.raises(_lib.HwgatError):
        m.cpu()(WG.synthetic_keypoints(2, 8, 64, seed=1))            # no CPU path
    m.cuda().train()
    for blk in m.layers:
        blk.attn.attn_drop.p = 0.1
    with pytest.raises(_lib.HwgatError, match="attention dropout"), torch.autocast("cuda", dtype=torch.bfloat16):
        m(WG.synthetic_keypoints(2, 8, 64, seed=1).cuda())


def test_weighted_pool_vs_torch_fp32():
    """K9 with token weights against LayerNorm + einsum in fp32 (1e-5), padded keypoint axis included."""
    from sl_hwgat_b200 import ops
    torch.manual_seed(0)
    B, F, K, d = 6, 5, 29, 128
    x = torch.randn(B, F, 32, d, device="cuda")
    gamma = (1 + 0.1 * torch.randn(d, device="cuda")).requires_grad_(True)
    beta = (0.1 * torch.randn(d, device="cuda")).requires_grad_(True)
    w = (torch.randn(1, F * K, device="cuda") / (F * K)).requires_grad_(True)
    wb = torch.randn(1, device="cuda").requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    out = ops.layer_norm_weighted_pool(xr, gamma, beta, w, wb, 1e-5, kp_real=K)
    g = torch.randn_like(out)
    (out * g).sum().backward()
    got = [out.detach(), xr.grad, gamma.grad, beta.grad, w.grad, wb.grad]
    x2 = x.double().requires_grad_(True)
    g2, b2, w2, wb2 = (t.detach().double().requires_grad_(True) for t in (gamma, beta, w, wb))
    ln = torch.nn.functional.layer_norm(x2[:, :, :K], (d,), g2, b2, 1e-5).reshape(B, F * K, d)
    ref = torch.einsum("btd,t->bd", ln, w2.reshape(-1)) + wb2
    (ref * g.double()).sum().backward()
    want = [ref.detach(), x2.grad, g2.grad, b2.grad, w2.grad, wb2.grad]
    for a, b_ in zip(got, want):
        assert rel_l2(a.double(), b_) < 1e-5
    assert float(xr.grad[:, :, K:].abs().max()) == 0.0           # padded rows get no gradient


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_band_models_cuda_graph_inference(name):
    """the per-sample evaluator (inference.py:88-95) through runtime.GraphedInference: replays equal the eager forward"""
    from sl_hwgat_b200.runtime import GraphedInference
    K = 64 if name == "wgate" else 29
    m, cfg, sd, p = build(name, 16, classes=50, depths=2)
    m.eval()
    fast = GraphedInference(m)
    for B in (4, 8):
        x = WG.synthetic_keypoints(B, 16, K, seed=B).cuda()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            want = m(x)
        assert torch.equal(fast(x), want)
        x2 = WG.synthetic_keypoints(B, 16, K, seed=B + 100).cuda()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            want2 = m(x2)
        assert torch.equal(fast(x2), want2) and not torch.equal(want, want2)


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_band_attention_full_size_replication(name):
    """The bench shape (B = 512, T = 64: 32768 CTAs per launch) through a size-independent property: every sample is the
    same sequence, so every sample's output and input gradient must equal sample 0's bit for bit, sample 0 must match the
    dense oracle, and the weight gradient must be 512 x the one-sample gradient."""
    from sl_hwgat_b200 import ops
    from sl_hwgat_b200.models.HGATE import _pad_kp
    B, F, d, h = 512, 64, 128, 8
    K, W = (64, 16) if name == "wgate" else (29, 32)
    rng = np.random.default_rng(11)
    x1 = torch.from_numpy(rng.standard_normal((1, F, K, d))).float()
    g1 = torch.from_numpy(rng.standard_normal((1, F, K, d))).float()
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * 0.2).float()
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1).float()
    mask = dense_mask(name, F)
    bits = ops.band_mask_pack(torch.from_numpy(mask).float().cuda(), F, W)
    xp = _pad_kp(x1, 2) if name == "gate" else x1
    gp = _pad_kp(g1, 2) if name == "gate" else g1
    x_ = xp.cuda().to(torch.bfloat16).expand(B, -1, -1, -1).contiguous().requires_grad_(True)
    w_, b_ = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = ops.band_graph_attention(x_, w_, b_, bits, h, W)
    y.backward(gp.cuda().to(torch.bfloat16).expand(B, -1, -1, -1).contiguous())
    assert torch.equal(y, y[:1].expand_as(y)) and torch.equal(x_.grad, x_.grad[:1].expand_as(x_.grad))
    rb = lambda t: t.bfloat16().double()
    xr, wr, br = rb(x1).cuda().requires_grad_(True), rb(w).cuda().requires_grad_(True), b.double().cuda().requires_grad_(True)
    ref = (WG.wgate_attention_core(xr, wr, br, h, mask, 16) if name == "wgate" else WG.gate_attention_core(xr, wr, br, h, mask))
    (ref * rb(g1).cuda()).sum().backward()
    assert rel_l2(y[0, :, :K], ref[0]) < BF16_TOL and rel_l2(x_.grad[0, :, :K], xr.grad[0]) < BF16_TOL
    assert rel_l2(w_.grad / B, wr.grad) < BF16_TOL and rel_l2(b_.grad / B, br.grad) < BF16_TOL


# ------------------------------------------------------------------ fp32 parity mode (no autocast)
FP32_TOL = 1e-5


@pytest.mark.parametrize("name", ["wgate", "gate"])
@pytest.mark.parametrize("d,h", [(128, 8), (128, 2), (256, 8)])
def test_band_attention_fp32_vs_reference_golden(golden_dir, name, d, h):
    """the true-fp32 band kernels against the reference's fp64 outputs (dense attention, additive mask): 1e-5"""
    from sl_hwgat_b200 import ops
    from sl_hwgat_b200.models.HGATE import _pad_kp
    G = np.load(os.path.join(golden_dir, "wgate_gate.npz"))
    key = f"{name}_d{d}_h{h}"
    xn, w, b, g = core_inputs(name, d, h)
    B, F, K, _ = xn.shape
    W = 16 if name == "wgate" else 32
    bits = ops.band_mask_pack(torch.from_numpy(dense_mask(name, F)).float().cuda(), F, W)
    x_ = xn.float().cuda().requires_grad_(True)
    w_ = w.float().cuda().requires_grad_(True)
    b_ = b.float().cuda().requires_grad_(True)
    xp = _pad_kp(x_, 2) if name == "gate" else x_
    y = ops.band_graph_attention(xp, w_, b_, bits, h, W)[:, :, :K]
    assert y.dtype == torch.float32
    (y * g.float().cuda()).sum().backward()

    def chk(t, nm, stride):
        a = t.detach().double().cpu().reshape(-1).numpy()[::stride]
        ref = G[key + "_" + nm]
        return float(np.abs(a - ref).max() / np.abs(ref).max())
    errs = (chk(y, "y", 53), chk(x_.grad, "dx", 53), chk(w_.grad, "dw", 251), chk(b_.grad, "db", 1))
    print(key, "fp32 max-rel y/dx/dw/db:", errs)
    assert max(errs) < FP32_TOL, errs


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_band_model_fp32_vs_reference_golden(golden_dir, name):
    """the drop-in model called without autocast (utils.py:102): logits, loss and every parameter gradient against the
    reference's fp64 run"""
    G = np.load(os.path.join(golden_dir, "wgate_gate.npz"))
    F, K = (8, 64) if name == "wgate" else (6, 29)
    m, cfg, sd, p = build(name, F)
    m.train()                                    # drop 0
    x = WG.synthetic_keypoints(2, F, K, seed=1001).cuda()
    y = torch.tensor([3, 7]).cuda()
    logits = m(x)
    loss = O.smoothed_cross_entropy(logits, y)
    loss.backward()
    assert logits.dtype == torch.float32
    assert rel_inf(logits, torch.from_numpy(G[name + "_model_logits"])) < FP32_TOL
    assert abs(loss.item() - float(G[name + "_model_loss"])) < 1e-5 * abs(float(G[name + "_model_loss"]))
    grads = dict(m.named_parameters())
    worst = 0.0
    for pname, norm, head in zip(G[name + "_gnames"], G[name + "_gnorms"], G[name + "_gheads"]):
        gr = grads[str(pname)].grad.double().cpu()
        worst = max(worst, abs(gr.norm().item() - norm) / norm)
        got = np.zeros(4); got[:min(4, gr.numel())] = gr.reshape(-1)[:4].numpy()
        assert np.abs(got - head).max() <= 1e-4 * max(np.abs(head).max(), 1e-12) + 1e-9, pname
    assert worst < 1e-4, worst
    # and the whole gradients against oracle autograd in fp64
    sd64 = {k: v.double().requires_grad_(k not in ("B", "pos_encoder.pe", "adj_mask")) for k, v in sd.items()}
    ref = (WG.wgate_forward if name == "wgate" else WG.gate_forward)(x.cpu().double(), sd64, cfg)
    O.smoothed_cross_entropy(ref, y.cpu()).backward()
    errs = {n: rel_l2(q.grad.cpu(), sd64[n].grad) for n, q in m.named_parameters() if q.grad is not None}
    assert max(errs.values()) < 1e-4, sorted(errs.items(), key=lambda kv: -kv[1])[:3]
