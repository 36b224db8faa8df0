"""Drop-in module on the GPU against outputs of the unmodified reference
(tests/golden/full_model.npz) and the fp64 oracle."""
import contextlib
import os

import numpy as np
import pytest
import torch

from oracle import hwgate_oracle as O
from tests._util import ADJ, rel_inf, rel_l2

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def patched_rand(values):
    """MSA.forward draws its training threshold with torch.rand(1).item() (HWGATE.py:96)."""
    it = iter(values)
    real = torch.rand

    def fake(*a, **k):
        return torch.tensor([next(it)])
    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = real


def build(T, classes, drop=0.0, std=0.05):
    from sl_hwgat_b200.models import HWGATE, model_params
    p = model_params.HWGATEParams({"num_class": classes, "src_len": T}, 2, "cuda")
    p.drop_rate = drop
    torch.manual_seed(0)
    m = HWGATE.Model(*p.get_model_params())
    cfg = O.HWGATEConfig(temporal_dim=T, num_classes=classes)
    sd = O.make_state_dict(cfg, seed=1001, weight_std=std)
    m.load_state_dict(sd, strict=True)        # same names and shapes as the reference
    return m.cuda(), cfg, sd


def test_params_adjacency_matches_reference(golden_dir):
    from sl_hwgat_b200.models import model_params
    p = model_params.HWGATEParams({"num_class": 262, "src_len": 64}, 2, "cuda")
    g = np.load(os.path.join(golden_dir, "masks.npz"))
    assert p.adj_mat.device.type == "cpu" and p.adj_mat.dtype == torch.float32
    assert np.array_equal(p.adj_mat.numpy(), g["adj"].astype(np.float32))
    assert np.array_equal(p.get_adj(0), g["adj"][0, :16, :16])
    assert len(p.get_model_params()) == 16


def test_state_dict_contract(golden_dir):
    G = np.load(os.path.join(golden_dir, "full_model.npz"))
    m, cfg, sd = build(64, 262)
    assert list(m.state_dict().keys()) == list(G["state_dict_names"])
    assert [str(tuple(v.shape)) for v in m.state_dict().values()] == list(G["state_dict_shapes"])
    # the float attn_mask buffers the module builds itself equal the reference's
    from sl_hwgat_b200.models import HWGATE
    blk = HWGATE.PartAttentionBlock(dim=128, num_kps=64, num_heads=2, window_size=16, temporal_patch_size=2,
                                    temporal_dim=16, shift_size=1)
    assert np.array_equal(blk.attn_mask.numpy() != 0, O.shift_window_mask(16, 4, 16, 2, 1))


def test_model_eval_fp32_matches_reference(golden_dir):
    G = np.load(os.path.join(golden_dir, "full_model.npz"))
    m, cfg, sd = build(64, 262)
    m.eval()
    x = O.synthetic_keypoints(2, 64, 2, seed=1001).cuda()
    with torch.no_grad():
        logits = m(x)
    ref = torch.from_numpy(G["include_eval_logits"])              # reference, fp64
    assert rel_inf(logits, ref) < 1e-5, rel_inf(logits, ref)
    ref32 = torch.from_numpy(G["include_eval_logits_fp32"])        # reference, fp32 on CPU
    assert rel_inf(logits, ref32) < 1e-5


def test_model_batch_one_long_sequence(golden_dir):
    """inference.py:95 calls the model with batch 1; FDMSE-ISL shape T=192, 2002 classes."""
    G = np.load(os.path.join(golden_dir, "full_model.npz"))
    m, cfg, sd = build(192, 2002)
    m.eval()
    x = O.synthetic_keypoints(1, 192, 2, seed=1001).cuda()
    with torch.no_grad():
        logits = m(x)
    assert rel_inf(logits, torch.from_numpy(G["fdmse_eval_logits"])) < 1e-5


def test_model_train_fp32_matches_reference(golden_dir):
    G = np.load(os.path.join(golden_dir, "full_model.npz"))
    m, cfg, sd = build(64, 262, drop=0.0)
    m.train()
    x = O.synthetic_keypoints(2, 64, 2, seed=1001).cuda()
    y = O.synthetic_labels(2, 262, seed=1001).cuda()
    thr = [float(t) for t in G["include_train_thr"]]
    with patched_rand(thr):
        logits = m(x)
    assert rel_inf(logits, torch.from_numpy(G["include_train_logits"])) < 1e-5
    loss = O.smoothed_cross_entropy(logits, y)
    assert abs(loss.item() - float(G["include_train_loss"])) < 1e-5 * abs(float(G["include_train_loss"]))
    loss.backward()
    params = dict(m.named_parameters())
    worst = 0.0
    for name, norm, head in zip(G["include_train_gnames"], G["include_train_gnorms"], G["include_train_gheads"]):
        g = params[str(name)].grad
        assert g is not None, name
        worst = max(worst, abs(g.norm().item() - norm) / norm)
        assert abs(g.norm().item() - norm) / norm < 1e-4, (name, g.norm().item(), norm)
        got = g.reshape(-1)[:4].double().cpu().numpy()
        assert np.abs(got - head).max() <= 1e-4 * max(np.abs(head).max(), norm / np.sqrt(g.numel())), name


def test_model_bf16_autocast_within_tolerance(golden_dir):
    G = np.load(os.path.join(golden_dir, "full_model.npz"))
    m, cfg, sd = build(64, 262)
    m.eval()
    x = O.synthetic_keypoints(2, 64, 2, seed=1001).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(x)
    err = rel_l2(logits.float(), torch.from_numpy(G["include_eval_logits"]))
    assert err < 2e-2, err


def test_model_bf16_train_grads_within_tolerance(golden_dir):
    G = np.load(os.path.join(golden_dir, "full_model.npz"))
    m, cfg, sd = build(64, 262, drop=0.0)
    m.train()
    x = O.synthetic_keypoints(2, 64, 2, seed=1001).cuda()
    y = O.synthetic_labels(2, 262, seed=1001).cuda()
    thr = [float(t) for t in G["include_train_thr"]]
    with patched_rand(thr), torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(x)
        loss = O.smoothed_cross_entropy(logits.float(), y)
    loss.backward()
    assert abs(loss.item() - float(G["include_train_loss"])) < 2e-2 * abs(float(G["include_train_loss"]))
    params = dict(m.named_parameters())
    # gradient norms of the attention parameters within the bf16 tolerance band (x a margin for
    # threshold decisions that flip under bf16 rounding of the logits)
    bad = []
    for name, norm in zip(G["include_train_gnames"], G["include_train_gnorms"]):
        g = params[str(name)].grad
        if abs(g.norm().item() - norm) / norm > 5e-2:
            bad.append((str(name), g.norm().item(), float(norm)))
    assert not bad, bad


def test_msa_reference_signature_equals_block_path():
    """MSA.forward(x, B, f, nW, mask) on the rolled + partitioned tensor == the fused block path."""
    from sl_hwgat_b200.models import HWGATE
    F, d, h = 8, 256, 4
    adj = torch.from_numpy(ADJ.astype(np.float32))
    torch.manual_seed(3)
    blk = HWGATE.PartAttentionBlock(dim=d, num_kps=64, num_heads=h, window_size=16, temporal_patch_size=2,
                                    temporal_dim=F, shift_size=1, adj_mat=torch.cat([adj] * (F // 2)).cuda(),
                                    drop=0.0, ff_ratio=2.).cuda().eval()
    x = torch.randn(2, F, 64, d, device="cuda")
    with torch.no_grad():
        xn = blk.norm1(x)
        fused = blk.attn.attend(xn, 1, blk._block_bits(x.device))
        xw = HWGATE.window_partition(torch.roll(xn, -1, 1), 16, 2).contiguous()
        yw = blk.attn(xw, 2, F // 2, 4, mask=blk.attn_mask)
        ref_style = torch.roll(HWGATE.window_reverse(yw, 16, 2, F, 64), 1, 1)
    assert rel_inf(ref_style, fused) < 1e-6


def test_threshold_draw_order_matches_reference():
    """one CPU-generator draw per MSA call, in block order (HWGATE.py:96): 8 per training forward."""
    m, cfg, sd = build(64, 262)
    m.train()
    x = O.synthetic_keypoints(1, 64, 2, seed=5).cuda()
    torch.manual_seed(1001)
    expect_after = None
    st = torch.get_rng_state()
    for _ in range(8):
        torch.rand(1)
    expect_after = torch.rand(1).item()
    torch.set_rng_state(st)
    m(x)
    assert torch.rand(1).item() == expect_after


# ------------------------------------------------------------------ K11: fused AdamW
@pytest.mark.parametrize("wd", [0.01, 0.0])
def test_fused_adamw_matches_torch(wd):
    """K11 against torch.optim.AdamW (the reference's optimizer, utils.py:73-75) on ragged tensor sizes, five steps,
    with a learning-rate change in between (CosineAnnealingLR drives param_groups[0]['lr'], utils.py:86-88)."""
    from sl_hwgat_b200.optim import AdamW
    g = torch.Generator().manual_seed(3)
    shapes = [(1,), (7,), (1000,), (4099,), (64, 128), (3, 5, 17), (4096 * 5 + 3,)]
    ours = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    ref = [p.detach().clone().requires_grad_(True) for p in ours]
    o1, o2 = AdamW(ours, lr=5e-4, weight_decay=wd), torch.optim.AdamW(ref, lr=5e-4, weight_decay=wd)
    for it in range(5):
        for p, q in zip(ours, ref):
            gr = torch.randn(p.shape, generator=g).cuda() * (10.0 ** (it - 2))
            p.grad, q.grad = gr.clone(), gr.clone()
        if it == 3:
            o1.param_groups[0]["lr"] = o2.param_groups[0]["lr"] = 1e-4
        o1.step(); o2.step()
    for p, q in zip(ours, ref):
        assert (p - q).abs().max().item() <= 2e-6 * max(1.0, q.abs().max().item())
    s1, s2 = o1.state_dict(), o2.state_dict()
    assert s1["param_groups"][0].keys() >= {"lr", "betas", "eps", "weight_decay"}
    for k in s2["state"]:
        assert set(s1["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(s1["state"][k]["step"]) == float(s2["state"][k]["step"]) == 5
        # (torch forms m with lerp, the kernel with b1*m + (1-b1)*g: same value, different rounding)
        for key in ("exp_avg", "exp_avg_sq"):
            a, b = s1["state"][k][key], s2["state"][k][key]
            assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item()
    # checkpoints interoperate: torch's state loads into ours and the next step still agrees
    import copy
    o1.load_state_dict(copy.deepcopy(s2))       # (a checkpoint round trip copies; load_state_dict itself aliases)
    for p, q in zip(ours, ref):
        p.data.copy_(q.data)
        gr = torch.randn(p.shape, generator=g).cuda()
        p.grad, q.grad = gr.clone(), gr.clone()
    o1.step(); o2.step()
    for p, q in zip(ours, ref):
        assert (p - q).abs().max().item() <= 2e-6 * max(1.0, q.abs().max().item())


def test_fused_adamw_grad_scale_and_skipped_params():
    from sl_hwgat_b200.optim import AdamW
    a = torch.ones(10, device="cuda", requires_grad=True)
    b = torch.ones(10, device="cuda", requires_grad=True)      # never gets a gradient
    r = torch.ones(10, device="cuda", requires_grad=True)
    o1, o2 = AdamW([a, b], lr=1e-2), torch.optim.AdamW([r], lr=1e-2)
    a.grad = torch.full((10,), 8.0, device="cuda"); r.grad = torch.full((10,), 2.0, device="cuda")
    o1.step(grad_scale=0.25); o2.step()
    assert torch.allclose(a, r, rtol=1e-6) and torch.equal(b, torch.ones_like(b))
    assert len(o1.state[b]) == 0


# ------------------------------------------------------------------ CUDA-graph inference (inference.py:95 path)
def test_graphed_inference_matches_eager_and_is_batch1_safe():
    from sl_hwgat_b200 import _lib
    from sl_hwgat_b200.runtime import GraphedInference
    T, classes = 16, 11
    model, _, _ = build(T, classes)
    model.train()
    with pytest.raises(RuntimeError):
        GraphedInference(model)                     # train mode: refused
    model.eval()
    fast = GraphedInference(model)
    for B in (1, 3, 1):                             # batch 1 is what inference.py:95 feeds; shapes are cached
        x = torch.rand(B, T, 64, 2, device="cuda")
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            want = model(x)
        got = fast(x)
        assert got.shape == (B, classes) and torch.equal(got, want)
    assert len(fast._graphs) == 2
    n0 = _lib.launch_count()
    fast(torch.rand(1, T, 64, 2, device="cuda"))
    assert _lib.launch_count() == n0                # replay: no host-side launches through the library


def test_graphed_inference_fp32_path():
    """the reference's per-sample evaluator runs without autocast (inference.py:95): the fp32 path (x3 GEMMs, whose
    plane scratch is stream-ordered and becomes allocation nodes of the graph) captured and replayed"""
    from sl_hwgat_b200.runtime import GraphedInference
    T, classes = 16, 11
    model, _, _ = build(T, classes)
    model.eval()
    fast = GraphedInference(model, autocast_dtype=None)
    for B in (1, 2, 1):
        x = torch.rand(B, T, 64, 2, device="cuda")
        with torch.no_grad():
            want = model(x)
        got = fast(x)
        assert got.dtype == torch.float32 and got.shape == (B, classes)
        assert torch.equal(got, want)
