"""The sibling model HGATE (hwgat/models/HGATE.py) on the general-window kernels: 29 keypoints
stored as 32, one 64-token window per block.  Parity against outputs of the unmodified reference
(tests/golden/hgate.npz) and the fp64 oracle (oracle/hgate_oracle.py, pinned to the same goldens on the CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import hgate_oracle as H
from oracle import hwgate_oracle as O
from tests._util import rel_inf, rel_l2

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def build(T=16, classes=10, drop=0.0):
    from sl_hwgat_b200.models import HGATE, model_params
    p = model_params.HGATEParams({"num_class": classes, "src_len": T}, 2, "cuda")
    p.drop_rate = drop
    torch.manual_seed(0)
    m = HGATE.Model(*p.get_model_params())
    cfg = H.HGATEConfig(temporal_dim=T, num_classes=classes)
    sd = H.make_state_dict(cfg, seed=1001, weight_std=0.05)
    m.load_state_dict(sd, strict=True)          # the reference's names and shapes
    return m.cuda(), cfg, sd, p


def test_hgate_params_and_state_dict_match_reference(golden_dir):
    G = np.load(os.path.join(golden_dir, "hgate.npz"))
    m, cfg, sd, p = build()
    assert p.adj_mat.device.type == "cpu" and tuple(p.adj_mat.shape) == (58, 58)
    assert np.array_equal(p.adj_mat.numpy(), G["adj"].astype(np.float32))           # reference get_adj_mat()
    assert np.array_equal(p.get_adj(), G["adj"][:29, :29].astype(np.float32))
    assert len(p.get_model_params()) == 15
    assert sorted(m.state_dict().keys()) == sorted(G["state_dict_names"])
    shapes = dict(zip(G["state_dict_names"], G["state_dict_shapes"]))
    assert all(str(tuple(v.shape)) == shapes[k] for k, v in m.state_dict().items())
    blk = m.layers[0].blocks[1]
    assert np.array_equal(blk.attn_mask.cpu().numpy() != 0, H.block_shift_mask(16, 29, 2, 1))


@pytest.mark.parametrize("d,h", [(128, 2), (256, 4)])
@pytest.mark.parametrize("shift", [0, 1])
def test_hgate_msa_reference_signature_vs_golden(golden_dir, d, h, shift):
    """MSA.forward(x, B, f, attn_mask) on the rolled + partitioned (B*f, 58, d) tensor, as the reference block calls
    it (HGATE.py:194-198), against the reference's fp64 outputs on the golden's seeded inputs."""
    from sl_hwgat_b200.models import HGATE
    G = np.load(os.path.join(golden_dir, "hgate.npz"))
    key = f"d{d}_s{shift}"
    B, F = 2, 4
    std = 0.2 if d == 128 else 0.1
    rng = np.random.default_rng(5000 + d + 10 * shift)
    xn = torch.from_numpy(rng.standard_normal((B, F, 29, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, 29, d)))
    adj = torch.from_numpy(G["adj"].astype(np.float32)).cuda()
    blk = HGATE.GraphAttentionBlock(dim=d, num_kps=29, num_heads=h, temporal_patch_size=2, temporal_dim=F,
                                    shift_size=shift, adj_mat=adj, drop=0.0).cuda()
    msa = blk.attn
    with torch.no_grad():
        msa.qkv.weight.copy_(w); msa.qkv.bias.copy_(b)
        msa.proj.weight.copy_(torch.eye(d)); msa.proj.bias.zero_()
    x_ = xn.float().cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        xs = torch.roll(x_, shifts=-shift, dims=1) if shift else x_
        yb = msa(HGATE.block_partition(xs, 2), B, F // 2, attn_mask=blk.attn_mask)
        y = HGATE.block_reverse(yb, 2, F, 29)
        y = torch.roll(y, shifts=shift, dims=1) if shift else y
    (y.float() * g.float().cuda()).sum().backward()

    def chk(t, name, stride):
        a = t.detach().double().cpu().reshape(-1).numpy()[::stride]
        ref = G[key + "_" + name]
        return float(np.linalg.norm(a - ref) / np.linalg.norm(ref))
    # the identity projection runs as a bf16 GEMM here, so y carries one more bf16 rounding than the core
    assert chk(y, "y", 53) < BF16_TOL and chk(x_.grad, "dx", 53) < 3e-2
    assert chk(msa.qkv.weight.grad, "dw", 251) < 3e-2


def test_hgate_model_vs_reference_golden_and_oracle_gradients(golden_dir):
    G = np.load(os.path.join(golden_dir, "hgate.npz"))
    m, cfg, sd, p = build()
    m.train()                                    # drop 0: HGATE has no threshold path, train == eval numerics
    x = H.synthetic_keypoints(2, 16, seed=1001).cuda()
    y = torch.tensor([3, 7]).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = m(x)
        loss = O.smoothed_cross_entropy(logits.float(), y)
    loss.backward()
    assert rel_l2(logits.float(), torch.from_numpy(G["model_logits"])) < BF16_TOL
    assert abs(loss.item() - float(G["model_loss"])) < 1e-2 * abs(float(G["model_loss"]))
    sd64 = {k: v.cuda().double().requires_grad_(k not in ("B", "pos_encoder.pe") and not k.endswith("attn_mask"))
            for k, v in sd.items()}
    ref = H.model_forward(x.double(), sd64, cfg)
    O.smoothed_cross_entropy(ref, y).backward()
    errs = {n: rel_l2(q.grad, sd64[n].grad) for n, q in m.named_parameters() if q.grad is not None}
    assert set(errs) == {k for k, v in sd64.items() if v.grad is not None}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print("HGATE worst gradient rel_l2:", worst)
    assert worst[0][1] < 5e-2, worst
    # eval forward, batch 1 (inference.py:95), T = 64
    m2, cfg2, sd2, _ = build(T=64, classes=262)
    m2.eval()
    x1 = H.synthetic_keypoints(1, 64, seed=5).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out = m2(x1).float()
    ref1 = H.model_forward(x1.cpu().double(), {k: v.double() for k, v in sd2.items()}, cfg2)
    assert rel_l2(out, ref1) < BF16_TOL


@pytest.mark.parametrize("d,h", [(128, 2), (256, 4)])
@pytest.mark.parametrize("shift", [0, 1])
def test_hgate_fp32_msa_vs_reference_golden(golden_dir, d, h, shift):
    """MSA.forward without autocast: the true-fp32 general-window kernels (attn_win_f32.cu) on HGATE's 58-token blocks
    stored as 64, against the reference's fp64 outputs: 1e-5"""
    from sl_hwgat_b200.models import HGATE
    G = np.load(os.path.join(golden_dir, "hgate.npz"))
    key = f"d{d}_s{shift}"
    B, F = 2, 4
    std = 0.2 if d == 128 else 0.1
    rng = np.random.default_rng(5000 + d + 10 * shift)
    xn = torch.from_numpy(rng.standard_normal((B, F, 29, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, 29, d)))
    adj = torch.from_numpy(G["adj"].astype(np.float32)).cuda()
    blk = HGATE.GraphAttentionBlock(dim=d, num_kps=29, num_heads=h, temporal_patch_size=2, temporal_dim=F,
                                    shift_size=shift, adj_mat=adj, drop=0.0).cuda()
    msa = blk.attn
    with torch.no_grad():
        msa.qkv.weight.copy_(w); msa.qkv.bias.copy_(b)
        msa.proj.weight.copy_(torch.eye(d)); msa.proj.bias.zero_()
    x_ = xn.float().cuda().requires_grad_(True)
    xs = torch.roll(x_, shifts=-shift, dims=1) if shift else x_
    yb = msa(HGATE.block_partition(xs, 2), B, F // 2, attn_mask=blk.attn_mask)
    y = HGATE.block_reverse(yb, 2, F, 29)
    y = torch.roll(y, shifts=shift, dims=1) if shift else y
    assert y.dtype == torch.float32
    (y * g.float().cuda()).sum().backward()

    def chk(t, name, stride):
        a = t.detach().double().cpu().reshape(-1).numpy()[::stride]
        ref = G[key + "_" + name]
        return float(np.abs(a - ref).max() / np.abs(ref).max())
    errs = (chk(y, "y", 53), chk(x_.grad, "dx", 53), chk(msa.qkv.weight.grad, "dw", 251), chk(msa.qkv.bias.grad, "db", 1))
    print(key, "HGATE fp32 max-rel y/dx/dw/db:", errs)
    assert max(errs) < 1e-5, errs


def test_hgate_fp32_model_vs_reference_golden(golden_dir):
    """the drop-in model called without autocast (utils.py:102): logits, loss and every gradient against the
    reference's fp64 run"""
    G = np.load(os.path.join(golden_dir, "hgate.npz"))
    m, cfg, sd, p = build()
    m.train()                                    # drop 0
    x = H.synthetic_keypoints(2, 16, seed=1001).cuda()
    y = torch.tensor([3, 7]).cuda()
    logits = m(x)
    assert logits.dtype == torch.float32
    loss = O.smoothed_cross_entropy(logits, y)
    loss.backward()
    assert rel_inf(logits, torch.from_numpy(G["model_logits"])) < 1e-5
    assert abs(loss.item() - float(G["model_loss"])) < 1e-5 * abs(float(G["model_loss"]))
    grads = dict(m.named_parameters())
    for name, norm, head in zip(G["gnames"], G["gnorms"], G["gheads"]):
        gr = grads[str(name)].grad.double().cpu()
        assert abs(gr.norm().item() - norm) <= 1e-4 * norm, name
        assert np.abs(gr.reshape(-1)[:4].numpy() - head).max() <= 1e-4 * max(np.abs(head).max(), 1e-12) + 1e-9, name


def test_hgate_cpu_is_refused():
    from sl_hwgat_b200 import _lib
    m, cfg, sd, p = build()
    m.eval()
    with pytest.raises(_lib.HwgatError):
        m.cpu()(H.synthetic_keypoints(2, 16, seed=1))
