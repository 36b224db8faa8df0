"""Round-2 parity tests (VERDICT r01 "what's weak" 1-4, "what's missing" 6):

 * the persistent MULTI-TILE path of K2/K3 (every CTA walks several tiles: barrier phases, accumulator
   parity and weight-ring wrap across tiles) at d=256 and d=512, B=512 and an odd tile count, against a
   small case that is itself checked against the oracle;
 * bf16 TRAIN mode pinned to the unmodified reference under torch.autocast (tests/golden/autocast_train.npz);
 * every parameter's FULL gradient (not its norm) against oracle autograd;
 * run-to-run determinism of outputs and input gradients, bounded non-determinism of the split-K weight gradients;
 * the way the reference instantiates and restores the model (utils.py:55-59, 185-214), replayed.
"""
import contextlib
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import hwgate_oracle as O
from tests._util import ADJ, core_inputs, cuda_core, device_bits, oracle_core, rel_inf, rel_l2

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BF16_TOL = 2e-2


@contextlib.contextmanager
def patched_rand(values):
    it = iter(values)
    real = torch.rand
    torch.rand = lambda *a, **k: torch.tensor([next(it)])
    try:
        yield
    finally:
        torch.rand = real


# ------------------------------------------------------------------ multi-tile K2 / K3 at every level
def _bf16_oracle_points(xn, w, b, g, h, F, shift, thr):
    mask = O.combined_mask(ADJ, F, 16, 2, shift)
    x_, w_, b_ = (t.clone().requires_grad_(True) for t in (xn, w, b))
    y = O.attention_core(x_, w_, b_, h, mask, 16, 2, shift, thr, bf16_points=True)
    (y * g).sum().backward()
    return y.detach(), x_.grad, w_.grad, b_.grad


@pytest.mark.parametrize("d,h,F", [(128, 2, 64), (256, 4, 32), (512, 8, 16)])
@pytest.mark.parametrize("shift", [0, 1])
@pytest.mark.parametrize("thr", [None, 0.05])
def test_attention_multi_tile_replication(d, h, F, shift, thr):
    """BASELINE configs[2] sizes (B=512 at each level's frame count) and an odd batch (B=37: 37*F/2 tiles do not
    divide over 148 CTAs): every sample is the same sequence, so each sample's output / input gradient must be
    BIT-identical to the one-sample case, which is checked against the oracle here, and dW / db must be B times it."""
    xn, w, b, g = core_inputs(d, shift, 0.05, B=1, F=F)
    xn, g = xn.to(torch.bfloat16).double(), g.to(torch.bfloat16).double()
    w = w.to(torch.bfloat16).double()
    b = b.float().double()
    ys, dxs, dws, dbs = cuda_core(xn, w, b, g, h, shift, thr, torch.bfloat16)
    # (1) the one-sample case against the oracle (F/2 tiles: one tile per CTA)
    by, bdx, bdw, bdb = _bf16_oracle_points(xn, w, b, g, h, F, shift, thr)
    errs = dict(y=rel_l2(ys, by), dx=rel_l2(dxs, bdx), dw=rel_l2(dws, bdw), db=rel_l2(dbs, bdb))
    assert all(e < BF16_TOL for e in errs.values()), errs
    # (2) many tiles per CTA
    for Bbig in (512, 37):
        yb, dxb, dwb, dbb = cuda_core(xn.expand(Bbig, -1, -1, -1).contiguous(), w, b,
                                      g.expand(Bbig, -1, -1, -1).contiguous(), h, shift, thr, torch.bfloat16)
        assert torch.equal(yb, ys.expand(Bbig, -1, -1, -1)), (d, Bbig)
        assert torch.equal(dxb, dxs.expand(Bbig, -1, -1, -1)), (d, Bbig)
        assert rel_l2(dwb, dws * Bbig) < 1e-3 and rel_l2(dbb, dbs * Bbig) < 1e-3
        del yb, dxb, dwb, dbb
        torch.cuda.empty_cache()


@pytest.mark.parametrize("d,h,F", [(256, 4, 8), (512, 8, 8)])
def test_attention_multi_tile_distinct_samples_vs_oracle(d, h, F):
    """Distinct samples, more tiles (B*F/2 = 304 and 608) than CTAs (148), every one checked against the oracle:
    a tile that picked up a neighbour's X chunk, weight stage or accumulator would show here."""
    B = 76 if d == 256 else 152
    xn, w, b, g = core_inputs(d, 1, 0.05, B=B, F=F)
    xn, g = xn.to(torch.bfloat16).double(), g.to(torch.bfloat16).double()
    w = w.to(torch.bfloat16).double()
    b = b.float().double()
    for thr in (None, 0.04):
        y, dx, dw, db = cuda_core(xn, w, b, g, h, 1, thr, torch.bfloat16)
        by, bdx, bdw, bdb = _bf16_oracle_points(xn, w, b, g, h, F, 1, thr)
        # per-sample errors: a single bad tile must not hide in the aggregate
        ey = ((y.double().cpu() - by).flatten(1).norm(dim=1) / by.flatten(1).norm(dim=1)).max().item()
        ex = ((dx.double().cpu() - bdx).flatten(1).norm(dim=1) / bdx.flatten(1).norm(dim=1)).max().item()
        # eval: continuous, every sample within the tolerance.  train: the threshold decisions of a few borderline
        # logits differ between the kernel's fp32 accumulation order and the oracle's (HWGATE.py:94-100 is
        # discontinuous), which moves whole rows of a sample: the aggregate stays within the tolerance, one sample
        # may not, so its bound is looser - a tile that read a neighbour's data would be O(1) off.
        per_sample = BF16_TOL if thr is None else 8e-2
        assert ey < per_sample and ex < per_sample, (thr, ey, ex)
        assert rel_l2(y, by) < BF16_TOL and rel_l2(dx, bdx) < BF16_TOL, (thr, rel_l2(y, by), rel_l2(dx, bdx))
        assert rel_l2(dw, bdw) < BF16_TOL and rel_l2(db, bdb) < BF16_TOL


# ------------------------------------------------------------------ full model: builders
def build(T, classes, drop=0.0, std=0.05):
    from sl_hwgat_b200.models import HWGATE, model_params
    p = model_params.HWGATEParams({"num_class": classes, "src_len": T}, 2, "cuda")
    p.drop_rate = drop
    torch.manual_seed(0)
    m = HWGATE.Model(*p.get_model_params())
    cfg = O.HWGATEConfig(temporal_dim=T, num_classes=classes)
    sd = O.make_state_dict(cfg, seed=1001, weight_std=std)
    m.load_state_dict(sd, strict=True)
    return m.cuda(), cfg, sd


def _oracle_grads(sd, cfg, x, y, thr, device="cuda"):
    """fp64 oracle forward + autograd of every trainable parameter (run on the box's GPU in fp64: the oracle is
    device-agnostic torch code; it stays the checker)."""
    frozen = ("B", "pos_encoder.pe")
    sd64 = {k: v.to(device).double().requires_grad_(k not in frozen and not k.endswith("attn_mask"))
            for k, v in sd.items()}
    logits = O.model_forward(x.to(device).double(), sd64, cfg, thresholds=thr)
    loss = O.smoothed_cross_entropy(logits, y.to(device))
    loss.backward()
    return logits.detach(), loss.item(), {k: v.grad for k, v in sd64.items() if v.grad is not None}


THR = [0.03, 0.05, 0.031, 0.2, 0.033, 0.04, 0.0312, 0.1]


def _run_model(m, x, y, thr, autocast):
    m.zero_grad(set_to_none=True)
    with patched_rand(thr), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits = m(x)
        loss = O.smoothed_cross_entropy(logits.float(), y)
    loss.backward()
    return logits.detach().float(), loss.item(), {n: p.grad for n, p in m.named_parameters() if p.grad is not None}


def test_model_fp32_train_every_gradient_vs_oracle():
    """fp32 path, train mode: rel_l2 of EVERY parameter's full gradient against oracle autograd."""
    m, cfg, sd = build(64, 262, drop=0.0)
    m.train()
    x = O.synthetic_keypoints(2, 64, 2, seed=1001).cuda()
    y = O.synthetic_labels(2, 262, seed=1001).cuda()
    logits, loss, grads = _run_model(m, x, y, THR, autocast=False)
    r_logits, r_loss, r_grads = _oracle_grads(sd, cfg, x, y, THR)
    assert rel_inf(logits, r_logits) < 1e-5
    assert set(grads) == set(r_grads)
    errs = {n: rel_l2(grads[n], r_grads[n]) for n in grads}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print("fp32 worst gradient rel_l2:", worst)
    assert worst[0][1] < 1e-4, worst


def test_model_bf16_train_every_gradient_vs_oracle():
    """bf16 autocast path, train mode: every parameter's full gradient against fp64 oracle autograd.  Bound: 5e-2
    per parameter in rel_l2 (a gradient is a sum over 8192 tokens x 8 blocks of bf16-rounded products, and
    threshold decisions that flip under bf16 rounding of the logits move whole rows, HWGATE.py:94-100; the
    reference's own autocast run is 2e-2 from its fp64 run on the LOGITS of this case, see the golden test below);
    direction is checked too: cosine > 0.998."""
    m, cfg, sd = build(64, 262, drop=0.0)
    m.train()
    x = O.synthetic_keypoints(2, 64, 2, seed=1001).cuda()
    y = O.synthetic_labels(2, 262, seed=1001).cuda()
    logits, loss, grads = _run_model(m, x, y, THR, autocast=True)
    r_logits, r_loss, r_grads = _oracle_grads(sd, cfg, x, y, THR)
    assert rel_l2(logits, r_logits) < 2.5e-2
    assert abs(loss - r_loss) < 2e-2 * abs(r_loss)
    assert set(grads) == set(r_grads)
    errs = {n: rel_l2(grads[n], r_grads[n]) for n in grads}
    cos = {n: float(torch.nn.functional.cosine_similarity(grads[n].double().flatten(), r_grads[n].flatten(), dim=0))
           for n in grads}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print("bf16 worst gradient rel_l2:", worst, "min cosine:", min(cos.values()))
    assert worst[0][1] < 5e-2, worst
    assert min(cos.values()) > 0.998, sorted(cos.items(), key=lambda kv: kv[1])[:5]


def test_model_bf16_train_vs_reference_autocast_golden(golden_dir):
    """The unmodified reference under torch.autocast(bfloat16), train mode, injected thresholds
    (tests/golden/autocast_train.npz, made by make_golden.py section 5).  Two independent bf16 evaluations of a
    discontinuous function: the reference's own autocast logits are 1.95e-2 (train) / 0.99e-2 (eval) from its fp64
    logits, so the assertions are (a) this path is no further from the fp64 reference than the reference's autocast
    run is (x1.25), and (b) the two bf16 runs are within the sum of their distances to fp64."""
    G = np.load(os.path.join(golden_dir, "autocast_train.npz"))
    F64 = np.load(os.path.join(golden_dir, "full_model.npz"))
    thr = [float(t) for t in G["thr"]]
    assert thr == [float(t) for t in F64["include_train_thr"]]
    m, cfg, sd = build(64, 262, drop=0.0)
    m.train()
    x = O.synthetic_keypoints(2, 64, 2, seed=1001).cuda()
    y = O.synthetic_labels(2, 262, seed=1001).cuda()
    logits, loss, grads = _run_model(m, x, y, thr, autocast=True)
    ref_ac, ref_64 = torch.from_numpy(G["logits"]), torch.from_numpy(F64["include_train_logits"])
    gap_ref = rel_l2(ref_ac, ref_64)                 # the reference's own bf16 gap (0.0195)
    gap_ours = rel_l2(logits, ref_64)
    gap_between = rel_l2(logits, ref_ac)
    print(f"train logits: reference autocast vs fp64 {gap_ref:.3e}; ours vs fp64 {gap_ours:.3e}; "
          f"ours vs reference autocast {gap_between:.3e}")
    assert gap_ours < max(BF16_TOL, 1.25 * gap_ref)
    assert gap_between < gap_ref + gap_ours + 5e-3
    assert abs(loss - float(G["loss"])) < 1e-2 * abs(float(G["loss"]))
    # gradients: strided samples of every parameter's gradient as the reference's autocast run produced them, next to
    # the same samples of the fp64 oracle gradients.  The reference's OWN autocast run is up to 11.5 % (LayerNorm
    # weights) and ~5 % (fc1 / proj / qkv weights of level 0) away from its fp64 gradients on these samples, so the
    # statement checked per parameter is: this path is at least as close to the fp64 reference as the reference's
    # autocast run is (x1.25 + 1e-2), and never further than 5e-2.
    r_logits, r_loss, r_grads = _oracle_grads(sd, cfg, x, y, thr)
    stride, offs = int(G["stride"]), G["goffsets"]
    bad, worst_ref, worst_ours = [], 0.0, 0.0
    for i, name in enumerate(G["gnames"]):
        g = grads[str(name)].detach().float().reshape(-1).cpu().numpy()[::stride]
        ac = G["gsamples"][offs[i]:offs[i + 1]]
        r64 = r_grads[str(name)].detach().reshape(-1).cpu().numpy()[::stride]
        assert g.shape == ac.shape == r64.shape, name
        nrm = max(np.linalg.norm(r64), 1e-30)
        gap_ref_i, gap_ours_i = np.linalg.norm(ac - r64) / nrm, np.linalg.norm(g - r64) / nrm
        worst_ref, worst_ours = max(worst_ref, gap_ref_i), max(worst_ours, gap_ours_i)
        nerr = abs(np.linalg.norm(grads[str(name)].double().cpu().numpy()) - G["gnorms"][i]) / G["gnorms"][i]
        if gap_ours_i > min(5e-2, 1.25 * gap_ref_i + 1e-2) or nerr > 3e-2:
            bad.append((str(name), float(gap_ours_i), float(gap_ref_i), float(nerr)))
    print(f"gradient samples vs fp64: worst reference-autocast {worst_ref:.3e}, worst ours {worst_ours:.3e}")
    assert not bad, bad
    # eval mode (no threshold): continuous, so the plain 2e-2 applies between the two bf16 runs
    m.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ev = m(x).float()
    assert rel_l2(ev, torch.from_numpy(G["eval_logits"])) < BF16_TOL
    assert rel_l2(ev, torch.from_numpy(F64["include_eval_logits"])) < BF16_TOL


# ------------------------------------------------------------------ determinism
def test_run_twice_determinism():
    """Outputs and input gradients are bit-reproducible (every tile is computed by one CTA in a fixed order).
    Weight / bias gradients are split over tokens and combined with fp32 atomic adds (gemm_tc.cu, block_fused.cu),
    so their summation ORDER varies from run to run: bounded here at 1e-5 relative (fp32 rounding of a sum of a
    few hundred partials), documented in DESIGN.md section 2.  The reference (cuBLAS/ATen) is deterministic."""
    d, h, F, B = 256, 4, 32, 64
    xn, w, b, g = core_inputs(d, 1, 0.05, B=1, F=F)
    xn = xn.expand(B, -1, -1, -1).contiguous() + 0.01 * torch.arange(B).double().reshape(B, 1, 1, 1)
    g = g.expand(B, -1, -1, -1).contiguous()
    runs = [cuda_core(xn, w, b, g, h, 1, 0.05, torch.bfloat16) for _ in range(3)]
    for r in runs[1:]:
        assert torch.equal(r[0], runs[0][0]) and torch.equal(r[1], runs[0][1])
        assert rel_l2(r[2], runs[0][2]) < 1e-5 and rel_l2(r[3], runs[0][3]) < 1e-5
    # whole model, bf16 train step with dropout: same seed -> same masks -> logits bit-identical
    m, cfg, sd = build(16, 10, drop=0.1)
    m.train()
    x = O.synthetic_keypoints(4, 16, 2, seed=3).cuda()
    outs = []
    for _ in range(2):
        torch.manual_seed(5)
        torch.cuda.manual_seed(5)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs.append(m(x).detach().clone())
    assert torch.equal(outs[0], outs[1])


def test_deterministic_mode_makes_parameter_gradients_bit_reproducible():
    """ops.set_deterministic(True): unsplit weight-gradient GEMMs + fixed-order column sums -> every gradient of a
    whole train step (bf16 autocast, dropout on, threshold drop on) is bit-identical between two runs; and the
    deterministic gradients agree with the default (atomic) ones to fp32 summation-order noise."""
    import sl_hwgat_b200.ops as ops
    from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss

    def step(m, x, y):
        torch.manual_seed(5)
        torch.cuda.manual_seed(5)
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = SmoothedCrossEntropyLoss(0.1)(m(x), y)
        loss.backward()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}

    m, cfg, sd = build(32, 10, drop=0.1)
    m.train()
    x = O.synthetic_keypoints(64, 32, 2, seed=3).cuda()
    y = torch.arange(64, device="cuda") % 10
    base = step(m, x, y)
    prev = ops.set_deterministic(True)
    try:
        a, b = step(m, x, y), step(m, x, y)
    finally:
        ops.set_deterministic(prev)
    assert a.keys() == b.keys() == base.keys() and len(a) >= 100
    for n in a:
        assert torch.equal(a[n], b[n]), n
        assert rel_l2(a[n].double(), base[n].double()) < 1e-4, n
    # the fp32 path too
    m32, _, _ = build(16, 10, drop=0.0)
    m32.train()
    x32 = O.synthetic_keypoints(8, 16, 2, seed=4).cuda()

    def step32():
        torch.manual_seed(7)
        m32.zero_grad(set_to_none=True)
        SmoothedCrossEntropyLoss(0.1)(m32(x32), y[:8]).backward()
        return {n: p.grad.detach().clone() for n, p in m32.named_parameters() if p.grad is not None}

    prev = ops.set_deterministic(True)
    try:
        a, b = step32(), step32()
    finally:
        ops.set_deterministic(prev)
    for n in a:
        assert torch.equal(a[n], b[n]), n


# ------------------------------------------------------------------ drop-in replay (utils.py:55-59, 185-214)
def test_dropin_replays_reference_load_model_and_checkpoint_filter(tmp_path):
    """Instantiate the model the way utils.load_model does - importlib on 'models.<model_type>' and
    'models.model_params' with the package directory standing in for hwgat/ - and restore a checkpoint the way
    utils.load_weights_from_pretrained does (strip 'model.', keep only same-name same-shape entries)."""
    pkg = os.path.join(ROOT, "sl_hwgat_b200")
    sys.path.insert(0, pkg)
    try:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        model_type, device = "HWGATE", torch.device("cuda:0")
        params_mod = importlib.import_module("models.model_params")                        # configs.py:80
        model_params = getattr(params_mod, model_type + "Params")({"num_class": 50, "src_len": 32}, 2, device)
        module = importlib.import_module("models." + model_type)                           # utils.py:56
        model = getattr(module, "Model")(*model_params.get_model_params())                  # utils.py:57
        model.to(device)                                                                   # utils.py:58
        assert os.path.isfile(os.path.join(pkg, "models", model_type + ".py"))              # utils.py:181 copies it
        assert os.path.isfile(os.path.join(pkg, "models", "model_params.py"))              # utils.py:182

        # a checkpoint written by a run with another class count: head.* must be skipped, the rest restored
        cfg_old = O.HWGATEConfig(temporal_dim=32, num_classes=77)
        sd_old = O.make_state_dict(cfg_old, seed=9, weight_std=0.05)
        path = str(tmp_path / "ckpt.pt")
        torch.save({"model_state_dict": {"model." + k: v for k, v in sd_old.items()} | {"model.extra.w": torch.ones(3)},
                    "epoch": 3}, path)
        ckpt = torch.load(path, map_location=device)["model_state_dict"]                   # utils.py:186
        pretrained = {k.replace("model.", ""): v for k, v in ckpt.items()}                 # utils.py:188
        model_dict = model.state_dict()
        before_head = model_dict["head.weight"].clone()
        tmp, skipped = {}, []
        for k, v in pretrained.items():                                                    # utils.py:194-202
            if k in model_dict:
                if v.shape == model_dict[k].shape:
                    tmp[k] = v
                else:
                    tmp[k] = model_dict[k]
                    skipped.append(k)
            else:
                skipped.append(k)
        assert sorted(skipped) == ["extra.w", "head.bias", "head.weight"]
        assert [k for k in model_dict if k not in pretrained] == []                        # utils.py:206-208
        model_dict.update(tmp)
        model.load_state_dict(model_dict)                                                  # utils.py:213
        model.to(dtype=torch.float)                                                        # utils.py:214
        assert torch.equal(model.state_dict()["head.weight"], before_head)
        assert torch.equal(model.state_dict()["layers.1.blocks.1.attn.qkv.weight"].cpu(),
                           sd_old["layers.1.blocks.1.attn.qkv.weight"])
        # and it runs: the features (everything but the skipped head) match the oracle with the checkpoint's weights
        model.eval()
        x = O.synthetic_keypoints(2, 32, 2, seed=4).to(device)
        with torch.no_grad():
            logits = model(x)                                                              # utils.py:128
            feats = model.forward_features(x)
        assert logits.shape == (2, 50)
        sd_id = dict(sd_old)
        sd_id["head.weight"], sd_id["head.bias"] = torch.eye(512), torch.zeros(512)
        cfg_id = O.HWGATEConfig(temporal_dim=32, num_classes=512)
        ref = O.model_forward(x.cpu().double(), {k: v.double() for k, v in sd_id.items()}, cfg_id)
        assert rel_inf(feats, ref) < 1e-5
    finally:
        sys.path.remove(pkg)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]


# ------------------------------------------------------------------ mask caches (VERDICT weak 15, ADVICE)
def test_adjacency_replaced_after_first_forward_is_honoured():
    from sl_hwgat_b200.models import HWGATE
    F, d, h = 8, 128, 2
    adj = torch.from_numpy(ADJ.astype(np.float32))
    torch.manual_seed(3)
    blk = HWGATE.PartAttentionBlock(dim=d, num_kps=64, num_heads=h, window_size=16, temporal_patch_size=2,
                                    temporal_dim=F, shift_size=1, adj_mat=torch.cat([adj] * (F // 2)).cuda(),
                                    drop=0.0, ff_ratio=2.).cuda().eval()
    x = torch.randn(2, F, 64, d, device="cuda")
    with torch.no_grad():
        y0 = blk(x)
        # (a) replaced by another tensor: a chain graph
        chain = [[[i, i + 1] for i in range(15)]] * 4
        adj2 = torch.from_numpy(O.window_adjacency(chain, 16, 2).astype(np.float32))
        blk.attn.adj_mat = torch.cat([adj2] * (F // 2)).cuda()
        y1 = blk(x)
        xn = blk.norm1(x)
        bits2 = device_bits_for(adj2, F, 1)
        want = x + blk.attn.attend(xn, 1, bits2)
        want = want + blk.ff(blk.norm2(want))
        assert not torch.equal(y0, y1) and rel_inf(y1, want) < 1e-6
        # (b) modified in place
        blk.attn.adj_mat.fill_(1.0)
        y2 = blk(x)
        assert not torch.equal(y1, y2)
        # (c) a per-temporal-group adjacency (not a replication of the first nW windows) goes through K1c
        per_group = torch.cat([adj, adj2, adj, adj2]).cuda()
        blk.attn.adj_mat = per_group
        y3 = blk(x)
        xw = HWGATE.window_partition(torch.roll(xn, -1, 1), 16, 2).contiguous()
        yw = blk.attn(xw, 2, F // 2, 4, mask=blk.attn_mask)
        want = x + torch.roll(HWGATE.window_reverse(yw, 16, 2, F, 64), 1, 1)
        want = want + blk.ff(blk.norm2(want))
        assert rel_inf(y3, want) < 1e-6


def device_bits_for(adj, F, shift):
    from sl_hwgat_b200 import ops
    return ops.mask_build(adj.cuda(), F, shift)


def test_non_binary_masks_are_refused():
    """The reference multiplies the logits by the float adj_mat / mask (HWGATE.py:102-108); a weighted mask cannot
    be represented by the packed bits and must raise instead of being binarised."""
    from sl_hwgat_b200 import ops
    adj = torch.from_numpy(ADJ.astype(np.float32)).cuda()
    with pytest.raises(NotImplementedError):
        ops.mask_build(adj * 0.5, 8, 0)
    with pytest.raises(NotImplementedError):
        ops.mask_pack(adj, torch.full((16, 32, 32), 2.0, device="cuda"), 16, 32, "cuda")
    ops.mask_pack(adj, torch.ones(16, 32, 32, device="cuda"), 16, 32, "cuda")


def test_input_gradient_under_autocast():
    """x.requires_grad under bf16 autocast (saliency / adversarial probes): the fused embedding kernel is
    forward-only, so the model takes the differentiable PyTorch embedding instead of raising."""
    m, cfg, sd = build(16, 10)
    m.eval()
    x = O.synthetic_keypoints(2, 16, 2, seed=3).cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x)
    out.float().sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_second_device_in_one_process():
    """The >48 KB dynamic shared-memory opt-in is a per-device attribute: forward + backward on cuda:1 after cuda:0."""
    from sl_hwgat_b200 import ops
    d, h, F = 512, 8, 8
    xn, w, b, g = core_inputs(d, 0, 0.05, B=2, F=F)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        adj = torch.from_numpy(ADJ.astype(np.float32)).to(dev)
        bits = ops.mask_build(adj, F, 0)
        x_ = xn.to(dev, torch.bfloat16).requires_grad_(True)
        w_ = w.float().to(dev).requires_grad_(True)
        y = ops.window_graph_attention(x_, w_, b.float().to(dev), bits, h)
        y.backward(g.to(dev, torch.bfloat16))
        hdn = torch.randn(256, d, device=dev, dtype=torch.bfloat16)
        ops.feed_forward_core(hdn, torch.randn(2 * d, d, device=dev) * 0.02, torch.zeros(2 * d, device=dev),
                              torch.randn(d, 2 * d, device=dev) * 0.02, 0.0, False)
        torch.cuda.synchronize(dev)
        outs.append((y.detach().cpu(), x_.grad.cpu(), w_.grad.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert rel_l2(outs[0][2], outs[1][2]) < 1e-5


# ------------------------------------------------------------------ K4 folded into K6 / K5'
@pytest.mark.parametrize("d,F,B", [(128, 8, 3), (256, 4, 2), (128, 64, 5)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_merge_fold_equals_k6_k4_k5(d, F, B, p):
    """bias_dropout_add_merge_ln (K6 storing TemporalMerging's layout + K5 over the merged rows; backward K5' storing
    un-merged + K6') against the three separate ops K6 -> K4 -> K5: same dropout stream, so forward is bit-identical
    and the gradients agree to fp32 rounding of the atomically summed column reductions."""
    from sl_hwgat_b200 import ops
    g = torch.Generator().manual_seed(d + F)
    res = torch.randn(B, F, 64, d, generator=g).cuda()
    a0 = torch.randn(B, F, 64, d, generator=g).to(torch.bfloat16).cuda()
    bias = torch.randn(d, generator=g).cuda()
    norm = torch.nn.LayerNorm(2 * d).cuda()
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(2 * d, generator=g))
        norm.bias.copy_(0.1 * torch.randn(2 * d, generator=g))
    gx = torch.randn(B, F // 2, 64, 2 * d, generator=g).cuda()
    gy = torch.randn(B, F // 2, 64, 2 * d, generator=g).to(torch.bfloat16).cuda()

    def run(folded):
        r, a, b_ = res.clone().requires_grad_(True), a0.clone().requires_grad_(True), bias.clone().requires_grad_(True)
        norm.zero_grad()
        torch.manual_seed(11); torch.cuda.manual_seed(11)
        if folded:
            xm, y = ops.bias_dropout_add_merge_ln(r, a, b_, norm, p, True)
        else:
            x1, _ = ops.bias_dropout_add_ln(r, a, b_, None, p, True)
            xm = ops.temporal_merge(x1)
            xm, y = ops.layer_norm_residual(xm, norm.weight, norm.bias, norm.eps)
        torch.autograd.backward([xm, y], [gx, gy])
        return xm.detach(), y.detach(), r.grad, a.grad, b_.grad, norm.weight.grad.clone(), norm.bias.grad.clone()

    f, u = run(True), run(False)
    assert torch.equal(f[0], u[0]) and torch.equal(f[1], u[1])
    assert torch.equal(f[2], u[2]) and torch.equal(f[3], u[3])          # d_res, d_a0: row-local, bit-identical
    for i in (4, 5, 6):
        assert rel_l2(f[i], u[i]) < 1e-5, i
    # and against PyTorch in fp64 when there is no dropout
    if p == 0.0:
        x1 = res.double() + a0.double() + bias.double()
        xm_ref = O.temporal_merge(x1.cpu(), 2)
        assert rel_inf(f[0], xm_ref) < 1e-6
        y_ref = torch.nn.functional.layer_norm(xm_ref, (2 * d,), norm.weight.double().cpu(), norm.bias.double().cpu(), 1e-5)
        assert rel_l2(f[1], y_ref) < 4e-3


def test_model_uses_merge_fold_and_matches_unfolded():
    """The full model takes the folded path under autocast (no K4 launches) and gives the same logits / gradients as
    the per-layer path that still launches K4."""
    from sl_hwgat_b200 import _lib
    m, cfg, sd = build(16, 10, drop=0.0)
    m.train()
    x = O.synthetic_keypoints(2, 16, 2, seed=3).cuda()
    y = O.synthetic_labels(2, 10, seed=3).cuda()
    thr = THR
    calls = {"merge": 0}
    lib = _lib.load()
    real = lib.hwgat_merge_fwd

    def counting(*a):
        calls["merge"] += 1
        return real(*a)
    lib.hwgat_merge_fwd = counting
    try:
        logits, loss, grads = _run_model(m, x, y, thr, autocast=True)
        assert calls["merge"] == 0
        # per-layer path: call the layers one by one (each PartAttentionLayer.forward ends in its own K4)
        m.zero_grad(set_to_none=True)
        with patched_rand(thr), torch.autocast("cuda", dtype=torch.bfloat16):
            from sl_hwgat_b200 import ops
            h = ops.fourier_embed(x, m.B, m.pos_encoder.pe, 0.0, True)
            for layer in m.layers:
                h = layer(h)
            feats = ops.layer_norm_mean_pool(h, m.norm.weight, m.norm.bias, m.norm.eps)
            logits2 = ops.linear_f32(feats, m.head.weight, m.head.bias)
            loss2 = O.smoothed_cross_entropy(logits2.float(), y)
        loss2.backward()
        assert calls["merge"] == 2
    finally:
        lib.hwgat_merge_fwd = real
    assert torch.equal(logits, logits2.detach().float())
    for n, p_ in m.named_parameters():
        if p_.grad is not None:
            assert rel_l2(p_.grad, grads[n]) < 1e-4, n


# ------------------------------------------------------------------ K13 / K14: head and loss
@pytest.mark.parametrize("n,d_in,d_out", [(2, 512, 262), (512, 512, 2002), (37, 128, 10), (5000, 64, 33)])
def test_head_linear_f32(n, d_in, d_out):
    from sl_hwgat_b200 import ops
    g = torch.Generator().manual_seed(n + d_out)
    x = torch.randn(n, d_in, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(d_out, d_in, generator=g) * 0.05).cuda().requires_grad_(True)
    b = torch.randn(d_out, generator=g).cuda().requires_grad_(True)
    gy = torch.randn(n, d_out, generator=g).cuda()
    y = ops.linear_f32(x, w, b)
    y.backward(gy)
    xr, wr, br = (t.detach().double().cpu().requires_grad_(True) for t in (x, w, b))
    yr = xr @ wr.t() + br
    yr.backward(gy.double().cpu())
    assert rel_inf(y, yr) < 1e-5
    assert rel_inf(x.grad, xr.grad) < 1e-5 and rel_inf(w.grad, wr.grad) < 1e-5 and rel_inf(b.grad, br.grad) < 1e-5


@pytest.mark.parametrize("rows,classes", [(2, 262), (512, 2002), (37, 10), (3, 5000)])
@pytest.mark.parametrize("smooth", [0.01, 0.0, 0.2])
def test_smooth_cross_entropy_kernel(rows, classes, smooth):
    """K14 against the reference formula (SmoothCrossEntropy.py:35-39, restated in the oracle) in fp64."""
    from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss
    g = torch.Generator().manual_seed(rows + classes)
    z = (torch.randn(rows, classes, generator=g) * 3).cuda().requires_grad_(True)
    t = torch.randint(0, classes, (rows,), generator=g).cuda()
    loss = SmoothedCrossEntropyLoss(smooth)(z, t)
    (loss * 1.7).backward()
    zr = z.detach().double().cpu().requires_grad_(True)
    ref = O.smoothed_cross_entropy(zr, t.cpu(), smooth)
    (ref * 1.7).backward()
    assert abs(loss.item() - ref.item()) < 1e-6 * max(1.0, abs(ref.item()))
    assert rel_inf(z.grad, zr.grad) < 1e-5
    # deterministic: two evaluations give the same bits
    assert SmoothedCrossEntropyLoss(smooth)(z.detach(), t).item() == SmoothedCrossEntropyLoss(smooth)(z.detach(), t).item()


def test_env_autocast_switch_runs_the_bf16_kernels_without_a_source_change():
    """HWGAT_AUTOCAST=bf16: Model.forward opens the autocast region itself, so the reference's unmodified loop
    (model(x) without autocast, utils.py:102) takes the tcgen05 path; off by default."""
    from sl_hwgat_b200.models import HWGATE
    m, cfg, sd = build(16, 10)
    m.eval()
    x = O.synthetic_keypoints(2, 16, 2, seed=3).cuda()
    with torch.no_grad():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            want = m(x)
        fp32 = m(x)
        HWGATE.AUTOCAST = "bf16"
        try:
            got = m(x)
        finally:
            HWGATE.AUTOCAST = ""
    assert torch.equal(got, want) and not torch.equal(got, fp32)
    assert rel_l2(got, fp32) < BF16_TOL


def test_fused_adamw_step_invalidates_cached_bf16_weights():
    """ops.cast_cached keys the bf16 weight copies on the parameter's `_version`; the fused AdamW writes parameters
    through raw pointers (no version bump), so it must drop the cache: the forward after a step sees the new weights."""
    from sl_hwgat_b200.optim import AdamW
    m, cfg, sd = build(16, 10)
    m.train()
    x = O.synthetic_keypoints(2, 16, 2, seed=3).cuda()
    y = O.synthetic_labels(2, 10, seed=3).cuda()
    opt = AdamW(m.parameters(), lr=1e-2)

    def fwd():
        with patched_rand(THR), torch.autocast("cuda", dtype=torch.bfloat16):
            return m(x)
    out0 = fwd()
    O.smoothed_cross_entropy(out0.float(), y).backward()
    opt.step()
    out1 = fwd().detach()
    assert not torch.equal(out0.detach(), out1)
    # the same weights through a fresh cast (cache cleared by hand) give the same bits: nothing stale was used
    from sl_hwgat_b200 import ops
    ops.invalidate_cast_cache()
    assert torch.equal(fwd().detach(), out1)
    # and the update moved the loss down on this batch
    l0 = O.smoothed_cross_entropy(out0.detach().float(), y).item()
    l1 = O.smoothed_cross_entropy(out1.float(), y).item()
    assert l1 < l0, (l0, l1)
