"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own,
so these fixtures are outputs of the reference module itself
(/root/reference/hwgat/models/HWGATE.py, model_params.py,
losses/SmoothCrossEntropy.py), imported with a 3-line shim for the missing
``timm`` package (only ``trunc_normal_`` at init, irrelevant because weights
are loaded by state_dict).  Inputs and weights are not stored: they are
regenerated from numpy PCG64 seeds by ``oracle.hwgate_oracle.make_state_dict``
/ ``synthetic_keypoints``.  Large outputs are stored as a strided sample plus
sums.  Everything is computed by the reference in float64 (``.double()``), so
that the discontinuous threshold / ``== 0`` tests are reproducible.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/hwgat"


def import_reference():
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})
    sys.path.insert(0, REF)
    import importlib
    hw = importlib.import_module("models.HWGATE")
    mp = importlib.import_module("models.model_params")
    sys.path.insert(0, os.path.join(REF, "losses"))
    sce = importlib.import_module("SmoothCrossEntropy")
    return hw, mp, sce


class patched_rand:
    """Make the reference's ``torch.rand(1).item()`` (HWGATE.py:96) return the
    injected thresholds, in call order."""

    def __init__(self, values):
        self.values = list(values)
        self.calls = 0

    def __enter__(self):
        self._orig = torch.rand

        def fake(*a, **k):
            v = self.values[self.calls]
            self.calls += 1
            return torch.tensor([v], dtype=torch.float64)

        torch.rand = fake
        return self

    def __exit__(self, *exc):
        torch.rand = self._orig


def sample(t, stride=7):
    a = t.detach().double().reshape(-1).numpy()
    return a[::stride].copy(), np.array([a.sum(), np.abs(a).sum(), float(a.size)])


def build_ref_model(hw, mp, cfg, sd, drop=0.0):
    from oracle import hwgate_oracle as O
    params = mp.HWGATEParams({"num_class": cfg.num_classes, "src_len": cfg.temporal_dim}, cfg.kp_dim, "cpu")
    params.drop_rate = drop
    params.depths, params.num_heads = list(cfg.depths), list(cfg.num_heads)
    params.embed_dim = cfg.embed_dim
    model = hw.Model(*params.get_model_params())
    missing = model.load_state_dict(sd, strict=True)
    model = model.double()
    for layer in model.layers:                       # adj_mat is a plain attribute (HWGATE.py:231)
        layer.adj_mat = layer.adj_mat.double()
        for blk in layer.blocks:
            blk.attn.adj_mat = blk.attn.adj_mat.double()
    return model, params


AUTOCAST_THR = [0.03, 0.05, 0.031, 0.2, 0.033, 0.04, 0.0312, 0.1]
AUTOCAST_STRIDE = 31


def autocast_train_golden(hw, mp, sce):
    """Section 5: the UNMODIFIED reference (fp32 parameters) under torch.autocast("cpu", bfloat16) - the only
    bf16 mode the reference can run in - in TRAIN mode with injected thresholds: logits, loss,
    and for every parameter the gradient norm and a strided sample of the gradient.  This pins what "bf16 training"
    means to the reference itself rather than to a rounding model of ours."""
    from oracle import hwgate_oracle as O
    cfg = O.HWGATEConfig(temporal_dim=64, num_classes=262)
    sd = O.make_state_dict(cfg, seed=1001, weight_std=0.05)
    params = mp.HWGATEParams({"num_class": 262, "src_len": 64}, 2, "cpu")
    params.drop_rate = 0.0
    model = hw.Model(*params.get_model_params())
    model.load_state_dict(sd, strict=True)
    model.train()
    x = O.synthetic_keypoints(2, 64, 2, seed=1001)
    y = O.synthetic_labels(2, 262, seed=1001)
    out = {}
    with patched_rand(AUTOCAST_THR) as pr, torch.autocast("cpu", dtype=torch.bfloat16):
        logits = model(x)
        loss = sce.SmoothedCrossEntropyLoss()(logits.float(), y)
    assert pr.calls == 8
    loss.backward()
    out["logits"] = logits.detach().float().numpy()
    out["loss"] = np.array(loss.item())
    out["thr"] = np.array(AUTOCAST_THR)
    names, norms, samples, offsets = [], [], [], [0]
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().float().reshape(-1).numpy()
        names.append(n)
        norms.append(float(np.linalg.norm(g.astype(np.float64))))
        samples.append(g[::AUTOCAST_STRIDE].astype(np.float32))
        offsets.append(offsets[-1] + samples[-1].size)
    out["gnames"] = np.array(names)
    out["gnorms"] = np.array(norms)
    out["gsamples"] = np.concatenate(samples)
    out["goffsets"] = np.array(offsets)
    out["stride"] = np.array(AUTOCAST_STRIDE)
    # the same model in eval mode under autocast (no threshold): the reference's own bf16 inference
    model.eval()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        out["eval_logits"] = model(x).float().numpy()
    np.savez_compressed(os.path.join(HERE, "autocast_train.npz"), **out)


def ref_params_with_window(mp, W, classes=10, T=16):
    """The reference's HWGATEParams with another window_size: the class hard-codes 16 in __init__ and derives adj_mat
    there, so the attribute is set afterwards and adj_mat re-derived by the reference's own get_adj_mat()
    (model_params.py:373-400), which reads self.window_size."""
    params = mp.HWGATEParams({"num_class": classes, "src_len": T}, 2, "cpu")
    params.window_size = W
    params.adj_mat = torch.tensor(params.get_adj_mat(), dtype=torch.float32)
    params.drop_rate = 0.0
    return params


def larger_windows_golden(hw, mp, sce):
    """Section 6: window_size 32 and 64 (N = 64 / 128 tokens per window; BASELINE configs[4] "larger temporal
    windows"), which the unmodified reference runs (HWGATE.py:30-36, 290-291): masks, attention core, full model."""
    from oracle import hwgate_oracle as O
    out = {}
    for W in (32, 64):
        params = ref_params_with_window(mp, W)
        nW = 64 // W
        out[f"adj_W{W}"] = params.adj_mat.numpy().astype(np.uint8)
        for F in (8, 4):
            for shift in (0, 1):
                blk = hw.PartAttentionBlock(dim=128, num_kps=64, num_heads=2, window_size=W, temporal_patch_size=2,
                                            temporal_dim=F, shift_size=shift, adj_mat=None)
                full = torch.concatenate([params.adj_mat for _ in range(F // 2)])
                if blk.attn_mask is not None:
                    full = full * blk.attn_mask
                out[f"bits_W{W}_F{F}_s{shift}"] = O.pack_mask_bits(full.numpy() != 0)
        for (d, h) in ((128, 2), (256, 4)):
            for shift in (0, 1):
                for thr in (None, 0.04):
                    B, F = 1, 4
                    std = 0.2 if d == 128 else 0.1     # logits of the same spread at both widths (std * sqrt(d))
                    rng = np.random.default_rng(3000 + d + 10 * shift + W)
                    xn = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
                    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
                    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
                    g = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
                    blk = hw.PartAttentionBlock(dim=d, num_kps=64, num_heads=h, window_size=W, temporal_patch_size=2,
                                                temporal_dim=F, shift_size=shift,
                                                adj_mat=torch.concatenate([params.adj_mat] * (F // 2)).double(),
                                                drop=0.0).double()
                    msa = blk.attn
                    with torch.no_grad():
                        msa.qkv.weight.copy_(w); msa.qkv.bias.copy_(b)
                        msa.proj.weight.copy_(torch.eye(d, dtype=torch.float64)); msa.proj.bias.zero_()
                    msa.train(thr is not None)
                    xn_ = xn.clone().requires_grad_(True)
                    xs = torch.roll(xn_, shifts=-shift, dims=1) if shift else xn_
                    xw = hw.window_partition(xs, W, 2)
                    with patched_rand([thr] if thr is not None else []):
                        yw = msa(xw, B, F // 2, nW, mask=blk.attn_mask)
                    y = hw.window_reverse(yw, W, 2, F, 64)
                    y = torch.roll(y, shifts=shift, dims=1) if shift else y
                    (y * g).sum().backward()
                    key = f"W{W}_d{d}_s{shift}_thr{thr}"
                    out[key + "_y"], out[key + "_ysum"] = sample(y, 61)
                    out[key + "_dx"], out[key + "_dxsum"] = sample(xn_.grad, 61)
                    out[key + "_dw"], out[key + "_dwsum"] = sample(msa.qkv.weight.grad, 251)
                    out[key + "_db"] = msa.qkv.bias.grad.numpy().copy()
        # full model, T = 16, 10 classes, eval and train (fp64)
        cfg = O.HWGATEConfig(temporal_dim=16, num_classes=10, window_size=W, edges=O.HWGATEConfig().edges[:nW])
        sd = O.make_state_dict(cfg, seed=1001, weight_std=0.05)
        params = ref_params_with_window(mp, W)
        model = hw.Model(*params.get_model_params())
        model.load_state_dict(sd, strict=True)
        model = model.double()
        for layer in model.layers:
            layer.adj_mat = layer.adj_mat.double()
            for blk in layer.blocks:
                blk.attn.adj_mat = blk.attn.adj_mat.double()
        x = O.synthetic_keypoints(2, 16, 2, seed=1001).double()
        model.eval()
        with torch.no_grad():
            out[f"model_W{W}_eval_logits"] = model(x).numpy()
        model.train()
        with patched_rand(AUTOCAST_THR):
            out[f"model_W{W}_train_logits"] = model(x).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "larger_windows.npz"), **out)


def hgate_golden(mp, sce):
    """Section 7: the sibling model HGATE (hwgat/models/HGATE.py, HGATEParams): 29 keypoints, blocks of 2 x 29 tokens,
    adjacency and shift mask both multiplicative, no threshold drop.  Unmodified reference, fp64."""
    import importlib
    from oracle import hgate_oracle as H
    hg = importlib.import_module("models.HGATE")
    out = {}
    params = mp.HGATEParams({"num_class": 10, "src_len": 16}, 2, "cpu")
    params.drop_rate = 0.0
    out["adj"] = params.adj_mat.numpy().astype(np.uint8)                         # (58, 58)
    for F in (8, 4):
        blk = hg.GraphAttentionBlock(dim=128, num_kps=29, num_heads=2, temporal_patch_size=2, temporal_dim=F,
                                     shift_size=1, adj_mat=None)
        out[f"shift_mask_F{F}"] = (blk.attn_mask.numpy() != 0)
    for (d, h) in ((128, 2), (256, 4)):
        for shift in (0, 1):
            B, F = 2, 4
            std = 0.2 if d == 128 else 0.1
            rng = np.random.default_rng(5000 + d + 10 * shift)
            xn = torch.from_numpy(rng.standard_normal((B, F, 29, d)))
            w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
            b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
            g = torch.from_numpy(rng.standard_normal((B, F, 29, d)))
            blk = hg.GraphAttentionBlock(dim=d, num_kps=29, num_heads=h, temporal_patch_size=2, temporal_dim=F,
                                         shift_size=shift, adj_mat=params.adj_mat.double(), drop=0.0).double()
            msa = blk.attn
            with torch.no_grad():
                msa.qkv.weight.copy_(w); msa.qkv.bias.copy_(b)
                msa.proj.weight.copy_(torch.eye(d, dtype=torch.float64)); msa.proj.bias.zero_()
            xn_ = xn.clone().requires_grad_(True)
            xs = torch.roll(xn_, shifts=-shift, dims=1) if shift else xn_
            xb = hg.block_partition(xs, 2)
            yb = msa(xb, B, F // 2, attn_mask=blk.attn_mask)
            y = hg.block_reverse(yb, 2, F, 29)
            y = torch.roll(y, shifts=shift, dims=1) if shift else y
            (y * g).sum().backward()
            key = f"d{d}_s{shift}"
            out[key + "_y"], out[key + "_ysum"] = sample(y, 53)
            out[key + "_dx"], out[key + "_dxsum"] = sample(xn_.grad, 53)
            out[key + "_dw"], out[key + "_dwsum"] = sample(msa.qkv.weight.grad, 251)
            out[key + "_db"] = msa.qkv.bias.grad.numpy().copy()
    cfg = H.HGATEConfig(temporal_dim=16, num_classes=10)
    sd = H.make_state_dict(cfg, seed=1001, weight_std=0.05)
    model = hg.Model(*params.get_model_params())
    model.load_state_dict(sd, strict=True)
    out["state_dict_names"] = np.array(list(model.state_dict().keys()))
    out["state_dict_shapes"] = np.array([str(tuple(v.shape)) for v in model.state_dict().values()])
    model = model.double()
    for layer in model.layers:
        layer.adj_mat = layer.adj_mat.double()
        for blk in layer.blocks:
            blk.attn.adj_mat = blk.attn.adj_mat.double()
    x = H.synthetic_keypoints(2, 16, seed=1001).double()
    y = torch.from_numpy(np.array([3, 7]))
    model.train()                                   # drop_rate 0: train == eval for HGATE (no threshold path)
    logits = model(x)
    loss = sce.SmoothedCrossEntropyLoss()(logits, y)
    loss.backward()
    out["model_logits"] = logits.detach().numpy()
    out["model_loss"] = np.array(loss.item())
    names, norms, heads = [], [], []
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(n); norms.append(p.grad.norm().item()); heads.append(p.grad.reshape(-1)[:4].numpy().copy())
    out["gnames"], out["gnorms"], out["gheads"] = np.array(names), np.array(norms), np.stack(heads)
    np.savez_compressed(os.path.join(HERE, "hgate.npz"), **out)


def wgate_gate_golden(mp, sce):
    """Section 8: the sibling models WGATE (hwgat/models/WGATE.py, WGATEParams) and GATE (hwgat/models/GATE.py,
    GATEParams): dense attention over all frames with the graph as an additive -10000 mask.  Unmodified reference,
    fp64: the additive masks, MSA forward + backward, the full models with the loss and every parameter gradient."""
    import importlib
    from oracle import wgate_oracle as WG
    wg = importlib.import_module("models.WGATE")
    ga = importlib.import_module("models.GATE")
    out = {}

    class Parent:                      # MSA.forward reads the mask off `parent` by name (WGATE.py:102, GATE.py:60)
        pass

    for name, F in (("wgate", 8), ("gate", 6)):
        if name == "wgate":
            params = mp.WGATEParams({"num_class": 10, "src_len": F}, 2, "cpu")
            K = 64
        else:
            params = mp.GATEParams({"num_class": 10, "src_len": F}, 2, "cpu")
            K = 29
        params.drop_rate = 0.0
        adj = params.adj_mat
        out[name + "_adj"] = adj.numpy().astype(np.uint8)
        mask = adj.masked_fill(adj == 0, float(-10000)).masked_fill(adj == 1, float(0))     # WGATE.py:190 / GATE.py:142
        parent = Parent()
        parent.adj_mask = mask.double() if name == "wgate" else mask.double()[None, None]
        for (d, h) in ((128, 8), (128, 2), (256, 8)):
            B = 2
            std = 0.2 if d == 128 else 0.1
            rng = np.random.default_rng(6000 + d + h + (0 if name == "wgate" else 1))
            xn = torch.from_numpy(rng.standard_normal((B, F, K, d)))
            w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
            b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
            g = torch.from_numpy(rng.standard_normal((B, F, K, d)))
            msa = (wg if name == "wgate" else ga).MSA(h, d, adj_mask="adj_mask").double()
            with torch.no_grad():
                msa.qkv.weight.copy_(w); msa.qkv.bias.copy_(b)
                msa.proj.weight.copy_(torch.eye(d, dtype=torch.float64)); msa.proj.bias.zero_()
            xn_ = xn.clone().requires_grad_(True)
            if name == "wgate":
                y = wg.window_reverse(msa(wg.window_partition(xn_, 16), B, 4, parent), 16, F, K)    # WGATE.py:155-157
            else:
                y = msa(xn_.reshape(B, F * K, d), parent).reshape(B, F, K, d)                       # GATE.py:198, 106
            (y * g).sum().backward()
            key = f"{name}_d{d}_h{h}"
            out[key + "_y"], out[key + "_ysum"] = sample(y, 53)
            out[key + "_dx"], out[key + "_dxsum"] = sample(xn_.grad, 53)
            out[key + "_dw"], out[key + "_dwsum"] = sample(msa.qkv.weight.grad, 251)
            out[key + "_db"] = msa.qkv.bias.grad.numpy().copy()
        # full model: T = 8 (WGATE) / 6 (GATE), depth cut to 3 blocks, fp64, loss and gradients
        params.depths = 3
        cfg = (WG.WGATEConfig if name == "wgate" else WG.GATEConfig)(temporal_dim=F, num_classes=10, depths=3)
        sd = WG.make_state_dict(cfg, seed=1001, weight_std=0.05)
        model = (wg if name == "wgate" else ga).Model(*params.get_model_params())
        model.load_state_dict(sd, strict=True)
        out[name + "_state_dict_names"] = np.array(list(model.state_dict().keys()))
        out[name + "_state_dict_shapes"] = np.array([str(tuple(v.shape)) for v in model.state_dict().values()])
        model = model.double()
        x = WG.synthetic_keypoints(2, F, K, seed=1001).double()
        y = torch.from_numpy(np.array([3, 7]))
        model.train()                               # drop_rate 0: train == eval (neither model has a threshold path)
        logits = model(x)
        loss = sce.SmoothedCrossEntropyLoss()(logits, y)
        loss.backward()
        out[name + "_model_logits"] = logits.detach().numpy()
        out[name + "_model_loss"] = np.array(loss.item())
        names, norms, heads = [], [], []
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            first = np.zeros(4); first[:min(4, p.numel())] = p.grad.reshape(-1)[:4].numpy()     # (weightedAvg.bias has 1)
            names.append(n); norms.append(p.grad.norm().item()); heads.append(first)
        out[name + "_gnames"], out[name + "_gnorms"], out[name + "_gheads"] = np.array(names), np.array(norms), np.stack(heads)
    np.savez_compressed(os.path.join(HERE, "wgate_gate.npz"), **out)


def main():
    from oracle import hwgate_oracle as O
    hw, mp, sce = import_reference()
    torch.manual_seed(0)
    if "--only-wgate" in sys.argv:
        wgate_gate_golden(mp, sce)
        return
    if "--only-hgate" in sys.argv:
        hgate_golden(mp, sce)
        return
    if "--only-autocast" in sys.argv:
        autocast_train_golden(hw, mp, sce)
        return
    if "--only-windows" in sys.argv:
        larger_windows_golden(hw, mp, sce)
        return
    out = {}

    # ---- 1. masks: reference float masks of every default block config, packed
    params = mp.HWGATEParams({"num_class": 10, "src_len": 64}, 2, "cpu")
    adj_ref = params.adj_mat.numpy()                                    # (4,32,32) float
    out_m = {"adj": adj_ref.astype(np.uint8)}
    for F in (64, 32, 16, 8, 4):
        for shift in (0, 1):
            blk = hw.PartAttentionBlock(dim=128, num_kps=64, num_heads=2, window_size=16,
                                        temporal_patch_size=2, temporal_dim=F, shift_size=shift,
                                        adj_mat=None)
            full = torch.concatenate([params.adj_mat for _ in range(F // 2)])   # HWGATE.py:309
            if blk.attn_mask is not None:
                full = full * blk.attn_mask
            bits = O.pack_mask_bits(full.numpy() != 0)
            out_m[f"bits_F{F}_s{shift}"] = bits
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **out_m)

    # ---- 2. index maps on an arange tensor
    x = torch.arange(2 * 8 * 64 * 3, dtype=torch.float64).reshape(2, 8, 64, 3)
    part = hw.window_partition(x, 16, 2)
    rev = hw.window_reverse(part, 16, 2, 8, 64)
    mer = hw.TemporalMerging(3, 2)(x)
    np.savez_compressed(os.path.join(HERE, "index_maps.npz"), partition=part.numpy().astype(np.int32),
                        reverse=rev.numpy().astype(np.int32), merge=mer.numpy().astype(np.int32))

    # ---- 3. attention core (block minus norm/proj/ffn), every level, shift, mode
    core = {}
    for (d, h) in ((128, 2), (256, 4), (512, 8)):
        for shift in (0, 1):
            for thr in (None, 0.02, 0.04, 0.2):
                for std in (0.02, 0.2):
                    if thr in (0.02, 0.2) and std == 0.02:
                        continue
                    B, F = 1, 4
                    rng = np.random.default_rng(1000 + d + 10 * shift + int(std * 100))
                    xn = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
                    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
                    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
                    g = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
                    blk = hw.PartAttentionBlock(dim=d, num_kps=64, num_heads=h, window_size=16,
                                                temporal_patch_size=2, temporal_dim=F, shift_size=shift,
                                                adj_mat=torch.concatenate([params.adj_mat] * (F // 2)).double(),
                                                drop=0.0).double()
                    msa = blk.attn
                    with torch.no_grad():
                        msa.qkv.weight.copy_(w); msa.qkv.bias.copy_(b)
                        msa.proj.weight.copy_(torch.eye(d, dtype=torch.float64)); msa.proj.bias.zero_()
                    msa.train(thr is not None)
                    xn_ = xn.clone().requires_grad_(True)
                    # the same sequence as PartAttentionBlock.forward, HWGATE.py:197-215, without norm1
                    xs = torch.roll(xn_, shifts=-shift, dims=1) if shift else xn_
                    xw = hw.window_partition(xs, 16, 2)
                    with patched_rand([thr] if thr is not None else []):
                        yw = msa(xw, B, F // 2, 4, mask=blk.attn_mask)
                    y = hw.window_reverse(yw, 16, 2, F, 64)
                    y = torch.roll(y, shifts=shift, dims=1) if shift else y
                    (y * g).sum().backward()
                    key = f"d{d}_s{shift}_thr{thr}_std{std}"
                    core[key + "_y"], core[key + "_ysum"] = sample(y, 127)
                    core[key + "_dx"], core[key + "_dxsum"] = sample(xn_.grad, 127)
                    core[key + "_dw"], core[key + "_dwsum"] = sample(msa.qkv.weight.grad, 509)
                    core[key + "_db"] = msa.qkv.bias.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "attention_core.npz"), **core)

    # ---- 4. full model, default hierarchy
    full = {}
    for name, T, classes, B, thr in (
            ("include_eval", 64, 262, 2, None),
            ("include_train", 64, 262, 2, [0.03, 0.05, 0.031, 0.2, 0.033, 0.04, 0.0312, 0.1]),
            ("fdmse_eval", 192, 2002, 1, None)):
        cfg = O.HWGATEConfig(temporal_dim=T, num_classes=classes)
        sd = O.make_state_dict(cfg, seed=1001, weight_std=0.05)
        model, _ = build_ref_model(hw, mp, cfg, sd)
        x = O.synthetic_keypoints(B, T, 2, seed=1001).double()
        y = O.synthetic_labels(B, classes, seed=1001)
        model.train(thr is not None)
        with patched_rand(thr or []) as pr:
            logits = model(x)
        full[name + "_logits"] = logits.detach().numpy()
        if thr is not None:
            assert pr.calls == 8
            loss = sce.SmoothedCrossEntropyLoss()(logits, y)
            loss.backward()
            full[name + "_loss"] = np.array(loss.item())
            full[name + "_thr"] = np.array(thr)
            names, norms, heads = [], [], []
            for n, p in model.named_parameters():
                if p.grad is None:
                    continue
                names.append(n)
                norms.append(p.grad.norm().item())
                heads.append(p.grad.reshape(-1)[:4].numpy().copy())
            full[name + "_gnames"] = np.array(names)
            full[name + "_gnorms"] = np.array(norms)
            full[name + "_gheads"] = np.stack(heads)
    # fp32 reference forward (what a user of the reference actually gets)
    cfg = O.HWGATEConfig(temporal_dim=64, num_classes=262)
    sd = O.make_state_dict(cfg, seed=1001, weight_std=0.05)
    params32 = mp.HWGATEParams({"num_class": 262, "src_len": 64}, 2, "cpu")
    params32.drop_rate = 0.0
    m32 = hw.Model(*params32.get_model_params())
    m32.load_state_dict(sd, strict=True)
    m32.eval()
    with torch.no_grad():
        full["include_eval_logits_fp32"] = m32(O.synthetic_keypoints(2, 64, 2, seed=1001)).numpy()
    full["state_dict_names"] = np.array(list(m32.state_dict().keys()))
    full["state_dict_shapes"] = np.array([str(tuple(v.shape)) for v in m32.state_dict().values()])
    np.savez_compressed(os.path.join(HERE, "full_model.npz"), **full)

    autocast_train_golden(hw, mp, sce)
    larger_windows_golden(hw, mp, sce)
    hgate_golden(mp, sce)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
