"""Pins oracle/hwgate_oracle.py to the reference's own outputs
(tests/golden/*.npz, produced by tests/golden/make_golden.py from the
unmodified reference).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import hwgate_oracle as O

# packed rows of one window as listed in SURVEY.md section 8 row a1 [probe]
SURVEY_ROWS_0_15 = [0x1000f, 0x20003, 0x40005, 0x80019, 0x100038, 0x200070, 0x4055e0, 0x80aac0,
                    0x1000740, 0x2000b80, 0x4001d40, 0x8002e80, 0x10007440, 0x2000b880,
                    0x4000d040, 0x8000e080]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_adjacency_matches_reference(golden_dir):
    g = _load(golden_dir, "masks.npz")
    cfg = O.HWGATEConfig()
    adj = O.window_adjacency(cfg.edges, 16, 2)
    assert adj.shape == (4, 32, 32)
    assert np.array_equal(adj, g["adj"].astype(bool))
    assert int(adj[0].sum()) == 164
    bits = O.pack_mask_bits(adj)[0, :, 0]
    assert [int(b) for b in bits[:16]] == SURVEY_ROWS_0_15
    # rows 16-31: same pattern with the halves swapped
    swapped = [((r >> 16) | (r << 16)) & 0xFFFFFFFF for r in SURVEY_ROWS_0_15]
    assert [int(b) for b in bits[16:]] == swapped


@pytest.mark.parametrize("F", [64, 32, 16, 8, 4])
@pytest.mark.parametrize("shift", [0, 1])
def test_combined_mask_bit_exact(golden_dir, F, shift):
    g = _load(golden_dir, "masks.npz")
    cfg = O.HWGATEConfig()
    adj = O.window_adjacency(cfg.edges, 16, 2)
    m = O.combined_mask(adj, F, 16, 2, shift)
    assert np.array_equal(O.pack_mask_bits(m), g[f"bits_F{F}_s{shift}"])


def test_index_maps(golden_dir):
    g = _load(golden_dir, "index_maps.npz")
    x = torch.arange(2 * 8 * 64 * 3, dtype=torch.float64).reshape(2, 8, 64, 3)
    part = O.window_partition(x, 16, 2)
    assert np.array_equal(part.numpy().astype(np.int32), g["partition"])
    assert torch.equal(O.window_reverse(part, 16, 2, 8, 64), x)
    mer = O.temporal_merge(x, 2)
    assert np.array_equal(mer.numpy().astype(np.int32), g["merge"])
    assert torch.equal(O.temporal_merge_backward(mer, 2), x)


def _core_cases():
    for (d, h) in ((128, 2), (256, 4), (512, 8)):
        for shift in (0, 1):
            for thr in (None, 0.02, 0.04, 0.2):
                for std in (0.02, 0.2):
                    if thr in (0.02, 0.2) and std == 0.02:
                        continue
                    yield d, h, shift, thr, std


def core_inputs(d, shift, std, B=1, F=4):
    rng = np.random.default_rng(1000 + d + 10 * shift + int(std * 100))
    xn = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
    return xn, w, b, g


@pytest.mark.parametrize("d,h,shift,thr,std", list(_core_cases()))
def test_attention_core_matches_reference(golden_dir, d, h, shift, thr, std):
    G = _load(golden_dir, "attention_core.npz")
    key = f"d{d}_s{shift}_thr{thr}_std{std}"
    xn, w, b, g = core_inputs(d, shift, std)
    adj = O.window_adjacency(O.HWGATEConfig().edges, 16, 2)
    mask = O.combined_mask(adj, 4, 16, 2, shift)
    xn_ = xn.clone().requires_grad_(True)
    w_ = w.clone().requires_grad_(True)
    b_ = b.clone().requires_grad_(True)
    y = O.attention_core(xn_, w_, b_, h, mask, 16, 2, shift, thr)
    (y * g).sum().backward()

    def chk(t, name, stride):
        a = t.detach().reshape(-1).numpy()
        np.testing.assert_allclose(a[::stride], G[key + "_" + name], rtol=1e-9, atol=1e-11)
        s = G[key + "_" + name + "sum"]
        np.testing.assert_allclose([a.sum(), np.abs(a).sum(), a.size], s, rtol=1e-9, atol=1e-9)

    chk(y, "y", 127)
    chk(xn_.grad, "dx", 127)
    chk(w_.grad, "dw", 509)
    np.testing.assert_allclose(b_.grad.numpy(), G[key + "_db"], rtol=1e-9, atol=1e-10)

    # closed-form backward == autograd (this is the formula K3 implements)
    dx, dw, db = O.attention_core_backward(xn, w, b, h, mask, 16, 2, shift, thr, g)
    np.testing.assert_allclose(dx.numpy(), xn_.grad.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(dw.numpy(), w_.grad.numpy(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(db.numpy(), b_.grad.numpy(), rtol=1e-9, atol=1e-9)


def test_fully_masked_rows_are_uniform():
    """threshold below 1/32 drops every logit: softmax over 32 fills of -10000
    is uniform over all 32 keys, neighbours or not."""
    d, h = 128, 2
    xn, w, b, g = core_inputs(d, 0, 0.02)
    adj = O.window_adjacency(O.HWGATEConfig().edges, 16, 2)
    mask = O.combined_mask(adj, 4, 16, 2, 0)
    y = O.attention_core(xn, w, b, h, mask, 16, 2, 0, threshold=0.001)
    xw = O.window_partition(xn, 16, 2)
    v = (xw @ w.t() + b)[..., 2 * d:]
    expect = O.window_reverse(v.mean(dim=1, keepdim=True).expand(-1, 32, -1), 16, 2, 4, 64)
    np.testing.assert_allclose(y.numpy(), expect.numpy(), rtol=1e-10, atol=1e-12)


def test_full_model_eval_matches_reference(golden_dir):
    G = _load(golden_dir, "full_model.npz")
    cfg = O.HWGATEConfig(temporal_dim=64, num_classes=262)
    sd = O.make_state_dict(cfg, seed=1001, weight_std=0.05)
    assert list(G["state_dict_names"]) == list(sd.keys())
    assert [str(tuple(v.shape)) for v in sd.values()] == list(G["state_dict_shapes"])
    x = O.synthetic_keypoints(2, 64, 2, seed=1001)
    sd64 = {k: v.double() for k, v in sd.items()}
    logits = O.model_forward(x.double(), sd64, cfg)
    np.testing.assert_allclose(logits.numpy(), G["include_eval_logits"], rtol=1e-9, atol=1e-10)
    # fp32 oracle against the fp32 reference: the north_star's fp32 tolerance
    l32 = O.model_forward(x, sd, cfg).numpy()
    ref32 = G["include_eval_logits_fp32"]
    assert np.abs(l32 - ref32).max() / np.abs(ref32).max() < 1e-5


def test_full_model_train_matches_reference(golden_dir):
    G = _load(golden_dir, "full_model.npz")
    cfg = O.HWGATEConfig(temporal_dim=64, num_classes=262)
    sd = {k: v.double().requires_grad_(v.dtype.is_floating_point and k not in ("B", "pos_encoder.pe")
                                        and not k.endswith("attn_mask"))
          for k, v in O.make_state_dict(cfg, seed=1001, weight_std=0.05).items()}
    x = O.synthetic_keypoints(2, 64, 2, seed=1001).double()
    y = O.synthetic_labels(2, 262, seed=1001)
    thr = [float(t) for t in G["include_train_thr"]]
    logits = O.model_forward(x, sd, cfg, thresholds=thr)
    np.testing.assert_allclose(logits.detach().numpy(), G["include_train_logits"], rtol=1e-9, atol=1e-10)
    loss = O.smoothed_cross_entropy(logits, y)
    np.testing.assert_allclose(loss.item(), float(G["include_train_loss"]), rtol=1e-10)
    loss.backward()
    for name, norm, head in zip(G["include_train_gnames"], G["include_train_gnorms"], G["include_train_gheads"]):
        g = sd[str(name)].grad
        np.testing.assert_allclose(g.norm().item(), norm, rtol=1e-8, err_msg=str(name))
        np.testing.assert_allclose(g.reshape(-1)[:4].numpy(), head, rtol=1e-7, atol=1e-12, err_msg=str(name))


def test_full_model_long_sequence_eval(golden_dir):
    G = _load(golden_dir, "full_model.npz")
    cfg = O.HWGATEConfig(temporal_dim=192, num_classes=2002)
    sd = {k: v.double() for k, v in O.make_state_dict(cfg, seed=1001, weight_std=0.05).items()}
    x = O.synthetic_keypoints(1, 192, 2, seed=1001).double()
    logits = O.model_forward(x, sd, cfg)
    np.testing.assert_allclose(logits.numpy(), G["fdmse_eval_logits"], rtol=1e-9, atol=1e-10)


def test_bf16_rounding_model_is_close_to_fp32():
    """The rounding points used by the bf16 kernels stay inside the
    north_star's 2e-2 relative tolerance against the unrounded oracle."""
    d, h = 256, 4
    xn, w, b, g = core_inputs(d, 1, 0.2)
    adj = O.window_adjacency(O.HWGATEConfig().edges, 16, 2)
    mask = O.combined_mask(adj, 4, 16, 2, 1)
    y = O.attention_core(xn.float(), w.float(), b.float(), h, mask, 16, 2, 1)
    yb = O.attention_core(xn.float(), w.float(), b.float(), h, mask, 16, 2, 1, bf16_points=True)
    assert (y - yb).norm() / y.norm() < 2e-2


# ------------------------------------------------------------------ larger windows (W = 32, 64; BASELINE configs[4])
def _adj_w(W):
    return O.window_adjacency(O.HWGATEConfig().edges[:64 // W], W, 2)


def window_core_inputs(d, shift, W, B=1, F=4, std=None):
    """Same seeded inputs as tests/golden/make_golden.py section 6."""
    std = (0.2 if d == 128 else 0.1) if std is None else std
    rng = np.random.default_rng(3000 + d + 10 * shift + W)
    xn = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
    return xn, w, b, g


@pytest.mark.parametrize("W", [32, 64])
def test_larger_window_masks_bit_exact(golden_dir, W):
    G = _load(golden_dir, "larger_windows.npz")
    adj = _adj_w(W)
    assert adj.shape == (64 // W, 2 * W, 2 * W)
    assert np.array_equal(adj, G[f"adj_W{W}"].astype(bool))
    for F in (8, 4):
        for shift in (0, 1):
            assert np.array_equal(O.pack_mask_bits(O.combined_mask(adj, F, W, 2, shift)), G[f"bits_W{W}_F{F}_s{shift}"])


@pytest.mark.parametrize("W", [32, 64])
@pytest.mark.parametrize("d,h", [(128, 2), (256, 4)])
@pytest.mark.parametrize("shift", [0, 1])
@pytest.mark.parametrize("thr", [None, 0.04])
def test_larger_window_attention_core_matches_reference(golden_dir, W, d, h, shift, thr):
    G = _load(golden_dir, "larger_windows.npz")
    key = f"W{W}_d{d}_s{shift}_thr{thr}"
    xn, w, b, g = window_core_inputs(d, shift, W)
    mask = O.combined_mask(_adj_w(W), 4, W, 2, shift)
    x_ = xn.clone().requires_grad_(True)
    w_ = w.clone().requires_grad_(True)
    b_ = b.clone().requires_grad_(True)
    y = O.attention_core(x_, w_, b_, h, mask, W, 2, shift, thr)
    (y * g).sum().backward()
    cdx, cdw, cdb = O.attention_core_backward(xn, w, b, h, mask, W, 2, shift, thr, g)

    def chk(t, name, stride):
        a = t.detach().reshape(-1).numpy()
        ref = G[key + "_" + name]
        assert np.abs(a[::stride] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (key, name)
        s = G[key + "_" + name + "sum"]
        assert abs(a.sum() - s[0]) <= 1e-9 * s[1] and a.size == int(s[2])

    chk(y, "y", 61)
    chk(x_.grad, "dx", 61)
    chk(w_.grad, "dw", 251)
    assert np.abs(b_.grad.numpy() - G[key + "_db"]).max() <= 1e-9 * np.abs(G[key + "_db"]).max()
    # closed-form backward == autograd
    assert torch.allclose(cdx, x_.grad, rtol=1e-9, atol=1e-12) and torch.allclose(cdw, w_.grad, rtol=1e-9, atol=1e-11)
    assert torch.allclose(cdb, b_.grad, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("W", [32, 64])
def test_larger_window_full_model_matches_reference(golden_dir, W):
    G = _load(golden_dir, "larger_windows.npz")
    cfg = O.HWGATEConfig(temporal_dim=16, num_classes=10, window_size=W, edges=O.HWGATEConfig().edges[:64 // W])
    sd = {k: v.double() for k, v in O.make_state_dict(cfg, seed=1001, weight_std=0.05).items()}
    x = O.synthetic_keypoints(2, 16, 2, seed=1001).double()
    ev = O.model_forward(x, sd, cfg)
    assert np.abs(ev.numpy() - G[f"model_W{W}_eval_logits"]).max() <= 1e-9 * np.abs(G[f"model_W{W}_eval_logits"]).max()
    thr = [0.03, 0.05, 0.031, 0.2, 0.033, 0.04, 0.0312, 0.1]
    tr = O.model_forward(x, sd, cfg, thresholds=thr)
    assert np.abs(tr.numpy() - G[f"model_W{W}_train_logits"]).max() <= 1e-9 * np.abs(G[f"model_W{W}_train_logits"]).max()


def test_autocast_golden_is_consistent_with_fp64_golden(golden_dir):
    """tests/golden/autocast_train.npz (the reference under torch.autocast(bfloat16), train mode) sits within the
    bf16 band of the fp64 goldens of the same case - the band the GPU bf16 tests are judged against."""
    A, F = _load(golden_dir, "autocast_train.npz"), _load(golden_dir, "full_model.npz")
    assert list(A["thr"]) == list(F["include_train_thr"])
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    assert rel(A["logits"], F["include_train_logits"]) < 2.5e-2
    assert rel(A["eval_logits"], F["include_eval_logits"]) < 2e-2
    assert list(A["gnames"]) == list(F["include_train_gnames"])
    assert np.abs(A["gnorms"] - F["include_train_gnorms"]).max() / F["include_train_gnorms"].max() < 2e-2
    assert int(A["goffsets"][-1]) == A["gsamples"].size


# ------------------------------------------------------------------ sibling model HGATE
def hgate_core_inputs(d, shift, B=2, F=4):
    """Same seeded inputs as tests/golden/make_golden.py section 7."""
    std = 0.2 if d == 128 else 0.1
    rng = np.random.default_rng(5000 + d + 10 * shift)
    xn = torch.from_numpy(rng.standard_normal((B, F, 29, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, 29, d)))
    return xn, w, b, g


def test_hgate_masks_match_reference(golden_dir):
    from oracle import hgate_oracle as H
    G = _load(golden_dir, "hgate.npz")
    adj = H.block_adjacency(H.HGATEConfig().edges, 29, 2)
    assert adj.shape == (58, 58) and np.array_equal(adj, G["adj"].astype(bool))
    assert len(H.HGATE_EDGES) == 34
    for F in (8, 4):
        assert np.array_equal(H.block_shift_mask(F, 29, 2, 1), G[f"shift_mask_F{F}"])


@pytest.mark.parametrize("d,h", [(128, 2), (256, 4)])
@pytest.mark.parametrize("shift", [0, 1])
def test_hgate_attention_core_matches_reference(golden_dir, d, h, shift):
    from oracle import hgate_oracle as H
    G = _load(golden_dir, "hgate.npz")
    key = f"d{d}_s{shift}"
    xn, w, b, g = hgate_core_inputs(d, shift)
    mask = H.block_mask(H.block_adjacency(H.HGATEConfig().edges, 29, 2), 4, 29, 2, shift)
    x_, w_, b_ = (t.clone().requires_grad_(True) for t in (xn, w, b))
    y = H.attention_core(x_, w_, b_, h, mask, 2, shift)
    (y * g).sum().backward()

    def chk(t, name, stride):
        a = t.detach().reshape(-1).numpy()
        ref = G[key + "_" + name]
        assert np.abs(a[::stride] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (key, name)
        s = G[key + "_" + name + "sum"]
        assert abs(a.sum() - s[0]) <= 1e-9 * s[1] and a.size == int(s[2])
    chk(y, "y", 53)
    chk(x_.grad, "dx", 53)
    chk(w_.grad, "dw", 251)
    assert np.abs(b_.grad.numpy() - G[key + "_db"]).max() <= 1e-9 * np.abs(G[key + "_db"]).max()


def test_hgate_full_model_matches_reference(golden_dir):
    from oracle import hgate_oracle as H
    G = _load(golden_dir, "hgate.npz")
    cfg = H.HGATEConfig(temporal_dim=16, num_classes=10)
    sd = H.make_state_dict(cfg, seed=1001, weight_std=0.05)
    assert sorted(sd.keys()) == sorted(G["state_dict_names"])
    shapes = dict(zip(G["state_dict_names"], G["state_dict_shapes"]))
    assert all(str(tuple(v.shape)) == shapes[k] for k, v in sd.items())
    sd64 = {k: v.double().requires_grad_(k not in ("B", "pos_encoder.pe") and not k.endswith("attn_mask"))
            for k, v in sd.items()}
    x = H.synthetic_keypoints(2, 16, seed=1001).double()
    logits = H.model_forward(x, sd64, cfg)
    assert np.abs(logits.detach().numpy() - G["model_logits"]).max() <= 1e-9 * np.abs(G["model_logits"]).max()
    loss = O.smoothed_cross_entropy(logits, torch.tensor([3, 7]))
    assert abs(loss.item() - float(G["model_loss"])) < 1e-9
    loss.backward()
    for name, norm, head in zip(G["gnames"], G["gnorms"], G["gheads"]):
        gr = sd64[str(name)].grad
        assert abs(gr.norm().item() - norm) <= 1e-8 * norm, name
        assert np.abs(gr.reshape(-1)[:4].numpy() - head).max() <= 1e-8 * max(np.abs(head).max(), 1e-12), name


# ------------------------------------------------------------------ sibling models WGATE and GATE (section 8)
def band_core_inputs(name, d, h, B=2):
    """Same seeded inputs as tests/golden/make_golden.py section 8."""
    F, K = (8, 64) if name == "wgate" else (6, 29)
    std = 0.2 if d == 128 else 0.1
    rng = np.random.default_rng(6000 + d + h + (0 if name == "wgate" else 1))
    xn = torch.from_numpy(rng.standard_normal((B, F, K, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, K, d)))
    return xn, w, b, g


def band_adjacency(name):
    from oracle import wgate_oracle as WG
    if name == "wgate":
        return WG.wgate_adjacency(WG.WGATEConfig().edges, 8, 16)
    return WG.gate_adjacency(WG.GATEConfig().edges, 6, 29)


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_wgate_gate_adjacency_matches_reference(golden_dir, name):
    G = _load(golden_dir, "wgate_gate.npz")
    adj = band_adjacency(name)
    assert adj.shape == G[name + "_adj"].shape and np.array_equal(adj != 0, G[name + "_adj"] != 0)
    assert set(np.unique(adj)) == {0.0, 1.0}


@pytest.mark.parametrize("name", ["wgate", "gate"])
@pytest.mark.parametrize("d,h", [(128, 8), (128, 2), (256, 8)])
def test_wgate_gate_attention_core_matches_reference(golden_dir, name, d, h):
    from oracle import wgate_oracle as WG
    G = _load(golden_dir, "wgate_gate.npz")
    key = f"{name}_d{d}_h{h}"
    xn, w, b, g = band_core_inputs(name, d, h)
    mask = WG.additive_mask(band_adjacency(name))
    x_, w_, b_ = (t.clone().requires_grad_(True) for t in (xn, w, b))
    if name == "wgate":
        y = WG.wgate_attention_core(x_, w_, b_, h, mask, 16)
    else:
        y = WG.gate_attention_core(x_, w_, b_, h, mask)
    (y * g).sum().backward()

    def chk(t, nm, stride):
        a = t.detach().reshape(-1).numpy()
        ref = G[key + "_" + nm]
        assert np.abs(a[::stride] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (key, nm)
        s = G[key + "_" + nm + "sum"]
        assert abs(a.sum() - s[0]) <= 1e-9 * s[1] and a.size == int(s[2])
    chk(y, "y", 53)
    chk(x_.grad, "dx", 53)
    chk(w_.grad, "dw", 251)
    assert np.abs(b_.grad.numpy() - G[key + "_db"]).max() <= 1e-9 * np.abs(G[key + "_db"]).max()


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_wgate_gate_full_model_matches_reference(golden_dir, name):
    from oracle import wgate_oracle as WG
    G = _load(golden_dir, "wgate_gate.npz")
    F, K = (8, 64) if name == "wgate" else (6, 29)
    cfg = (WG.WGATEConfig if name == "wgate" else WG.GATEConfig)(temporal_dim=F, num_classes=10, depths=3)
    sd = WG.make_state_dict(cfg, seed=1001, weight_std=0.05)
    assert sorted(sd.keys()) == sorted(G[name + "_state_dict_names"])
    shapes = dict(zip(G[name + "_state_dict_names"], G[name + "_state_dict_shapes"]))
    assert all(str(tuple(v.shape)) == shapes[k] for k, v in sd.items())
    sd64 = {k: v.double().requires_grad_(k not in ("B", "pos_encoder.pe", "adj_mask")) for k, v in sd.items()}
    x = WG.synthetic_keypoints(2, F, K, seed=1001).double()
    logits = (WG.wgate_forward if name == "wgate" else WG.gate_forward)(x, sd64, cfg)
    ref = G[name + "_model_logits"]
    assert np.abs(logits.detach().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()
    loss = O.smoothed_cross_entropy(logits, torch.tensor([3, 7]))
    assert abs(loss.item() - float(G[name + "_model_loss"])) < 1e-9
    loss.backward()
    for pname, norm, head in zip(G[name + "_gnames"], G[name + "_gnorms"], G[name + "_gheads"]):
        gr = sd64[str(pname)].grad
        assert abs(gr.norm().item() - norm) <= 1e-8 * norm, pname
        got = np.zeros(4); got[:min(4, gr.numel())] = gr.reshape(-1)[:4].numpy()
        assert np.abs(got - head).max() <= 1e-8 * max(np.abs(head).max(), 1e-12), pname


@pytest.mark.parametrize("name", ["wgate", "gate"])
def test_band_restriction_equals_dense_additive_softmax(name):
    """The claim the product kernels rest on (band_attn.cu header): with the reference's graphs, softmax(logits +
    additive mask) over all F*k tokens equals the softmax over the 3-frame band's edges alone - in fp32 the -10000
    terms underflow to exactly 0 once the row maximum (an edge) is subtracted."""
    from oracle import wgate_oracle as WG
    adj = band_adjacency(name)
    adj = adj[0] if name == "wgate" else adj
    N = adj.shape[0]
    rng = np.random.default_rng(7)
    s = torch.from_numpy(rng.standard_normal((N, N)).astype(np.float32) * 8.0)            # logits up to ~ +-35
    dense = torch.softmax(s + torch.from_numpy(WG.additive_mask(adj).astype(np.float32)), dim=-1)
    edges = torch.from_numpy(adj != 0)
    banded = torch.softmax(torch.where(edges, s, torch.full_like(s, float("-inf"))), dim=-1)
    assert torch.equal(dense == 0, ~edges)                     # every non-edge weighs exactly zero
    assert torch.allclose(dense, banded, rtol=1e-6, atol=0)
    assert bool(edges.any(dim=-1).all())                       # no empty row (else the softmax would go dense)
