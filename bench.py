#!/usr/bin/env python
"""HWGATE fwd+bwd throughput on B200 (BASELINE.json metric: sequences/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME]   # our arm
    python bench.py --impl reference [--steps K] [--warmup W]             # CPU reference arm (oracle port)

--config train512 (default; BASELINE configs[2] / [3]): one step = zero_grad + forward +
SmoothedCrossEntropyLoss + backward of the full HWGATE model (reference hierarchy: depths [2,2,4],
heads [2,4,8], W=16, TP=2) on a synthetic keypoint batch, bf16 autocast, dropout 0.1 and the training
threshold path on (model.train()), as the reference trains.  N > 1: batch-sharded (weak scaling: 512
sequences per GPU) with the NCCL gradient all-reduce inside the step.
--config infer256_t192 (configs[1]): eval forward, batch 256 per GPU, T=192, 2002 classes, bf16; N > 1
shards the batch with no collective.
--config train_t256 / train_t256_w32 / train_t256_w64 (configs[4]): T=256, 128 sequences per GPU
(global 1024 on 8 GPUs), window_size 16 / 32 / 64 ("larger temporal windows": N = 32 / 64 / 128 tokens).
--strong: fixed GLOBAL batch (--batch) split over the ranks.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "HWGAT sequences/sec fwd+bwd"
UNIT = "sequences/s"
T_FRAMES, KPS, CLASSES = 64, 64, 262
DEPTHS, HEADS, EMBED = [2, 2, 4], [2, 4, 8], 128
WINDOW = 16

CONFIGS = {
    # name: (mode, frames, classes, per-GPU batch, window_size, what BASELINE.json calls it)
    "train512": ("train", 64, 262, 512, 16,
                 "BASELINE configs[2]: HWGATE training fwd+bwd, T=64 frames x 64 keypoints x 2, 262 classes "
                 "(INCLUDE shape), depths [2,2,4], heads [2,4,8], W=16, TP=2, train mode (threshold drop + dropout 0.1)"),
    "infer256_t192": ("infer", 192, 2002, 256, 16,
                      "BASELINE configs[1]: HWGATE inference forward, batch 256 x T=192 frames x 64 keypoints x 2, "
                      "2002 classes (FDMSE-ISL shape), bf16, eval mode"),
    "train_t256": ("train", 256, 262, 128, 16,
                   "BASELINE configs[4] at the reference window (W=16): training fwd+bwd, T=256 frames, 128 sequences "
                   "per GPU (1024 on 8 GPUs), 262 classes, train mode"),
    "train_t256_w32": ("train", 256, 262, 128, 32,
                       "BASELINE configs[4], larger windows W=32 (N=64 tokens per window): training fwd+bwd, T=256, "
                       "128 sequences per GPU, 262 classes, train mode"),
    "train_t256_w64": ("train", 256, 262, 128, 64,
                       "BASELINE configs[4], larger windows W=64 (N=128 tokens per window): training fwd+bwd, T=256, "
                       "128 sequences per GPU, 262 classes, train mode"),
    # the sibling model HGATE: 29 keypoints stored as 32, one 64-token window per block
    "hgate_train512": ("train", 64, 262, 512, 32,
                       "sibling model HGATE (hwgat/models/HGATE.py): training fwd+bwd, T=64 frames x 29 keypoints x 2, "
                       "262 classes, depths [2,2,4], heads [2,4,8], TP=2, dropout 0.1"),
    # the sibling models WGATE / GATE: one width, 8 blocks, attention over all frames with an
    # additive graph mask - evaluated here as a 3-frame band (K15 / K16)
    "wgate_train512": ("train", 64, 262, 512, 16,
                       "sibling model WGATE (hwgat/models/WGATE.py): training fwd+bwd, T=64 frames x 64 keypoints x 2, "
                       "262 classes, 8 blocks, d=128, 8 heads, windows of 16 keypoints x all frames, dropout 0.1"),
    "gate_train512": ("train", 64, 262, 512, 32,
                      "sibling model GATE (hwgat/models/GATE.py): training fwd+bwd, T=64 frames x 29 keypoints x 2, "
                      "262 classes, 8 blocks, d=128, 8 heads, attention over all 1856 tokens, dropout 0.1"),
    # The reference's own loop (utils.py:102, 128): fp32, NO autocast.  The drop-in then runs its fp32 path, whose GEMMs
    # are six bf16 tcgen05 products of hi / mid / lo planes (the "x3" mode, csrc/gemm_x3.cu; HWGAT_FP32=ffma selects
    # the true-fp32 FFMA kernels instead).  The incumbent beside it is gpu_eager_baseline.fp32.
    "train128_fp32": ("train", 64, 262, 128, 16,
                      "the reference's unmodified fp32 loop (no autocast): HWGATE training fwd+bwd, T=64, 262 classes, "
                      "128 sequences per GPU, train mode (threshold drop + dropout 0.1), fp32 tensors end to end"),
    "infer256_fp32": ("infer", 64, 262, 256, 16,
                      "the reference's unmodified fp32 evaluation (no autocast): HWGATE eval forward, T=64, 262 classes, "
                      "256 sequences per GPU, fp32 tensors end to end"),
}
MODEL = "HWGATE"
AMP = True          # bf16 autocast around the forward (every config but the *_fp32 ones)


def set_config(name):
    global T_FRAMES, CLASSES, WINDOW, METRIC, MODEL, KPS, AMP
    mode, T_FRAMES, CLASSES, batch, WINDOW, what = CONFIGS[name]
    AMP = not name.endswith("_fp32")
    MODEL = {"hgate": "HGATE", "wgate": "WGATE", "gate": "GATE"}.get(name.split("_")[0], "HWGATE")
    KPS = 32 if MODEL in ("HGATE", "GATE") else 64          # stored keypoints per frame
    METRIC = "HWGAT sequences/sec fwd+bwd" if mode == "train" else "HWGAT sequences/sec inference forward"
    return mode, batch, what


def peaks():
    p = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            m = json.load(open(path))
            p.update({k: float(m[k]) for k in ("bf16_tflops", "bf16_tflops_sustained", "hbm_gbs") if k in m})
            p["source"] = "measured"
        except Exception:
            pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def attn_flops(B, level, backward):
    """algorithmic FLOPs of K2 (x2 for K3) for one block of `level`."""
    F = T_FRAMES >> level
    d, h = EMBED << level, HEADS[level]
    n, nwin, N = B * F * KPS, B * (F // 2) * (KPS // WINDOW), 2 * WINDOW
    fwd = 6.0 * n * d * d + 4.0 * nwin * h * N * N * 64
    return fwd * (2.0 if backward else 1.0)


def merge_bytes(B, level, elem):
    F, d = T_FRAMES >> level, EMBED << level
    return 2.0 * B * F * KPS * d * elem


class KernelTimer:
    """CUDA events around every C-ABI call of the path, recorded on the launching (current) stream."""

    def __init__(self, lib):
        import torch
        self.torch, self.lib, self.on, self.rec = torch, lib, False, []
        self.orig = {}
        for name in ("hwgat_attn_fwd", "hwgat_attn_bwd", "hwgat_attn2_fwd", "hwgat_attn2_bwd", "hwgat_bda_merge_fwd",
                     "hwgat_ln_bwd_unmerge", "hwgat_merge_fwd", "hwgat_merge_bwd", "hwgat_ln_fwd",
                     "hwgat_ln_bwd", "hwgat_bda_ln_fwd", "hwgat_bda_ln_bwd", "hwgat_bias_gelu_dropout_fwd",
                     "hwgat_bias_gelu_dropout_bwd", "hwgat_ffn_fwd", "hwgat_ffn_bwd", "hwgat_proj_fwd",
                     "hwgat_proj_bwd", "hwgat_band_attn_fwd", "hwgat_band_attn_bwd"):
            fn = getattr(lib, name)
            self.orig[name] = fn
            setattr(lib, name, self._wrap(name, fn))

    # (columns of the tensor, extra info) from the C-ABI argument list of each entry point
    DIM_ARG = {
        "hwgat_attn_fwd": lambda a: (int(a[11]), 0), "hwgat_attn_bwd": lambda a: (int(a[14]), 0),
        "hwgat_attn2_fwd": lambda a: (int(a[12]), 0), "hwgat_attn2_bwd": lambda a: (int(a[15]), 0),
        "hwgat_band_attn_fwd": lambda a: (int(a[10]), int(a[7]) * int(a[8]) * int(a[9]), int(a[12])),
        "hwgat_band_attn_bwd": lambda a: (int(a[15]), int(a[12]) * int(a[13]) * int(a[14]), int(a[17])),   # (.., W, diag, stream)
        "hwgat_bda_merge_fwd": lambda a: (int(a[5]), int(a[4]), False),
        "hwgat_ln_bwd_unmerge": lambda a: (int(a[10]), int(a[9])),
        "hwgat_merge_fwd": lambda a: (int(a[5]), 0), "hwgat_merge_bwd": lambda a: (int(a[5]), 0),
        "hwgat_ln_fwd": lambda a: (int(a[7]), int(a[6])), "hwgat_ln_bwd": lambda a: (int(a[10]), int(a[9])),
        "hwgat_bda_ln_fwd": lambda a: (int(a[10]), int(a[9]), bool(a[3])),
        "hwgat_bda_ln_bwd": lambda a: (int(a[12]), int(a[11]), bool(a[5])),
        "hwgat_bias_gelu_dropout_fwd": lambda a: (int(a[4]), int(a[3])),
        "hwgat_bias_gelu_dropout_bwd": lambda a: (int(a[6]), int(a[5])),
        "hwgat_ffn_fwd": lambda a: (int(a[8]), int(a[7]), int(a[9])),
        "hwgat_ffn_bwd": lambda a: (int(a[13]), int(a[12]), int(a[14])),
        "hwgat_proj_fwd": lambda a: (int(a[4]), int(a[3]), int(a[5])),
        "hwgat_proj_bwd": lambda a: (int(a[8]), int(a[7]), int(a[9])),
    }
    # algorithmic HBM bytes per element of the bandwidth-bound kernels
    BYTES_PER_ELEM = {"hwgat_ln_fwd": 6, "hwgat_ln_bwd": 14, "hwgat_ln_bwd_unmerge": 14, "hwgat_bias_gelu_dropout_fwd": 4,
                      "hwgat_bias_gelu_dropout_bwd": 6}

    def _wrap(self, name, fn):
        def call(*a):
            if not self.on:
                return fn(*a)
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a)
            e1.record()
            d = self.DIM_ARG[name](a)
            self.rec.append((name, d, e0, e1))
            return r
        return call

    def restore(self):
        for name, fn in self.orig.items():
            setattr(self.lib, name, fn)

    def table(self, B, steps):
        agg = {}
        for name, d, e0, e1 in self.rec:
            k = (name, d)
            ms = e0.elapsed_time(e1)
            a = agg.setdefault(k, [0.0, 0])
            a[0] += ms
            a[1] += 1
        out = []
        for (name, key), (ms, cnt) in sorted(agg.items()):
            d = key[0]
            avg = ms / cnt
            if "band_attn" in name:
                # QKV projection + banded core (W queries x 3W keys per frame): 6 n d^2 + 4 n 3W d FLOP, backward twice
                n_rows, W = key[1], key[2]
                fl = (6.0 * n_rows * d * d + 12.0 * n_rows * W * d) * (2 if name.endswith("bwd") else 1)
                out.append({"kernel": ("K16 " if name.endswith("bwd") else "K15 ") + f"{name} d={d}", "bound": "tensor",
                            "calls_per_step": cnt / steps, "avg_ms": avg, "alg_flops": fl,
                            "achieved": fl / (avg * 1e-3) / 1e12, "unit": "TFLOP/s", "total_ms": ms})
                continue
            if "attn" in name:
                level = {128: 0, 256: 1, 512: 2}[d]
                fl = attn_flops(B, level, name.endswith("bwd"))
                out.append({"kernel": ("K3 " if name.endswith("bwd") else "K2 ") + f"{name} d={d}", "bound": "tensor",
                            "calls_per_step": cnt / steps, "avg_ms": avg, "alg_flops": fl,
                            "achieved": fl / (avg * 1e-3) / 1e12, "unit": "TFLOP/s", "total_ms": ms})
                continue
            if "ffn" in name:
                # FeedForward GEMMs: fc1 + fc2 forward = 4 n d hidden FLOP, backward twice that
                n_rows, hidden = key[1], key[2]
                fl = 4.0 * n_rows * d * hidden * (2 if name.endswith("bwd") else 1)
                out.append({"kernel": f"K10 {name} d={d}", "bound": "tensor", "calls_per_step": cnt / steps,
                            "avg_ms": avg, "alg_flops": fl, "achieved": fl / (avg * 1e-3) / 1e12, "unit": "TFLOP/s",
                            "total_ms": ms, "attention": False})
                continue
            if "proj" in name:
                # output projection: 2 n d_in d_out FLOP forward, twice that backward
                n_rows, d_out = key[1], key[2]
                fl = 2.0 * n_rows * d * d_out * (2 if name.endswith("bwd") else 1)
                out.append({"kernel": f"K12 {name} d={d}", "bound": "tensor", "calls_per_step": cnt / steps,
                            "avg_ms": avg, "alg_flops": fl, "achieved": fl / (avg * 1e-3) / 1e12, "unit": "TFLOP/s",
                            "total_ms": ms, "attention": False})
                continue
            if name in ("hwgat_merge_fwd", "hwgat_merge_bwd"):
                label, by = "K4 " + f"{name} d={d}", merge_bytes(B, {128: 0, 256: 1}[d], 4)  # fp32 residual stream
            elif "bda_" in name:
                n_rows, with_ln = key[1], key[2]
                per = (12 if with_ln else 10) if name.endswith("fwd") else (16 if with_ln else 6)
                label, by = f"K6 {name} d={d} ln={int(with_ln)}", float(per) * n_rows * d
            else:
                n_rows = key[1]
                label = ("K5 " if "ln" in name else "K7 ") + f"{name} cols={d}"
                by = float(self.BYTES_PER_ELEM[name]) * n_rows * d
            out.append({"kernel": label, "bound": "hbm", "calls_per_step": cnt / steps, "avg_ms": avg,
                        "alg_bytes": by, "achieved": by / (avg * 1e-3) / 1e9, "unit": "GB/s", "total_ms": ms})
        return out


def build_model(device, drop=0.1):
    import torch
    from sl_hwgat_b200.models import HWGATE, model_params
    if MODEL in ("HGATE", "WGATE", "GATE"):
        import importlib
        mod = importlib.import_module("sl_hwgat_b200.models." + MODEL)
        p = getattr(model_params, MODEL + "Params")({"num_class": CLASSES, "src_len": T_FRAMES}, 2, device)
        p.drop_rate = drop
        torch.manual_seed(1001)
        return mod.Model(*p.get_model_params()).to(device)
    p = model_params.HWGATEParams({"num_class": CLASSES, "src_len": T_FRAMES}, 2, device)
    p.drop_rate = drop
    if WINDOW != p.window_size:
        p.set_window_size(WINDOW)
    torch.manual_seed(1001)                       # the reference's seed (configs.py:55)
    return HWGATE.Model(*p.get_model_params()).to(device)


def synthetic_batch(B):
    """U(0,1) keypoints (B,T,29,2) gathered to the 64-slot window layout (dataTransform.py:428-441)."""
    import numpy as np
    import torch
    head, larm, rarm = [0, 1, 2], [3, 5, 7], [4, 6, 8]
    lh, rh = list(range(9, 19)), list(range(19, 29))
    gather = np.array(head + larm + lh + head + rarm + rh + head + larm + rh + head + rarm + lh)
    rng = np.random.default_rng(1001)
    raw = rng.random((B, T_FRAMES, 29, 2), dtype=np.float32)
    x = torch.from_numpy(np.ascontiguousarray(raw if MODEL in ("HGATE", "GATE") else raw[:, :, gather, :]))
    y = torch.from_numpy(rng.integers(0, CLASSES, size=(B,), dtype=np.int64))
    return x, y


def _oracle_setup(batch, device=None):
    """(forward(x, sd, thresholds or None, drop), sd, x, y, n_thresholds) of the oracle for the configured model"""
    import torch
    from oracle import hwgate_oracle as O
    if MODEL == "HGATE":
        from oracle import hgate_oracle as H
        cfg = H.HGATEConfig(temporal_dim=T_FRAMES, num_classes=CLASSES)
        sd = H.make_state_dict(cfg, seed=1001)
        x = H.synthetic_keypoints(batch, T_FRAMES, seed=1001)
        fwd = lambda xx, s_, thr, drop: H.model_forward(xx, s_, cfg, drop=drop, training=thr is not None)
    elif MODEL in ("WGATE", "GATE"):
        from oracle import wgate_oracle as WG
        cfg = (WG.WGATEConfig if MODEL == "WGATE" else WG.GATEConfig)(temporal_dim=T_FRAMES, num_classes=CLASSES)
        sd = WG.make_state_dict(cfg, seed=1001)
        x = WG.synthetic_keypoints(batch, T_FRAMES, cfg.num_kps, seed=1001)
        model_fwd = WG.wgate_forward if MODEL == "WGATE" else WG.gate_forward
        fwd = lambda xx, s_, thr, drop: model_fwd(xx, s_, cfg, drop=drop, training=thr is not None)
    else:
        cfg = O.HWGATEConfig(temporal_dim=T_FRAMES, num_classes=CLASSES, window_size=WINDOW,
                             edges=O.HWGATEConfig().edges[:64 // WINDOW])
        sd = O.make_state_dict(cfg, seed=1001)
        x = O.synthetic_keypoints(batch, T_FRAMES, 2, seed=1001)
        fwd = lambda xx, s_, thr, drop: O.model_forward(xx, s_, cfg, thresholds=thr, drop=drop)
    y = O.synthetic_labels(batch, CLASSES, seed=1001)
    if device is not None:
        sd = {k: v.to(device) for k, v in sd.items()}
        x, y = x.to(device), y.to(device)
    return fwd, sd, x, y, (sum(cfg.depths) if MODEL in ("HWGATE", "HGATE") else 0)


def cpu_reference_arm(steps, warmup, batch=8, mode="train"):
    """The reference's CPU path (oracle port, fp32; train mode with dropout 0.1, or the eval forward) on the host cores."""
    import torch
    from oracle import hwgate_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, sd, x, y, n_thr = _oracle_setup(batch)
    for k, v in sd.items():
        if mode == "train" and k not in ("B", "pos_encoder.pe", "adj_mask") and not k.endswith("attn_mask"):
            v.requires_grad_(True)
    torch.manual_seed(1001)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if mode == "train":
            for v in sd.values():
                v.grad = None
            thr = [torch.rand(1).item() for _ in range(n_thr)]
            loss = O.smoothed_cross_entropy(fwd(x, sd, thr, 0.1), y)
            loss.backward()
        else:
            with torch.no_grad():
                fwd(x, sd, None, 0.0)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port of the reference {MODEL} (fp32, " +
                      ("train mode, dropout 0.1" if mode == "train" else "eval forward") +
                      f"), batch {batch} x T={T_FRAMES} x {29 if MODEL in ('HGATE', 'GATE') else 64} kp x 2, {CLASSES} classes, "
                      f"window_size {WINDOW}, "
                      f"{len(times)} " + ("fwd+bwd" if mode == "train" else "forward") + f" steps after {warmup} warm-up",
            "ms_per_step": 1e3 * total / len(times), "batch": batch}


def gpu_eager_baseline(dev, mode, batch):
    """The incumbent BASELINE.md section 3 names: the reference's op sequence in eager PyTorch on this B200
    (cuBLAS / ATen kernels, no kernel of ours).  /root/reference cannot travel to the box, so the op sequence is the
    oracle's restatement of it (oracle.model_forward: roll, partition, Linear, softmax, ..., one ATen call per
    reference op) run on CUDA tensors, in fp32 (what main.py runs, utils.py:102) and under autocast(bf16), same
    seeds, dropout 0.1, threshold path on.  A reported baseline, timed with CUDA events after our own timed regions."""
    import torch
    from oracle import hwgate_oracle as O
    train = mode == "train"
    fwd, sd, _, _, n_thr = _oracle_setup(8, dev)
    for k, v in sd.items():
        if train and k not in ("B", "pos_encoder.pe", "adj_mask") and not k.endswith("attn_mask"):
            v.requires_grad_(True)
    out = {}
    if MODEL in ("WGATE", "GATE"):
        batch = min(batch, 32)     # dense (F*k)^2 logits: 1 GiB per block per 32 sequences in fp32 at T = 64, kept for backward
    for name, autocast in (("fp32", False), ("bf16_autocast", True)):
        B = batch
        while B >= 8:
            try:
                _, _, x, y, _ = _oracle_setup(B, dev)
                torch.manual_seed(1001)

                def step():
                    if not train:
                        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                            return fwd(x, sd, None, 0.0)
                    for v in sd.values():
                        v.grad = None
                    thr = [torch.rand(1).item() for _ in range(n_thr)]
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                        logits = fwd(x, sd, thr, 0.1)
                    loss = O.smoothed_cross_entropy(logits.float(), y)
                    loss.backward()
                    return loss

                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 3
                e0.record()
                for _ in range(n):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                out[name] = {"value": B / (ms * 1e-3), "unit": UNIT, "batch": B, "ms_per_step": ms,
                             "max_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
                break
            except torch.OutOfMemoryError:
                B //= 2
            finally:
                for v in sd.values():
                    v.grad = None
                x = y = None
                torch.cuda.empty_cache()
    out["what"] = ("eager PyTorch (cuBLAS/ATen) restatement of the reference's op sequence on the same B200, "
                   + ("train fwd+bwd, dropout 0.1, threshold path" if train else "eval forward"))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mode, _, what = set_config(args.config)
    r = cpu_reference_arm(args.steps, args.warmup, mode=mode)
    cfg = workload_config(1, r["batch"], what, mode, False)
    cfg["precision"] = "fp32 on the host cores (oracle port)"
    cfg["note"] = ("a bounded CPU sample of the same workload: batch %d per step on the host cores (the GPU arm runs "
                   "its own per-GPU batch); sequences/s normalises it" % r["batch"])
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _fp32_mode():
    try:
        from sl_hwgat_b200 import ops
        return ops.fp32_mode()
    except Exception:      # the CPU reference arm runs without the library
        return "n/a"


def workload_config(n_gpus, per_gpu_batch, what, mode, strong):
    par = "single GPU"
    if n_gpus > 1:
        par = (f"dp{n_gpus} batch-sharded, NCCL grad all-reduce" if mode == "train"
               else f"dp{n_gpus} batch-sharded inference, no collective")
    return {"workload": what, "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * n_gpus,
            "frames": T_FRAMES, "window_size": WINDOW, "tokens_per_window": 2 * WINDOW, "parallelism": par,
            "scaling_mode": "strong (fixed global batch)" if strong else "weak (fixed per-GPU batch)",
            "precision": "bf16 autocast" if AMP else "fp32, no autocast (fp32 mode %s)" % _fp32_mode(),
            "l2": "activations per block (>= 0.5 GB) exceed the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="train512", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU (default: the config's); with --strong: global")
    ap.add_argument("--strong", action="store_true", help="strong scaling: --batch is the GLOBAL batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--trace-allreduce", action="store_true", help="CUDA events around every gradient bucket")
    ap.add_argument("--frames", type=int, default=0, help="override the config's sequence length T (sweeps)")
    ap.add_argument("--window", type=int, default=0, help="override the config's window_size W: 16, 32, 64 (sweeps)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    mode, cfg_batch, what = set_config(args.config)
    if args.frames or args.window:
        global T_FRAMES, WINDOW
        T_FRAMES, WINDOW = args.frames or T_FRAMES, args.window or WINDOW
        what += f" [overridden for a sweep: T={T_FRAMES}, window_size={WINDOW}]"
    train = mode == "train"

    # stdout carries exactly one JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG >= VERSION, so a
    # bare VERSION setting is dropped and anything NCCL does log goes to stderr
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        del os.environ["NCCL_DEBUG"]
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    from sl_hwgat_b200 import _lib, parallel
    from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU fallback")
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    B = args.batch or cfg_batch
    if args.strong:
        if B % world:
            raise SystemExit("--strong needs a global batch divisible by the number of GPUs")
        B //= world
    model = build_model(dev)
    model = model.train() if train else model.eval()
    parallel.broadcast_parameters(model)
    parallel.sync_threshold_rng(1001, model)      # same threshold sequence on every rank
    criterion = SmoothedCrossEntropyLoss()
    sync = parallel.GradientAllReduce(model, trace=args.trace_allreduce) if (world > 1 and train) else None
    xg, yg = synthetic_batch(B * world)
    sl = parallel.shard_batch(B * world, rank, world)
    x_host, y_host = xg[sl].contiguous().pin_memory(), yg[sl].contiguous().pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    pred_host = torch.empty(B, dtype=torch.int64).pin_memory()

    def step(xd, yd):
        if not train:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=AMP):
                return model(xd)
        if sync is not None:
            sync.zero_grad()
        else:
            model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=AMP):
            logits = model(xd)
        loss = criterion(logits, yd)
        loss.backward()
        if sync is not None:
            sync.finish()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(x_dev, y_dev)
    timer = KernelTimer(lib)
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timer.on = True
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            step(x_dev, y_dev)
        e1.record()
        barrier()
    timer.on = False
    launches = _lib.launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    # gradient all-reduce trace of ONE extra step: when each bucket's collective ran relative to backward
    ar_trace = None
    if sync is not None and args.trace_allreduce:
        barrier()
        t0, t_bwd = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        sync.zero_grad()
        t0.record()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=AMP):
            logits = model(x_dev)
        loss = criterion(logits, y_dev)
        loss.backward()
        t_bwd.record()
        sync.finish()
        t_end.record()
        torch.cuda.synchronize()
        ar_trace = {"backward_done_ms": t0.elapsed_time(t_bwd), "step_done_ms": t0.elapsed_time(t_end),
                    "buckets": [{"bucket": bi, "bytes": nb, "start_ms": a, "end_ms": b}
                                for bi, nb, a, b in sync.trace_report(t0)]}

    # end to end through the public API with host buffers: H2D of the step's inputs from pinned
    # memory and the D2H read of the loss (training, as utils.py:99-109 does) or of the predictions (inference,
    # utils.py:128-133) inside the timed region
    xd, yd = torch.empty_like(x_dev), torch.empty_like(y_dev)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = 0.0
    for _ in range(args.steps):
        xd.copy_(x_host, non_blocking=True)
        if train:
            yd.copy_(y_host, non_blocking=True)
            last = step(xd, yd).item()
        else:
            pred_host.copy_(step(xd, yd).argmax(dim=1))      # blocking D2H of the predicted classes
            last = float(pred_host[0])
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms2 = float(ms2.item())
    timer.restore()

    if rank == 0:
        pk = peaks()
        kern = timer.table(B, args.steps)
        attn = [k for k in kern if k["bound"] == "tensor" and k.pop("attention", True)]
        top = max(attn, key=lambda k: k["total_ms"]) if attn else None
        roof = None
        if top is not None:
            peak = pk["bf16_tflops_sustained"]     # timed inside a long step
            roof = {"kernel": top["kernel"], "bound": "tensor", "achieved": top["achieved"], "peak": peak,
                    "unit": "TFLOP/s", "frac": top["achieved"] / peak, "traffic": None,
                    "peak_source": pk["source"] + " (sustained bf16 GEMM)", "avg_ms": top["avg_ms"]}
            tr = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tr) and args.config == "train512":
                try:
                    roof["traffic"] = json.load(open(tr)).get(top["kernel"])
                except Exception:
                    pass
        tot_fl = sum(k["alg_flops"] * k["calls_per_step"] for k in attn)
        tot_ms = sum(k["total_ms"] for k in attn) / args.steps
        for k in kern:
            k["frac"] = k["achieved"] / (pk["bf16_tflops_sustained"] if k["bound"] == "tensor" else pk["hbm_gbs"])
            k.pop("total_ms", None)
        h2d = x_host.numel() * 4 + (y_host.numel() * 8 if train else 0)
        line = {"metric": METRIC, "value": B * world * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
                "dtype": "bf16" if AMP else "f32", "data": "synthetic",
                "config": workload_config(world, B, what, mode, args.strong),
                "e2e": {"value": B * world * args.steps / (ms2 * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 if train else B * 8,
                        "ms_per_step": ms2 / args.steps, "last_loss" if train else "last_pred": last},
                "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof,
                "attn_tensor_frac": {"alg_tflop_per_step": tot_fl / 1e12, "attn_ms_per_step": tot_ms,
                                     "achieved_tflops": tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms else None,
                                     "frac_of_sustained_peak": (tot_fl / (tot_ms * 1e-3) / 1e12 /
                                                                pk["bf16_tflops_sustained"]) if tot_ms else None},
                "kernels": kern}
        if ar_trace is not None:
            line["allreduce_trace"] = ar_trace
        if world == 1 and not args.no_eager_baseline:
            del model
            torch.cuda.empty_cache()
            try:
                line["gpu_eager_baseline"] = gpu_eager_baseline(dev, mode, min(B, 128))
                eb = line["gpu_eager_baseline"].get("bf16_autocast") or line["gpu_eager_baseline"].get("fp32")
                if eb and AMP:
                    line["speedup_vs_gpu_eager_bf16"] = line["value"] / eb["value"]
                if not AMP and line["gpu_eager_baseline"].get("fp32"):
                    line["speedup_vs_gpu_eager_fp32"] = line["value"] / line["gpu_eager_baseline"]["fp32"]["value"]
            except Exception as exc:      # a reported baseline must not take the bench line down with it
                line["gpu_eager_baseline"] = {"error": repr(exc)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_arm(3, 1, mode=mode)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
