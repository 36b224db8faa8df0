"""Shared helpers of the GPU parity tests."""
import numpy as np
import torch

from oracle import hwgate_oracle as O

CFG = O.HWGATEConfig()
ADJ = O.window_adjacency(CFG.edges, 16, 2)          # (4,32,32) bool, pinned to the reference by test_oracle_golden


def rel_inf(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  - the 'relative' of the north_star tolerances."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def core_inputs(d, shift, std, B=1, F=4):
    """Same seeded inputs as tests/golden/make_golden.py section 3."""
    rng = np.random.default_rng(1000 + d + 10 * shift + int(std * 100))
    xn = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
    w = torch.from_numpy(rng.standard_normal((3 * d, d)) * std)
    b = torch.from_numpy(rng.standard_normal((3 * d,)) * 0.1)
    g = torch.from_numpy(rng.standard_normal((B, F, 64, d)))
    return xn, w, b, g


def oracle_core(xn, w, b, g, heads, F, shift, thr, bf16_points=False):
    """fp64 oracle forward + closed-form backward of the attention core."""
    mask = O.combined_mask(ADJ, F, 16, 2, shift)
    y = O.attention_core(xn.double(), w.double(), b.double(), heads, mask, 16, 2, shift, thr, bf16_points)
    dx, dw, db = O.attention_core_backward(xn.double(), w.double(), b.double(), heads, mask, 16, 2, shift, thr,
                                           g.double())
    return y, dx, dw, db


def device_bits(F, shift, dev="cuda"):
    from sl_hwgat_b200 import ops
    adj = torch.from_numpy(ADJ.astype(np.float32)).to(dev)
    return ops.mask_build(adj, F, shift)


def cuda_core(xn, w, b, g, heads, shift, thr, dtype, layout=0):
    """K2 + K3 through the C ABI (sl_hwgat_b200.ops -> ctypes -> libhwgat_b200.so)."""
    from sl_hwgat_b200 import ops
    B, F, K, d = xn.shape
    bits = device_bits(F, shift)
    x_ = xn.to("cuda", dtype).requires_grad_(True)
    w_ = w.to("cuda", torch.float32).requires_grad_(True)
    b_ = b.to("cuda", torch.float32).requires_grad_(True)
    if dtype == torch.bfloat16:
        # weights are parameters kept in fp32 and cast inside the op; feed the bf16-rounded
        # values so that the oracle and the kernel see identical numbers
        w_ = w.to(torch.bfloat16).float().to("cuda").requires_grad_(True)
    y = ops.window_graph_attention(x_, w_, b_, bits, heads, shift=shift, threshold=thr, layout=layout)
    y.backward(g.to("cuda", dtype))
    torch.cuda.synchronize()
    return y.detach(), x_.grad, w_.grad, b_.grad
