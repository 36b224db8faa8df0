"""fused (K2 + K3) against hybrid (K2 keeping q, k, v + K3b) and tc2 (K2b + K3b), alone, CUDA events: python tools/prof_hybrid.py"""
import sys, torch
sys.path.insert(0, '.')
from tests._util import device_bits
from sl_hwgat_b200 import ops
B = 512
for lvl, (d, h, F) in enumerate([(128, 2, 64), (256, 4, 32), (512, 8, 16)]):
    x = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16).requires_grad_(True)
    w = (torch.randn(3 * d, d, device='cuda') * 0.05).requires_grad_(True)
    b = (torch.randn(3 * d, device='cuda') * 0.05).requires_grad_(True)
    g = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16)
    bits = device_bits(F, 1)
    for impl in ("fused", "hybrid", "tc2", "fused", "hybrid"):
        def fwd():
            return ops.window_graph_attention(x, w, b, bits, h, shift=1, threshold=0.05, impl=impl)
        for _ in range(2):
            fwd().backward(g)
        torch.cuda.synchronize()
        tf = tb = 0.0
        n = 5
        for _ in range(n):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); y = fwd(); e[1].record(); y.backward(g); e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        print(f"level {lvl} d={d} {impl:7s} fwd {tf / n:.3f} ms  bwd {tb / n:.3f} ms  sum {(tf + tb) / n:.3f}", flush=True)
    del x, w, b, g
    torch.cuda.empty_cache()
