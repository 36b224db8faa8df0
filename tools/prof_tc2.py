"""K2b / K3b (attn_core_tc2.cu) alone: python tools/prof_tc2.py LEVEL W [--time]"""
import sys, torch
sys.path.insert(0, '.')
import numpy as np
from oracle import hwgate_oracle as O
from sl_hwgat_b200 import ops
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
W = int(sys.argv[2]) if len(sys.argv) > 2 else 16
d, h, F = [(128, 2, 64), (256, 4, 32), (512, 8, 16)][lvl]
B = 512
x = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16).requires_grad_(True)
w = (torch.randn(3 * d, d, device='cuda') * 0.05).requires_grad_(True)
b = (torch.randn(3 * d, device='cuda') * 0.05).requires_grad_(True)
g = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16)
adj = torch.from_numpy(O.window_adjacency(O.HWGATEConfig().edges[:64 // W], W, 2).astype(np.float32)).cuda()
bits = ops.mask_build(adj, F, 1, W, 2)
def step():
    y = ops.window_graph_attention(x, w, b, bits, h, shift=1, threshold=0.05, window=W, impl="tc2")
    y.backward(g)
for i in range(2):
    step()
torch.cuda.synchronize()
if "--time" in sys.argv:
    from sl_hwgat_b200 import _lib
    lib = _lib.load()
    ev = []
    for name in ("hwgat_attn2_fwd", "hwgat_attn2_bwd"):
        fn = getattr(lib, name)
        def wrap(fn=fn, name=name):
            def call(*a):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = fn(*a); e1.record(); ev.append((name, e0, e1)); return r
            return call
        setattr(lib, name, wrap())
    for i in range(5):
        step()
    torch.cuda.synchronize()
    for name in ("hwgat_attn2_fwd", "hwgat_attn2_bwd"):
        t = [a.elapsed_time(b_) for n, a, b_ in ev if n == name]
        print(f"lvl {lvl} W {W} {name}: {sum(t) / len(t):.3f} ms")
print('done')
