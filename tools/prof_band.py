"""K15 / K16 (band_attn.cu) alone on the WGATE / GATE step shape: python tools/prof_band.py {wgate|gate} [--time]"""
import sys, torch
sys.path.insert(0, '.')
import numpy as np
from oracle import wgate_oracle as WG
from sl_hwgat_b200 import ops
name = sys.argv[1] if len(sys.argv) > 1 else "wgate"
B, F, d, h = 512, 64, 128, 8
K, W = (64, 16) if name == "wgate" else (32, 32)
x = torch.randn(B, F, K, d, device='cuda', dtype=torch.bfloat16).requires_grad_(True)
w = (torch.randn(3 * d, d, device='cuda') * 0.05).requires_grad_(True)
b = (torch.randn(3 * d, device='cuda') * 0.05).requires_grad_(True)
g = torch.randn(B, F, K, d, device='cuda', dtype=torch.bfloat16)
adj = WG.wgate_adjacency(WG.WGATEConfig().edges, F, 16) if name == "wgate" else WG.gate_adjacency(WG.GATEConfig().edges, F, 29)
bits = ops.band_mask_pack(torch.from_numpy(WG.additive_mask(adj)).float().cuda(), F, W)
def step():
    y = ops.band_graph_attention(x, w, b, bits, h, W)
    y.backward(g)
for i in range(2):
    step()
torch.cuda.synchronize()
if "--time" in sys.argv:
    from sl_hwgat_b200 import _lib
    lib = _lib.load()
    ev = []
    for fname in ("hwgat_band_attn_fwd", "hwgat_band_attn_bwd"):
        fn = getattr(lib, fname)
        def wrap(fn=fn, fname=fname):
            def call(*a):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = fn(*a); e1.record(); ev.append((fname, e0, e1)); return r
            return call
        setattr(lib, fname, wrap())
    for i in range(5):
        step()
    torch.cuda.synchronize()
    for fname in ("hwgat_band_attn_fwd", "hwgat_band_attn_bwd"):
        t = [a.elapsed_time(b_) for n, a, b_ in ev if n == fname]
        print(f"{name} {fname}: {sum(t) / len(t):.3f} ms")
else:
    step()
    torch.cuda.synchronize()
print('done')
