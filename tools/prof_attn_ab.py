"""K2 / K3 timing at the three levels of the B=512, T=64 model (train mode): medians over interleaved rounds.
A/B two builds with HWGAT_B200_LIB=<other .so>:  for i in 1 2; do python tools/prof_attn_ab.py; HWGAT_B200_LIB=... python tools/prof_attn_ab.py; done"""
import statistics
import sys

import torch

sys.path.insert(0, '.')
from tests._util import device_bits
from sl_hwgat_b200 import ops

B = 512
cfgs = [(128, 2, 64), (256, 4, 32), (512, 8, 16)]
state = []
for d, h, F in cfgs:
    x = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16).requires_grad_(True)
    w = (torch.randn(3 * d, d, device='cuda') * 0.05).requires_grad_(True)
    b = (torch.randn(3 * d, device='cuda') * 0.05).requires_grad_(True)
    g = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16)
    state.append((x, w, b, g, device_bits(F, 1), h))
res = {(l, k): [] for l in range(3) for k in ('fwd', 'bwd')}
ev = lambda: torch.cuda.Event(enable_timing=True)
for rnd in range(8):
    for l, (x, w, b, g, bits, h) in enumerate(state):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        y = ops.window_graph_attention(x, w, b, bits, h, shift=1, threshold=0.05)
        e1.record()
        y.backward(g)
        e2.record()
        torch.cuda.synchronize()
        if rnd >= 2:
            res[(l, 'fwd')].append(e0.elapsed_time(e1))
            res[(l, 'bwd')].append(e1.elapsed_time(e2))
print(' '.join(f'L{l}{k} {statistics.median(v):.3f}' for (l, k), v in res.items()))
