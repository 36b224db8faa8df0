"""The fp32 attention (x3 QKV + FFMA core) forward + backward at one level of the batch-128 model, for ncu."""
import sys, torch
sys.path.insert(0, '.')
import numpy as np
from sl_hwgat_b200 import ops
from tests._util import device_bits
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
d, heads, F = [128, 256, 512][lvl], [2, 4, 8][lvl], 64 >> lvl
B = 128
xn = torch.randn(B, F, 64, d, device='cuda', requires_grad=True)
w = (torch.randn(3 * d, d, device='cuda') / d ** 0.5).requires_grad_(True)
b = torch.zeros(3 * d, device='cuda', requires_grad=True)
g = torch.randn(B, F, 64, d, device='cuda')
bits = device_bits(F, 0)
for _ in range(2):
    ops.window_graph_attention(xn, w, b, bits, heads, shift=0, threshold=0.5, layout=0).backward(g)
torch.cuda.synchronize()
print('done')
