"""One K10 forward + backward (train mode, p = 0.1) at one level of the B=512, T=64 model, for ncu --set full."""
import sys
import torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d = [128, 256, 512][lvl]
n, hid = 512 * 64 * 64 * 128 // d, 2 * d
h = torch.randn(n, d, device='cuda').to(torch.bfloat16).requires_grad_(True)
w1 = (torch.randn(hid, d, device='cuda') / d ** 0.5).requires_grad_(True)
b1 = torch.zeros(hid, device='cuda', requires_grad=True)
w2 = (torch.randn(d, hid, device='cuda') / hid ** 0.5).requires_grad_(True)
gv = torch.randn(n, d, device='cuda').to(torch.bfloat16)
for _ in range(2):
    ops.feed_forward_core(h, w1, b1, w2, 0.1, True).backward(gv)
torch.cuda.synchronize()
print('done')
