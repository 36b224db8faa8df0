import sys, torch
sys.path.insert(0, '.')
from tests._util import device_bits
from sl_hwgat_b200 import ops
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d, h, F = [(128, 2, 64), (256, 4, 32), (512, 8, 16)][lvl]
B = 512
x = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16).requires_grad_(True)
w = (torch.randn(3 * d, d, device='cuda') * 0.05).requires_grad_(True)
b = (torch.randn(3 * d, device='cuda') * 0.05).requires_grad_(True)
g = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16)
bits = device_bits(F, 1)
for i in range(2):
    y = ops.window_graph_attention(x, w, b, bits, h, shift=1, threshold=0.05)
    y.backward(g)
torch.cuda.synchronize()
print('done')
