import sys, torch
sys.path.insert(0, '.')
from tests._util import core_inputs, device_bits
from sl_hwgat_b200 import ops
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
d, h, F = [(128, 2, 64), (256, 4, 32), (512, 8, 16)][lvl]
B = 512
x = torch.randn(B, F, 64, d, device='cuda', dtype=torch.bfloat16)
w = torch.randn(3 * d, d, device='cuda') * 0.05
b = torch.randn(3 * d, device='cuda') * 0.05
bits = device_bits(F, 1)
for i in range(3):
    y = ops.window_graph_attention(x, w, b, bits, h, shift=1, threshold=(None if "--eval" in sys.argv else 0.05))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5):
    y = ops.window_graph_attention(x, w, b, bits, h, shift=1, threshold=(None if "--eval" in sys.argv else 0.05))
e1.record(); torch.cuda.synchronize()
print('lvl', lvl, 'ms', e0.elapsed_time(e1) / 5)
