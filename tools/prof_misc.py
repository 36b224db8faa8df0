"""K1 (adjacency / mask bits), K6 with the merge folded in, K5' with the un-merge folded in, K4: alone, for ncu dram__bytes."""
import sys, torch
sys.path.insert(0, '.')
import numpy as np
from oracle import hwgate_oracle as O
from sl_hwgat_b200 import ops
B, F, d = 512, 64, 128
adj = ops.adjacency_build(O.HWGATEConfig().edges, 16, 2, "cuda")
res = torch.randn(B, F, 64, d, device="cuda").requires_grad_(True)
a0 = torch.randn(B, F, 64, d, device="cuda", dtype=torch.bfloat16).requires_grad_(True)
bias = torch.zeros(d, device="cuda", requires_grad=True)
norm = torch.nn.LayerNorm(2 * d).cuda()
for i in range(3):
    bits = ops.mask_build(adj, 64, 1)
    xm, y = ops.bias_dropout_add_merge_ln(res, a0, bias, norm, 0.1, True)
    torch.autograd.backward([xm, y], [torch.ones_like(xm), torch.ones_like(y)])
    z = ops.temporal_merge(res.detach())
torch.cuda.synchronize()
print("done")
