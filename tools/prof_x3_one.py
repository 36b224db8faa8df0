"""One fp32 Linear forward + backward in x3 mode (gemm_nt_x3 / gemm_tn_x3 / split3) at a level of the batch-128 model,
and one inference FeedForward (K10f), for ncu --set full."""
import sys, torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
what = sys.argv[1] if len(sys.argv) > 1 else "x3"
lvl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = [128, 256, 512][lvl]
if what == "x3":
    ops.set_fp32_mode("x3")
    n = 128 * 64 * 64 * 128 // d
    x = torch.randn(n, d, device='cuda', requires_grad=True)
    w = (torch.randn(3 * d, d, device='cuda') / d ** 0.5).requires_grad_(True)
    b = torch.zeros(3 * d, device='cuda', requires_grad=True)
    g = torch.randn(n, 3 * d, device='cuda')
    for _ in range(2):
        ops.linear_f32(x, w, b).backward(g)
else:
    n = 256 * 192 * 64 * 128 // d
    h = torch.randn(n, d, device='cuda').to(torch.bfloat16)
    w1 = torch.randn(2 * d, d, device='cuda') / d ** 0.5
    b1 = torch.zeros(2 * d, device='cuda')
    w2 = torch.randn(d, 2 * d, device='cuda') / (2 * d) ** 0.5
    with torch.no_grad():
        for _ in range(3):
            ops.feed_forward_core(h, w1, b1, w2, 0.0, False)
torch.cuda.synchronize()
print('done')
