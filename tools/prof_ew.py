"""Time the bandwidth-bound kernels (K4-K7) alone at the B=512, T=64 sizes: achieved GB/s of algorithmic bytes."""
import sys, torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

N = 512 * 64 * 64   # tokens at level 0
for d in (128, 256, 512):
    n = N * 128 // d
    x = torch.randn(n, d, device='cuda').requires_grad_(True)
    gm = torch.ones(d, device='cuda', requires_grad=True); bt = torch.zeros(d, device='cuda', requires_grad=True)
    a = torch.randn(n, d, device='cuda').to(torch.bfloat16).requires_grad_(True)
    u = torch.randn(n, 2 * d, device='cuda').to(torch.bfloat16).requires_grad_(True)
    el = n * d
    t = timeit(lambda: ops.layer_norm_residual(x, gm, bt)); print(f'd={d} ln_fwd        {t:.3f} ms {6*el/t/1e6:8.0f} GB/s')
    xa, y = ops.layer_norm_residual(x, gm, bt); gy = torch.randn_like(y); gr = torch.randn_like(x)
    def lnb():
        torch.autograd.grad((xa, y), (x,), (gr, gy), retain_graph=True)
    t = timeit(lnb); print(f'd={d} ln_bwd        {t:.3f} ms {14*el/t/1e6:8.0f} GB/s')
    for p in (0.0, 0.1):
        t = timeit(lambda: ops.dropout_add(x, a, p, True)); print(f'd={d} dropadd_fwd p={p} {t:.3f} ms {10*el/t/1e6:8.0f} GB/s')
        o = ops.dropout_add(x, a, p, True); g = torch.randn_like(o)
        t = timeit(lambda: torch.autograd.grad(o, (a,), (g,), retain_graph=True)); print(f'd={d} dropadd_bwd p={p} {t:.3f} ms {6*el/t/1e6:8.0f} GB/s')
        t = timeit(lambda: ops.gelu_dropout(u, p, True)); print(f'd={d} gelu_fwd p={p}    {t:.3f} ms {8*el/t/1e6:8.0f} GB/s')
        o = ops.gelu_dropout(u, p, True); g = torch.randn_like(o)
        t = timeit(lambda: torch.autograd.grad(o, (u,), (g,), retain_graph=True)); print(f'd={d} gelu_bwd p={p}    {t:.3f} ms {12*el/t/1e6:8.0f} GB/s')
    del x, a, u
