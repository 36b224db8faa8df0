"""K10 (FeedForward on tcgen05 with fused epilogues) against the path it replaces (cuBLAS F.linear + K7 + F.linear)
at the B=512, T=64 sizes, forward and forward+backward, train mode p=0.1.  Also the bare NT GEMM against cuBLAS."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from sl_hwgat_b200 import _lib, ops


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


N = 512 * 64 * 64
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for d in (128, 256, 512):
    n = N * 128 // d
    hid = 2 * d
    h = torch.randn(n, d, device='cuda').to(torch.bfloat16).requires_grad_(True)
    w1 = (torch.randn(hid, d, device='cuda') / d ** 0.5).requires_grad_(True)
    b1 = torch.zeros(hid, device='cuda', requires_grad=True)
    w2 = (torch.randn(d, hid, device='cuda') / hid ** 0.5).requires_grad_(True)
    gv = torch.randn(n, d, device='cuda').to(torch.bfloat16)
    w1b, w2b = w1.detach().to(torch.bfloat16), w2.detach().to(torch.bfloat16)

    def old():
        u0 = torch.nn.functional.linear(h, w1b.requires_grad_(True))
        g = ops.bias_gelu_dropout(u0, b1, 0.1, True)
        return torch.nn.functional.linear(g, w2b.requires_grad_(True))

    def new():
        return ops.feed_forward_core(h, w1, b1, w2, 0.1, True)

    fl = 4.0 * n * d * hid
    for name, fn in (("cuBLAS+K7", old), ("K10", new)):
        tf = timeit(lambda: fn())
        tb = timeit(lambda: fn().backward(gv))
        print(f'd={d} {name:10s} fwd {tf:.3f} ms ({fl / tf / 1e9:6.0f} TF/s)  fwd+bwd {tb:.3f} ms ({3 * fl / tb / 1e9:6.0f} TF/s)',
              flush=True)
    # bare GEMMs: C[n, N] = A[n, K] . Bt[N, K]^T
    for (Nn, Kk) in ((hid, d), (d, hid), (d, d)):
        A = torch.randn(n, Kk, device='cuda').to(torch.bfloat16)
        Bt = torch.randn(Nn, Kk, device='cuda').to(torch.bfloat16)
        C = torch.empty(n, Nn, device='cuda', dtype=torch.bfloat16)
        t0 = timeit(lambda: torch.nn.functional.linear(A, Bt))
        t1 = timeit(lambda: lib.hwgat_debug_gemm_nt(A.data_ptr(), Bt.data_ptr(), C.data_ptr(), n, Nn, Kk, st))
        ref = torch.nn.functional.linear(A[:256], Bt).float()
        err = ((C[:256].float() - ref).norm() / ref.norm()).item()
        f2 = 2.0 * n * Nn * Kk
        print(f'd={d} NT gemm N={Nn} K={Kk}: cuBLAS {t0:.3f} ms ({f2 / t0 / 1e9:6.0f} TF/s)  hwgat {t1:.3f} ms '
              f'({f2 / t1 / 1e9:6.0f} TF/s) relerr {err:.1e}', flush=True)
    del h, w1, w2, gv
