"""CTA-pair (cta_group::2) GEMMs against their single-CTA forms and cuBLAS: correctness on small / ragged shapes,
then timing at the B=512, T=64 sizes of the model.  `python tools/prof_gemm_pair.py [nt|tn|all]`"""
import sys

import torch

sys.path.insert(0, '.')
from sl_hwgat_b200 import _lib


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


lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else 'all'

if which in ('nt', 'all'):
    for (M, N, K) in ((256, 256, 64), (384, 256, 128), (1280, 512, 512), (128 * 301, 1024, 256)):
        g = torch.Generator().manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
        Bt = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
        ref = A.float() @ Bt.float().t()
        for pair in (0, 1):
            lib.hwgat_debug_set_gemm_pair(pair)
            C = torch.full((M, N), float('nan'), dtype=torch.bfloat16, device='cuda')
            _lib.check(lib.hwgat_debug_gemm_nt(A.data_ptr(), Bt.data_ptr(), C.data_ptr(), M, N, K, st), 'gemm_nt')
            torch.cuda.synchronize()
            err = ((C.float() - ref).norm() / ref.norm()).item()
            print(f'NT M={M} N={N} K={K} pair={pair} rel {err:.2e}', flush=True)
            assert err < 3e-3, err
    N0 = 512 * 64 * 64
    for d in (256, 512):
        n = N0 * 128 // d
        for (Nn, Kk) in ((2 * d, d), (d, 2 * d), (d, d), (d, 3 * d)):
            A = torch.randn(n, Kk, device='cuda').to(torch.bfloat16)
            Bt = torch.randn(Nn, Kk, device='cuda').to(torch.bfloat16)
            C = torch.empty(n, Nn, device='cuda', dtype=torch.bfloat16)
            f2 = 2.0 * n * Nn * Kk
            t0 = timeit(lambda: torch.nn.functional.linear(A, Bt))
            ts = []
            for pair in (0, 1):
                lib.hwgat_debug_set_gemm_pair(pair)
                ts.append(timeit(lambda: lib.hwgat_debug_gemm_nt(A.data_ptr(), Bt.data_ptr(), C.data_ptr(), n, Nn, Kk, st)))
            ref = torch.nn.functional.linear(A[:512], Bt).float()
            err = ((C[:512].float() - ref).norm() / ref.norm()).item()
            print(f'd={d} NT N={Nn} K={Kk}: cuBLAS {t0:.3f} ms ({f2 / t0 / 1e9:5.0f} TF/s) single {ts[0]:.3f} ms '
                  f'({f2 / ts[0] / 1e9:5.0f}) pair {ts[1]:.3f} ms ({f2 / ts[1] / 1e9:5.0f}) relerr {err:.1e}', flush=True)
            del A, Bt, C

if which in ('tn', 'all'):
    for (M, N, Kd) in ((256, 256, 64), (768, 256, 4096), (1536, 512, 8192), (384, 256, 64 * 1001), (512, 1024, 64 * 333)):
        g = torch.Generator().manual_seed(M + N + Kd)
        A = torch.randn(Kd, M, generator=g).to(torch.bfloat16).cuda()
        B = torch.randn(Kd, N, generator=g).to(torch.bfloat16).cuda()
        ref = A.double().t() @ B.double()
        cref = A.double().sum(0)
        for pair in (0, 1):
            lib.hwgat_debug_set_gemm_pair(pair)
            C = torch.full((M, N), float('nan'), dtype=torch.float32, device='cuda')
            cs = torch.full((M,), float('nan'), dtype=torch.float32, device='cuda')
            _lib.check(lib.hwgat_debug_gemm_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), cs.data_ptr(), M, N, Kd, st), 'gemm_tn')
            torch.cuda.synchronize()
            err = ((C.double() - ref).norm() / ref.norm()).item()
            cerr = ((cs.double() - cref).abs().max() / cref.abs().max()).item()
            print(f'TN M={M} N={N} Kd={Kd} pair={pair} rel {err:.2e} colsum {cerr:.2e}', flush=True)
            if 'nocheck' not in sys.argv:
                assert err < 1e-5 and cerr < 1e-4, (err, cerr)
    N0 = 512 * 64 * 64
    import statistics
    for d in (256, 512):
        n = N0 * 128 // d
        for (Mm, Nn) in ((3 * d, d), (2 * d, d), (d, 2 * d)):
            A = torch.randn(n, Mm, device='cuda').to(torch.bfloat16)
            B = torch.randn(n, Nn, device='cuda').to(torch.bfloat16)
            C = torch.empty(Mm, Nn, device='cuda', dtype=torch.float32)
            cs = torch.empty(Mm, device='cuda', dtype=torch.float32)
            f2 = 2.0 * n * Nn * Mm

            def run(pair, csum):
                lib.hwgat_debug_set_gemm_pair(pair)
                return timeit(lambda: lib.hwgat_debug_gemm_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(),
                                                              cs.data_ptr() if csum else None, Mm, Nn, n, st), n=5)
            variants = {'cuBLAS': lambda: timeit(lambda: A.t() @ B, n=5), 'single+sum': lambda: run(0, True),
                        'single': lambda: run(0, False), 'pair+sum': lambda: run(1, True), 'pair': lambda: run(1, False)}
            res = {k: [] for k in variants}
            for _ in range(7):     # interleaved rounds: every variant sees the same thermal / power state
                for k, fn in variants.items():
                    res[k].append(fn())
            print(f'd={d} TN M={Mm} N={Nn}: ' + '  '.join(f'{k} {statistics.median(v):.3f} ms ({f2 / statistics.median(v) / 1e9:4.0f})'
                                                       for k, v in res.items()), flush=True)
            del A, B, C
print('done')
