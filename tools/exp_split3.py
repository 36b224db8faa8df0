"""Experiment: fp32-accurate products from bf16 tcgen05 MMAs (hi / mid / lo split, 3 or 6 partial products), through the
existing TN GEMM (fp32 output), against fp64; and where the time of the fp32 parity mode goes today."""
import sys, time, torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import _lib

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def split3(x):
    hi = x.to(torch.bfloat16)
    r = x - hi.float()
    mid = r.to(torch.bfloat16)
    lo = (r - mid.float()).to(torch.bfloat16)
    return hi, mid, lo


def gemm_tn(A, B):
    Kd, M = A.shape
    N = B.shape[1]
    C = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    _lib.check(lib.hwgat_debug_gemm_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), None, M, N, Kd, st), "tn")
    return C


def rel(a, b):
    return float((a.double() - b).norm() / b.norm()), float((a.double() - b).abs().max() / b.abs().max())


for Kd, M, N in [(64 * 8, 256, 256), (64 * 128, 256, 256), (64 * 2048, 384, 128)]:
    g = torch.Generator().manual_seed(Kd)
    x = torch.randn(Kd, M, generator=g).cuda()
    y = (torch.randn(Kd, N, generator=g) + 0.5).cuda()        # a non-zero mean makes the sums grow
    ref = x.double().t() @ y.double()
    xh, xm, xl = split3(x)
    yh, ym, yl = split3(y)
    c1 = gemm_tn(xh, yh)
    c3 = gemm_tn(torch.cat([xm, xh, xh]), torch.cat([yh, ym, yh]))
    c6 = gemm_tn(torch.cat([xl, xh, xm, xm, xh, xh]), torch.cat([yh, yl, ym, yh, ym, yh]))
    # small terms in a separate accumulation, added in fp32 afterwards
    c6s = gemm_tn(torch.cat([xl, xh, xm, xm, xh]), torch.cat([yh, yl, ym, yh, ym])) + c1
    f32 = x.t() @ y
    torch.cuda.synchronize()
    print(f"Kd={Kd} M={M} N={N}: 1 product l2/max {rel(c1, ref)}, 3 products {rel(c3, ref)}, 6 products {rel(c6, ref)}, "
          f"5 + 1 separately {rel(c6s, ref)}, torch fp32 matmul {rel(f32, ref)}", flush=True)

# ---- the fp32 parity mode today: time per step and the top kernels
import numpy as np
from oracle import hwgate_oracle as O
from sl_hwgat_b200.models import HWGATE as M_, model_params as P_
params = P_.HWGATEParams({'num_class': 262, 'src_len': 64}, 2, "cuda")
torch.manual_seed(1001)
model = M_.Model(*params.get_model_params()).cuda().train()
Bsz = 64
x = torch.rand(Bsz, 64, 64, 2, device="cuda")
tgt = torch.randint(0, 262, (Bsz,), device="cuda")
from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss as L_
crit = L_()


def step():
    model.zero_grad(set_to_none=True)
    loss = crit(model(x), tgt)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(3):
    step()
torch.cuda.synchronize()
dt = (time.time() - t0) / 3
print(f"fp32 parity mode, train fwd+bwd, batch {Bsz}: {dt * 1e3:.1f} ms/step = {Bsz / dt:.0f} sequences/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
