"""Other BASELINE.json configs, for DESIGN.md: configs[1] (inference, B=256, T=192, 2002 classes, bf16) and the
per-GPU share of configs[4] (T=256, 128 sequences per GPU, training).  Device-timed, inputs resident."""
import sys, json, torch
sys.path.insert(0, '.')
from sl_hwgat_b200.models import HWGATE, model_params
from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss

def build(T, classes):
    p = model_params.HWGATEParams({"num_class": classes, "src_len": T}, 2, "cuda")
    torch.manual_seed(1001)
    return HWGATE.Model(*p.get_model_params()).cuda()

def timeit(fn, warm=3, n=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = {}
m = build(192, 2002).eval()
x = torch.rand(256, 192, 64, 2, device="cuda")
def infer():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return m(x)
ms = timeit(infer)
out["configs[1] inference B=256 T=192 bf16"] = {"ms": ms, "seq_per_s": 256 / ms * 1e3}
def infer32():
    with torch.no_grad():
        return m(x[:64])
ms = timeit(infer32, 1, 2)
out["inference B=64 T=192 fp32 parity kernels"] = {"ms": ms, "seq_per_s": 64 / ms * 1e3}
del m, x
torch.cuda.empty_cache()
m = build(256, 262).train()
crit = SmoothedCrossEntropyLoss()
x = torch.rand(128, 256, 64, 2, device="cuda"); y = torch.randint(0, 262, (128,), device="cuda")
def train():
    m.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l = crit(m(x), y)
    l.backward()
ms = timeit(train)
out["configs[4] per-GPU share: train B=128 T=256 bf16"] = {"ms": ms, "seq_per_s": 128 / ms * 1e3,
                                                           "max_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
print(json.dumps(out, indent=1))
