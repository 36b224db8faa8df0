"""localise the x3 outlier of sample 102: per-module output difference x3 vs ffma vs fp64"""
import sys, torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
from sl_hwgat_b200.models import HWGATE as M_, model_params as P_
B = 128
params = P_.HWGATEParams({'num_class': 262, 'src_len': 64}, 2, "cuda")
torch.manual_seed(1001)
model = M_.Model(*params.get_model_params()).cuda().eval()
x = torch.rand(B, 64, 64, 2, device="cuda")[102:103].contiguous()
acts = {}
def hook(name):
    def f(m, i, o):
        acts.setdefault(name, []).append((i[0].detach().clone(), o.detach().clone()))
    return f
for n_, m_ in model.named_modules():
    if isinstance(m_, (torch.nn.Linear, torch.nn.LayerNorm, M_.PartAttentionBlock, M_.MSA, M_.FeedForward)):
        m_.register_forward_hook(hook(n_))
with torch.no_grad():
    for mode in ("ffma", "x3"):
        ops.set_fp32_mode(mode)
        model(x)
for k, v in acts.items():
    if len(v) != 2: continue
    (i0, o0), (i1, o1) = v
    di = float((i0.double() - i1.double()).abs().max() / i0.double().abs().max())
    do = float((o0.double() - o1.double()).abs().max() / o0.double().abs().max())
    flag = " <<<" if do > 20 * max(di, 1e-7) else ""
    print(f"{k:40s} in {di:.2e} out {do:.2e}{flag}")
