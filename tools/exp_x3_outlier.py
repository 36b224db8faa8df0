"""per-sample error of the fp32 modes against the fp64 oracle (looking for outliers)"""
import sys, torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
from sl_hwgat_b200.models import HWGATE as M_, model_params as P_
from oracle import hwgate_oracle as O
B = 128
params = P_.HWGATEParams({'num_class': 262, 'src_len': 64}, 2, "cuda")
torch.manual_seed(1001)
model = M_.Model(*params.get_model_params()).cuda().eval()
x = torch.rand(B, 64, 64, 2, device="cuda")
cfg = O.HWGATEConfig(temporal_dim=64, num_classes=262)
sd = {k: v.detach() for k, v in model.state_dict().items()}
with torch.no_grad():
    ref = torch.cat([O.model_forward(x[i:i + 16].double(), {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, cfg) for i in range(0, B, 16)])
    eager = O.model_forward(x, sd, cfg).double()
    outs = {}
    for mode in ("ffma", "x3", "x3"):
        ops.set_fp32_mode(mode)
        o = model(x).double()
        if mode in outs:
            print("x3 run-to-run identical:", bool((o == outs[mode]).all()))
        outs[mode] = o
scale = ref.abs().max()
for name, o in [("ffma", outs["ffma"]), ("x3", outs["x3"]), ("eager", eager)]:
    e = (o - ref).abs().amax(1) / scale
    print(name, "per-sample max rel: max %.2e median %.2e" % (e.max(), e.median()), "worst samples", e.topk(5).indices.tolist(), ["%.1e" % v for v in e.topk(5).values.tolist()])
