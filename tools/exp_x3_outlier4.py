import sys, torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
from sl_hwgat_b200.models import HWGATE as M_, model_params as P_
B = 128
params = P_.HWGATEParams({'num_class': 262, 'src_len': 64}, 2, "cuda")
torch.manual_seed(1001)
model = M_.Model(*params.get_model_params()).cuda().eval()
x = torch.rand(B, 64, 64, 2, device="cuda")[102:103].contiguous()
blk = model.layers[2].blocks[0]
keep = {}
blk.norm1.register_forward_hook(lambda m, i, o: keep.__setitem__("xn", o.detach().clone()))
blk.attn.proj_drop.register_forward_hook(lambda m, i, o: keep.__setitem__("proj_out", i[0].detach().clone()))
blk.norm2.register_forward_hook(lambda m, i, o: keep.__setitem__("x1", i[0].detach().clone()))
res = {}
with torch.no_grad():
    for mode in ("ffma", "x3"):
        ops.set_fp32_mode(mode)
        model(x)
        res[mode] = dict(keep)
    for k in ("xn", "proj_out", "x1"):
        a, b = res["ffma"][k].double(), res["x3"][k].double()
        d = (a - b).abs()
        print(k, "max abs diff %.3e absmax %.3e" % (d.max(), a.abs().max()), "n elements > 1e-4:", int((d > 1e-4).sum()), "shape", tuple(a.shape))
    d = (res["ffma"]["proj_out"].double() - res["x3"]["proj_out"].double()).abs().reshape(-1, 512)
    rows = (d.amax(1) > 1e-4).nonzero().flatten().tolist()
    print("bad rows", rows[:40], "count", len(rows))
    cols = (d.amax(0) > 1e-4).nonzero().flatten().tolist()
    print("bad cols", cols[:40], "count", len(cols))
