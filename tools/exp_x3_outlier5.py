import sys, torch, numpy as np
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
from sl_hwgat_b200.models import HWGATE as M_, model_params as P_
from oracle import hwgate_oracle as O
from tests._util import ADJ
B = 128
params = P_.HWGATEParams({'num_class': 262, 'src_len': 64}, 2, "cuda")
torch.manual_seed(1001)
model = M_.Model(*params.get_model_params()).cuda().eval()
x = torch.rand(B, 64, 64, 2, device="cuda")[102:103].contiguous()
blk = model.layers[2].blocks[0]
keep = {}
blk.norm1.register_forward_hook(lambda m, i, o: keep.__setitem__("xn", o.detach().clone()))
res = {}
with torch.no_grad():
    for mode in ("ffma", "x3"):
        ops.set_fp32_mode(mode)
        model(x)
        res[mode] = keep["xn"]
    bits = blk._block_bits(x.device)
    w, b = blk.attn.qkv.weight, blk.attn.qkv.bias
    mask = O.combined_mask(ADJ, 16, 16, 2, 0)
    for src in ("ffma", "x3"):
        xn = res[src]
        ref = O.attention_core(xn.double().cpu(), w.double().cpu(), b.double().cpu(), 8, mask, 16, 2, 0, None).cuda()
        flat = xn.reshape(-1, 512)
        qref = flat.double() @ w.double().t() + b.double()
        for mode in ("ffma", "x3"):
            ops.set_fp32_mode(mode)
            q = ops.linear_f32(flat, w, b).double()
            e = (q - qref).abs()
            c = ops.window_graph_attention(xn, w, b, bits, 8, shift=0, threshold=None, layout=ops.LAYOUT_BFKD, window=16).double()
            ec = (c - ref).abs().reshape(-1, 512)
            print(f"xn from {src}, mode {mode}: qkv max err {e.max():.2e} (row {int(e.amax(1).argmax())}); context max err {ec.max():.2e} at row {int(ec.amax(1).argmax())}, rows > 1e-4: {(ec.amax(1) > 1e-4).nonzero().flatten().tolist()}")
    # the logits of row 740, head by head, in fp64: any exact zero / tie?
    xn = res["x3"].double().reshape(-1, 512)
    qkv = xn @ w.double().t() + b.double()
    r = 740
    win0 = (r // 32) * 32     # only valid for the BFKD layout if rows are window-major; print the q row norm instead
    print("row 740 q/k/v abs max:", float(qkv[r, :512].abs().max()), float(qkv[r, 512:1024].abs().max()), float(qkv[r, 1024:].abs().max()))
