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
blk.norm1.register_forward_hook(lambda m, i, o: keep.setdefault("xn", o.detach().clone()))
with torch.no_grad():
    ops.set_fp32_mode("ffma")
    model(x)
    xn = keep["xn"]
    flat = xn.reshape(-1, xn.shape[-1])
    for name, lin in (("qkv", blk.attn.qkv), ("proj", blk.attn.proj), ("fc1", blk.ff.fc1)):
        ref = flat.double() @ lin.weight.double().t() + lin.bias.double()
        for mode in ("ffma", "x3"):
            ops.set_fp32_mode(mode)
            y = ops.linear_f32(flat, lin.weight, lin.bias).double()
            err = (y - ref).abs()
            idx = err.argmax()
            r, c = int(idx // ref.shape[1]), int(idx % ref.shape[1])
            print(f"{name} {mode}: max abs err {err.max():.3e} at row {r} col {c} (ref {ref[r, c]:.4e}, got {y[r, c]:.4e}); ref absmax {ref.abs().max():.3e}; rows with err > 1e-4*absmax: {int((err.amax(1) > 1e-4 * ref.abs().max()).sum())}")
    # attention context in both modes against each other
    bits = blk._block_bits(xn.device)
    outs = {}
    for mode in ("ffma", "x3"):
        ops.set_fp32_mode(mode)
        outs[mode] = ops.window_graph_attention(xn, blk.attn.qkv.weight, blk.attn.qkv.bias, bits, blk.attn.num_heads, shift=0, threshold=None, layout=ops.LAYOUT_BFKD, window=16).double()
    d = (outs["x3"] - outs["ffma"]).abs()
    idx = d.argmax()
    print("attention context x3 vs ffma: max abs", float(d.max()), "absmax", float(outs["ffma"].abs().max()), "at flat index", int(idx), "token", int(idx) // 512, "col", int(idx) % 512,
          "tokens with diff > 1e-4:", int((d.reshape(-1, 512).amax(1) > 1e-4 * outs["ffma"].abs().max()).sum()))
    # logits of the worst token/head: look for an exactly-zero or near-tie pattern
    tok = int(idx) // 512
    qkv = (flat.double() @ blk.attn.qkv.weight.double().t() + blk.attn.qkv.bias.double())
    print("worst token row of q (first 8):", qkv[tok, :8].tolist())
