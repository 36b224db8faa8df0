import sys, torch, numpy as np
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops, _lib
from sl_hwgat_b200.models import HWGATE as M_, model_params as P_
lib = _lib.load()
B = 128
params = P_.HWGATEParams({'num_class': 262, 'src_len': 64}, 2, "cuda")
torch.manual_seed(1001)
model = M_.Model(*params.get_model_params()).cuda().eval()
x = torch.rand(B, 64, 64, 2, device="cuda")[102:103].contiguous()
blk = model.layers[2].blocks[0]
keep = {}
blk.norm1.register_forward_hook(lambda m, i, o: keep.__setitem__("xn", o.detach().clone()))
with torch.no_grad():
    ops.set_fp32_mode("x3")
    model(x)
    xn = keep["xn"].contiguous()
    bits = blk._block_bits(x.device)
    w, b = blk.attn.qkv.weight.detach().contiguous(), blk.attn.qkv.bias.detach().contiguous()
    flat = xn.reshape(-1, 512)
    qref = flat.double() @ w.double().t() + b.double()
    st = torch.cuda.current_stream().cuda_stream
    for trial in range(3):
        ws_bytes = lib.hwgat_attn_workspace_bytes(1, 16, 64, 512, 8, 0, 0)
        ws = torch.full((ws_bytes // 4,), float("nan"), dtype=torch.float32, device="cuda")
        out = torch.empty_like(xn)
        _lib.check(lib.hwgat_attn_fwd(xn.data_ptr(), w.data_ptr(), b.data_ptr(), bits.data_ptr(), -1.0, out.data_ptr(), ws.data_ptr(), ws_bytes, 1, 16, 64, 512, 8, 16, 2, 0, 0, 0, st), "fwd")
        torch.cuda.synchronize()
        q = ws[:1024 * 1536].reshape(1024, 1536).double()
        e = (q - qref).abs()
        bad = (e > 1e-4).nonzero()
        print(f"trial {trial}: ws floats {ws.numel()}, qkv in workspace max err {e.max():.3e}, nan {int(torch.isnan(q).sum())}, bad entries {bad.shape[0]}", bad[:6].tolist(), [ (float(q[r, c]), float(qref[r, c])) for r, c in bad[:4].tolist()])
        # the S row of token 740 per head, from the workspace q / k
        r = 740
    # logits of row 740 in fp64 for each head: smallest |S| among live keys
    win = r // 32
    print("geometry note: BFKD layout, rows are tokens (b, f, k); window rows are not contiguous")
    mask = None
