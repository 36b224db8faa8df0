import sys, torch, numpy as np
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops, _lib
from sl_hwgat_b200.models import HWGATE as M_, model_params as P_
from tests._util import ADJ
lib = _lib.load()
B = 128
params = P_.HWGATEParams({'num_class': 262, 'src_len': 64}, 2, "cuda")
torch.manual_seed(1001)
model = M_.Model(*params.get_model_params()).cuda().eval()
x = torch.rand(B, 64, 64, 2, device="cuda")[102:103].contiguous()
blk = model.layers[2].blocks[0]
keep = {}
blk.norm1.register_forward_hook(lambda m, i, o: keep.__setitem__("xn", o.detach().clone()))
torch.set_printoptions(precision=4, linewidth=200)
with torch.no_grad():
    ops.set_fp32_mode("x3")
    model(x)
    xn = keep["xn"].contiguous()
    bits = blk._block_bits(x.device)
    w, b = blk.attn.qkv.weight.detach().contiguous(), blk.attn.qkv.bias.detach().contiguous()
    flat = xn.reshape(-1, 512)
    q64 = flat.double() @ w.double().t() + b.double()
    qs = {}
    outs = {}
    for mode in ("ffma", "x3"):
        ops.set_fp32_mode(mode)
        qs[mode] = ops.linear_f32(flat, w, b).double()
        outs[mode] = ops.window_graph_attention(xn, w, b, bits, 8, shift=0, threshold=None, layout=ops.LAYOUT_BFKD, window=16).double().reshape(-1, 512)
    f, k = 11, 36
    fi, tp, wdx, kk = f // 2, f % 2, k // 16, k % 16
    rows = [(2 * fi + t) * 64 + wdx * 16 + j for t in range(2) for j in range(16)]
    me = tp * 16 + kk
    assert rows[me] == 740
    adj = torch.from_numpy(ADJ[wdx].astype(np.float64)).cuda()
    d = (outs["x3"][740] - outs["ffma"][740]).abs().reshape(8, 64).amax(1)
    print("context diff of row 740 per head:", d.tolist())
    for h in range(8):
        if d[h] < 1e-4: continue
        for name, Q in (("fp64", q64), ("ffma", qs["ffma"]), ("x3", qs["x3"])):
            qv = Q[740, h * 64:(h + 1) * 64] * 0.125
            K = Q[rows][:, 512 + h * 64: 512 + (h + 1) * 64]
            S = K @ qv
            live = adj[me] * (S != 0)
            Sm = torch.where(live > 0, S, torch.full_like(S, -10000.0))
            P = torch.softmax(Sm, 0)
            print(f"head {h} {name}: live S", S[adj[me] > 0].tolist(), "P live", P[adj[me] > 0].tolist())
        # in fp32 arithmetic as the kernel does it
        for name in ("ffma", "x3"):
            Q = qs[name].float()
            qv = Q[740, h * 64:(h + 1) * 64] * 0.125
            K = Q[rows][:, 512 + h * 64: 512 + (h + 1) * 64]
            S = K @ qv
            print(f"head {h} {name} fp32 S all:", S.tolist())
