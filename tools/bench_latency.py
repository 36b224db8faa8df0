"""Batch-1 (and small-batch) inference latency, eval mode, bf16 autocast: eager launches vs CUDA-graph replay
(sl_hwgat_b200.runtime.GraphedInference) - the per-sample loop of inference.py:88-95."""
import json, sys, time, torch
sys.path.insert(0, '.')
from sl_hwgat_b200.models import HWGATE, model_params
from sl_hwgat_b200.runtime import GraphedInference

T, classes = 64, 262
p = model_params.HWGATEParams({"num_class": classes, "src_len": T}, 2, "cuda")
torch.manual_seed(1001)
m = HWGATE.Model(*p.get_model_params()).cuda().eval()
fast = GraphedInference(m)
out = {}
for B in (1, 8, 64):
    x = torch.rand(B, T, 64, 2, device="cuda")
    def eager():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return m(x)
    res = {}
    for name, fn in (("eager", eager), ("graph", lambda: fast(x))):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 50
        for _ in range(n):
            fn().argmax(-1).cpu()          # the evaluator reads the prediction back every sample
        res[name + "_ms"] = (time.perf_counter() - t0) / n * 1e3
    res["speedup"] = res["eager_ms"] / res["graph_ms"]
    out[f"B={B}"] = res
print(json.dumps(out, indent=1))
