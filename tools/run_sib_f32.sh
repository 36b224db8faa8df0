#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_band.py tests/test_gpu_hgate.py -x -q -m gpu > gpurun_out/r02ai_band_tests.log 2>&1; tail -5 gpurun_out/r02ai_band_tests.log
python - <<'P'
import sys, time, torch, importlib
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
from sl_hwgat_b200.models import model_params as P_
from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss
for name, B, K in (("WGATE", 64, 64), ("GATE", 64, 29), ("HGATE", 128, 29)):
    mod = importlib.import_module("sl_hwgat_b200.models." + name)
    params = getattr(P_, name + "Params")({'num_class': 262, 'src_len': 64}, 2, "cuda")
    torch.manual_seed(1001)
    model = mod.Model(*params.get_model_params()).cuda().train()
    x = torch.rand(B, 64, K, 2, device="cuda"); tgt = torch.randint(0, 262, (B,), device="cuda"); crit = SmoothedCrossEntropyLoss()
    def step():
        model.zero_grad(set_to_none=True); crit(model(x), tgt).backward()
    for mode in ("ffma", "x3"):
        ops.set_fp32_mode(mode)
        for _ in range(2): step()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): step()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 3
        print(f"{name} fp32 mode {mode}: train fwd+bwd batch {B}: {t:.1f} ms = {B / t * 1e3:.0f} sequences/s", flush=True)
P
