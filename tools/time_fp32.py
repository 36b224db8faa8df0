"""The model WITHOUT autocast (the reference's own loop, utils.py:102): train fwd+bwd and eval forward, fp32 modes
"ffma" (true-fp32 parity kernels + cuBLAS sgemm) and "x3" (tcgen05, six bf16 products), against eager PyTorch fp32."""
import sys, time, torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
from sl_hwgat_b200.models import HWGATE as M_, model_params as P_
from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
params = P_.HWGATEParams({'num_class': 262, 'src_len': 64}, 2, "cuda")
torch.manual_seed(1001)
model = M_.Model(*params.get_model_params()).cuda()
x = torch.rand(B, 64, 64, 2, device="cuda")
tgt = torch.randint(0, 262, (B,), device="cuda")
crit = SmoothedCrossEntropyLoss()


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def train_step():
    model.zero_grad(set_to_none=True)
    crit(model(x), tgt).backward()


def eval_step():
    with torch.no_grad():
        model(x)


outs = {}
for mode in ("ffma", "x3"):
    ops.set_fp32_mode(mode)
    model.train()
    t = timed(train_step)
    model.eval()
    e = timed(eval_step)
    with torch.no_grad():
        outs[mode] = model(x).double()
    print(f"fp32 mode {mode}: train fwd+bwd batch {B}: {t:.1f} ms = {B / t * 1e3:.0f} sequences/s; "
          f"eval forward {e:.1f} ms = {B / e * 1e3:.0f} sequences/s", flush=True)
d = (outs["x3"] - outs["ffma"]).abs().max() / outs["ffma"].abs().max()
print(f"eval logits x3 vs ffma: max rel {d:.2e}")
# both against the fp64 oracle (eager, on the GPU) on the first samples, and eager fp32 PyTorch beside them
from oracle import hwgate_oracle as O
cfg = O.HWGATEConfig(temporal_dim=64, num_classes=262)
nb = min(B, 16)
sd = {k: v.detach() for k, v in model.state_dict().items()}
with torch.no_grad():
    ref = O.model_forward(x[:nb].double(), {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, cfg)
    eager = O.model_forward(x[:nb], sd, cfg).double()
rel = lambda a: float((a - ref).abs().max() / ref.abs().max())
print(f"eval logits against the fp64 oracle (first {nb} samples): ffma {rel(outs['ffma'][:nb]):.2e}, "
      f"x3 {rel(outs['x3'][:nb]):.2e}, eager PyTorch fp32 {rel(eager):.2e}")
sdf = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}


def eager_train():
    for v in sdf.values():
        v.grad = None
    thr = [0.5] * 8
    out = O.model_forward(x, sdf, cfg, thresholds=thr, drop=0.1)
    crit(out, tgt).backward()


def eager_eval():
    with torch.no_grad():
        O.model_forward(x, sd, cfg)


try:
    t, e = timed(eager_train), timed(eager_eval)
    print(f"eager PyTorch fp32 (oracle op sequence on CUDA): train {t:.1f} ms = {B / t * 1e3:.0f} sequences/s; "
          f"eval {e:.1f} ms = {B / e * 1e3:.0f} sequences/s")
except Exception as ex:      # noqa: BLE001
    print("eager timing failed:", repr(ex)[:300])
from torch.profiler import profile, ProfilerActivity
model.train()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    train_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
