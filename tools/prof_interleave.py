"""Experiment: two half batches on two CUDA streams (HBM-bound kernels of one overlapping tensor-bound kernels of the
other) against one full batch on one stream.  Same thresholds for both halves (one scalar per MSA call for the batch)."""
import sys, time, torch
sys.path.insert(0, '.')
import bench
from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss
bench.set_config("train512")
dev = torch.device("cuda", 0)
model = bench.build_model(dev).train()
crit = SmoothedCrossEntropyLoss()
x, y = bench.synthetic_batch(512)
x, y = x.to(dev), y.to(dev)
n_thr = 8


class Thr:
    def __init__(self): self.vals, self.i = [], 0
    def draw(self, real):
        def f(*a, **k):
            if self.i >= len(self.vals): self.vals.append(real(1))
            v = self.vals[self.i]; self.i += 1
            return v
        return f


def full():
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = crit(model(x), y)
    loss.backward()


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def halves():
    model.zero_grad(set_to_none=True)
    real = torch.rand
    thr = Thr()
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    losses = []
    for st, sl in ((s1, slice(0, 256)), (s2, slice(256, 512))):
        thr.i = 0
        torch.rand = thr.draw(real)
        try:
            with torch.cuda.stream(st), torch.autocast("cuda", dtype=torch.bfloat16):
                losses.append(crit(model(x[sl]), y[sl]) * 0.5)
        finally:
            torch.rand = real
    for st, l in zip((s1, s2), losses):
        with torch.cuda.stream(st):
            l.backward()
    cur.wait_stream(s1); cur.wait_stream(s2)


def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for r in range(2):
    print("full batch, one stream   ms/step", round(timeit(full), 2), flush=True)
    print("two halves, two streams  ms/step", round(timeit(halves), 2), flush=True)
