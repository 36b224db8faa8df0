"""K10 forward in inference (no_grad: GELU-only epilogue) and in training (p = 0.1), alone, CUDA events."""
import sys, torch
sys.path.insert(0, '.')
from sl_hwgat_b200 import ops
tokens128 = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 192 * 64     # configs[1]: rows at d = 128
for d in (128, 256, 512):
    n, hid = tokens128 * 128 // d, 2 * d
    h = torch.randn(n, d, device='cuda').to(torch.bfloat16)
    w1 = (torch.randn(hid, d, device='cuda') / d ** 0.5)
    b1 = torch.zeros(hid, device='cuda')
    w2 = (torch.randn(d, hid, device='cuda') / hid ** 0.5)
    def t(fn, reps=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    def ev():
        with torch.no_grad(): ops.feed_forward_core(h, w1, b1, w2, 0.1, False)
    w1g = w1.clone().requires_grad_(True)
    def tr():
        ops.feed_forward_core(h, w1g, b1, w2, 0.1, True)
    print(f"d={d} n={n}: inference forward {t(ev):.3f} ms, training forward {t(tr):.3f} ms", flush=True)
