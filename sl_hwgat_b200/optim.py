"""AdamW with the reference's optimizer interface on one multi-tensor kernel (K11).

The reference builds `torch.optim.AdamW(model.parameters(), lr=cfg.lr)` (hwgat/utils.py:73-75) and calls
`optimizer.zero_grad()` / `optimizer.step()` once per batch (utils.py:105-107); checkpoints store
`optimizer.state_dict()` (utils.py:164-176).  `AdamW` below is a `torch.optim.Optimizer` with the same constructor
defaults and the same state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter), so `state_dict()` /
`load_state_dict()` interoperate with torch's AdamW and `CosineAnnealingLR` (utils.py:86-88) drives it unchanged.
`step()` is ONE launch of `hwgat_adamw_step` per parameter group instead of PyTorch's per-tensor kernels.
There is no CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is not supported by the fused kernel (the reference does not use it)")
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"Invalid betas: {betas}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """One AdamW update.  `grad_scale` multiplies every gradient inside the kernel (e.g. 1/world_size after a
        summed all-reduce)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            ps, gs, ms, vs = [], [], [], []
            step = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise _lib.HwgatError("sl_hwgat_b200.optim.AdamW updates CUDA parameters only (no CPU fallback)")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise _lib.HwgatError("the AdamW kernel takes dense float32 parameters and gradients")
                if not p.is_contiguous():
                    raise _lib.HwgatError("the AdamW kernel takes contiguous parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                t = int(st["step"].item())
                if step is None:
                    step = t
                elif step != t:          # parameters that joined later: their own launch below
                    step = -1
                ps.append(p); gs.append(p.grad.contiguous()); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            if not ps:
                continue
            b1, b2 = group["betas"]
            batches = [(ps, gs, ms, vs, step)] if step != -1 else \
                [([p], [g], [m], [v], int(self.state[p]["step"].item())) for p, g, m, v in zip(ps, gs, ms, vs)]
            for bp, bg, bm, bv, t in batches:
                n = len(bp)
                arr = ctypes.c_void_p * n
                sizes = (ctypes.c_longlong * n)(*[p.numel() for p in bp])
                dev = bp[0].device
                with torch.cuda.device(dev):
                    check(lib.hwgat_adamw_step(n, arr(*[p.data_ptr() for p in bp]), arr(*[g.data_ptr() for g in bg]),
                                               arr(*[m.data_ptr() for m in bm]), arr(*[v.data_ptr() for v in bv]),
                                               sizes, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                               float(group["weight_decay"]), t, float(grad_scale),
                                               torch.cuda.current_stream().cuda_stream), "hwgat_adamw_step")
        # the kernel wrote the parameters through raw pointers: torch's version counters did not move, so the cached
        # bf16 copies the ops keep per parameter version (ops.cast_cached) must be dropped explicitly
        from . import ops
        ops.invalidate_cast_cache()
        return loss
