"""Label-smoothed cross entropy with the reference's class name and semantics
(hwgat/losses/SmoothCrossEntropy.py:15-39, smooth_factor 0.01): the criterion
the fwd+bwd metric is quoted with.  It is kernel K14
(sl_hwgat_b200.ops.smooth_cross_entropy: one pass per row, deterministic mean,
fused backward).  Like the rest of the package it has no CPU path: logits that
are not on a CUDA device raise."""
import torch.nn as nn

from sl_hwgat_b200 import ops


class SmoothedCrossEntropyLoss(nn.Module):
    def __init__(self, smooth_factor=0.01):
        super().__init__()
        self.smooth_factor = smooth_factor

    def forward(self, input, target):
        # (rows, classes) logits, (rows,) class indices; anything else is flattened to that like log_softmax(dim=-1)
        if input.dim() != 2:
            input, target = input.reshape(-1, input.shape[-1]), target.reshape(-1)
        return ops.smooth_cross_entropy(input, target, self.smooth_factor)
