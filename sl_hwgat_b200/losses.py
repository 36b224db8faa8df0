"""Label-smoothed cross entropy with the reference's class name and semantics
(hwgat/losses/SmoothCrossEntropy.py:15-39, smooth_factor 0.01): the criterion
the fwd+bwd metric is quoted with.  On CUDA tensors it is kernel K14
(sl_hwgat_b200.ops.smooth_cross_entropy: one pass per row, deterministic mean,
fused backward); CPU tensors - the reference evaluates its loss wherever its
tensors live - take the same four-line formula in PyTorch."""
import torch
import torch.nn as nn

from sl_hwgat_b200 import ops


class SmoothedCrossEntropyLoss(nn.Module):
    def __init__(self, smooth_factor=0.01):
        super().__init__()
        self.smooth_factor = smooth_factor

    def forward(self, input, target):
        if input.is_cuda and input.dim() == 2:
            return ops.smooth_cross_entropy(input, target, self.smooth_factor)
        logp = torch.log_softmax(input.float(), dim=-1)
        nll = -logp.gather(-1, target.unsqueeze(-1)).squeeze(-1)
        uniform = -logp.mean(dim=-1)
        return ((1.0 - self.smooth_factor) * nll + self.smooth_factor * uniform).mean()
