"""Label-smoothed cross entropy with the reference's class name and semantics
(hwgat/losses/SmoothCrossEntropy.py:15-39, smooth_factor 0.01): the criterion
the fwd+bwd metric is quoted with.  Plain PyTorch - it is not on the hot path."""
import torch
import torch.nn as nn


class SmoothedCrossEntropyLoss(nn.Module):
    def __init__(self, smooth_factor=0.01):
        super().__init__()
        self.smooth_factor = smooth_factor

    def forward(self, input, target):
        logp = torch.log_softmax(input.float(), dim=-1)
        nll = -logp.gather(-1, target.unsqueeze(-1)).squeeze(-1)
        uniform = -logp.mean(dim=-1)
        return ((1.0 - self.smooth_factor) * nll + self.smooth_factor * uniform).mean()
