"""Stand-in for `hl-gauss-pytorch` HLGaussLoss (x_transformers_rl.py:42,356-361,427,454-462,
843).  TEST INFRASTRUCTURE.  SURVEY.md Appendix A.3; parity unpinned (third-party).
Decisions: sigma = 2 * bin width (U2); the loss is returned PER ELEMENT (U3) because the
reference consumes it element-wise (x_transformers_rl.py:470-475, 973)."""
from math import sqrt

import torch
import torch.nn.functional as F
from torch import nn


class HLGaussLoss(nn.Module):
    def __init__(self, min_value, max_value, num_bins, sigma=None, sigma_to_bin_ratio=2.,
                 eps=1e-10, clamp_to_range=False):
        super().__init__()
        self.eps = eps
        support = torch.linspace(min_value, max_value, num_bins + 1).float()
        bin_size = (support[1] - support[0]).item()
        self.sigma = sigma if sigma is not None else sigma_to_bin_ratio * bin_size
        self.min_value, self.max_value, self.clamp_to_range = min_value, max_value, clamp_to_range
        self.register_buffer('support', support, persistent=False)
        self.register_buffer('centers', (support[:-1] + support[1:]) / 2, persistent=False)
        self.sigma_times_sqrt_two = sqrt(2.) * self.sigma

    def transform_to_probs(self, target):
        if self.clamp_to_range:
            target = target.clamp(self.min_value, self.max_value)
        cdf = torch.special.erf((self.support - target[..., None]) / self.sigma_times_sqrt_two)
        z = cdf[..., -1] - cdf[..., 0]
        return (cdf[..., 1:] - cdf[..., :-1]) / z.clamp(min=self.eps)[..., None]

    def transform_from_probs(self, probs):
        return (probs * self.centers).sum(dim=-1)

    def forward(self, logits, target=None, reduction='none'):
        if target is None:
            return self.transform_from_probs(logits.softmax(dim=-1))
        probs = self.transform_to_probs(target)
        loss = -(probs * logits.log_softmax(dim=-1)).sum(dim=-1)
        if reduction == 'mean':
            return loss.mean()
        return loss
