"""Stand-in for `assoc-scan` (x_transformers_rl.py:49,634-636).  TEST INFRASTRUCTURE.
h_t = g_t * h_{t-1} + x_t over dim 1, h_{-1} = 0; reverse=True runs right-to-left.
The package's non-accelerated path is a log-depth associative scan; this shim evaluates the
same recurrence sequentially in the input dtype (differences are fp32 reassociation ~1e-7).
SURVEY.md Appendix A.2.  Parity unpinned (third-party)."""
import torch
from torch import nn


class AssocScan(nn.Module):
    def __init__(self, reverse=False, use_accelerated=False):
        super().__init__()
        self.reverse = reverse
        if use_accelerated:
            raise RuntimeError('assoc_scan shim: the accelerated (Triton) backend is not available')

    def forward(self, gates, inputs, prev=None):
        if self.reverse:
            gates, inputs = gates.flip(1), inputs.flip(1)
        h = torch.zeros_like(inputs[:, 0]) if prev is None else prev
        out = []
        for t in range(inputs.shape[1]):
            h = gates[:, t] * h + inputs[:, t]
            out.append(h)
        out = torch.stack(out, dim=1)
        return out.flip(1) if self.reverse else out
