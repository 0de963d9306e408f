"""Stand-in for the `accelerate.Accelerator` surface the reference uses
(x_transformers_rl.py:51,1106,885,981,1019,1359,1362).  Single process only.
TEST INFRASTRUCTURE.  SURVEY.md Appendix A.7."""
import torch
from torch import nn


class Accelerator:
    def __init__(self, cpu=True, **kwargs):
        self.device = torch.device('cpu')
        self.is_main_process = True
        self.num_processes = 1
        self.process_index = 0

    def prepare(self, *args):
        return args

    def backward(self, loss):
        loss.backward()

    def clip_grad_norm_(self, params, max_norm):
        return nn.utils.clip_grad_norm_(params, max_norm)

    def log(self, logs):
        self.last_logs = logs

    def wait_for_everyone(self):
        pass

    def reduce(self, t, reduction='sum'):
        return t
