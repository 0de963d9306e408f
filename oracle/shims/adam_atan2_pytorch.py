"""Stand-in for `adam-atan2-pytorch` AdoptAtan2 (x_transformers_rl.py:40,749).
TEST INFRASTRUCTURE.  SURVEY.md Appendix A.4 (U4); parity unpinned (third-party)."""
import torch
from torch.optim.optimizer import Optimizer


class AdoptAtan2(Optimizer):
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.99), weight_decay=0., regen_reg_rate=0.,
                 decoupled_wd=True, cautious_factor=1., a=1.27, b=1.):
        self._init_lr = lr
        self.decoupled_wd = decoupled_wd
        super().__init__(params, dict(lr=lr, betas=betas, a=a, b=b, weight_decay=weight_decay,
                                      regen_reg_rate=regen_reg_rate, cautious_factor=cautious_factor))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                grad, lr, wd, regen = p.grad, group['lr'], group['weight_decay'], group['regen_reg_rate']
                cautious, (beta1, beta2), a, b = group['cautious_factor'], group['betas'], group['a'], group['b']
                state, init_lr = self.state[p], self._init_lr
                if self.decoupled_wd:
                    wd /= init_lr
                if wd > 0.:
                    p.mul_(1. - lr * wd)
                if regen > 0. and 'param_init' in state:
                    p.lerp_(state['param_init'], lr / init_lr * regen)
                if len(state) == 0:
                    state['steps'] = 0
                    state['m'] = torch.zeros_like(grad)
                    state['v'] = grad * grad
                    if regen > 0.:
                        state['param_init'] = p.clone()
                m, v, steps = state['m'], state['v'], state['steps']
                if steps == 0:           # ADOPT: the first step only seeds v
                    state['steps'] += 1
                    continue
                steps += 1
                next_m = grad.atan2(b * v.sqrt())
                m.lerp_(next_m, 1. - beta1)
                scale = 1.
                if cautious < 1.:
                    align = (grad * m) > 0
                    scale = torch.where(align, torch.ones_like(grad), torch.full_like(grad, cautious))
                    scale = scale / scale.mean().clamp(min=1e-5)
                p.add_(m * scale, alpha=-lr * a)
                v.lerp_(grad * grad, 1. - beta2)
                state['steps'] = steps
