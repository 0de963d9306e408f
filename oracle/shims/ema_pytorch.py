"""Stand-in for `ema-pytorch` EMA (x_transformers_rl.py:38,747,753,1194,1269).
TEST INFRASTRUCTURE.  SURVEY.md Appendix A.5; parity unpinned (third-party)."""
from copy import deepcopy

import torch
from torch import nn


class EMA(nn.Module):
    def __init__(self, model, beta=0.9999, update_after_step=100, update_every=10,
                 inv_gamma=1., power=2 / 3, min_value=0., include_online_model=True,
                 update_model_with_ema_every=None, update_model_with_ema_beta=0.,
                 forward_method_names=()):
        super().__init__()
        self.beta = beta
        self.include_online_model = include_online_model
        if include_online_model:
            self.online_model = model
        else:
            self.online_model = [model]
        self.ema_model = deepcopy(model)
        for p in self.ema_model.parameters():
            p.detach_()
        self.update_every, self.update_after_step = update_every, update_after_step
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.update_model_with_ema_every = update_model_with_ema_every
        self.update_model_with_ema_beta = update_model_with_ema_beta
        for name in forward_method_names:
            setattr(self, name, getattr(self.ema_model, name))
        self.register_buffer('initted', torch.tensor(False))
        self.register_buffer('step', torch.tensor(0))

    @property
    def model(self):
        return self.online_model if self.include_online_model else self.online_model[0]

    def eval(self):
        return self.ema_model.eval()

    @torch.no_grad()
    def forward_eval(self, *args, **kwargs):
        training = self.ema_model.training
        self.ema_model.eval()
        out = self.ema_model(*args, **kwargs)
        self.ema_model.train(training)
        return out

    def add_to_optimizer_post_step_hook(self, optimizer):
        optimizer.register_step_post_hook(lambda *_: self.update())

    def _pairs(self):
        ema, on = self.ema_model, self.model
        for (_, a), (_, b) in zip(ema.named_parameters(), on.named_parameters()):
            yield a, b
        for (_, a), (_, b) in zip(ema.named_buffers(), on.named_buffers()):
            yield a, b

    @torch.no_grad()
    def copy_params_from_model_to_ema(self):
        for a, b in self._pairs():
            a.data.copy_(b.data)

    @torch.no_grad()
    def update_model_with_ema(self, decay=0.):
        for a, b in self._pairs():
            if torch.is_floating_point(b):
                b.data.lerp_(a.data, 1. - decay)
            else:
                b.data.copy_(a.data)

    def get_current_decay(self):
        epoch = max(self.step.item() - self.update_after_step - 1, 0.)
        if epoch <= 0:
            return 0.
        value = 1 - (1 + epoch / self.inv_gamma) ** -self.power
        return min(max(value, self.min_value), self.beta)

    @torch.no_grad()
    def update(self):
        step = self.step.item()
        self.step += 1
        if not self.initted.item():
            self.copy_params_from_model_to_ema()
            self.initted.data.copy_(torch.tensor(True))
            return
        should_update = (step % self.update_every) == 0
        if should_update and step <= self.update_after_step:
            self.copy_params_from_model_to_ema()
            return
        if should_update:
            decay = self.get_current_decay()
            for a, b in self._pairs():
                if torch.is_floating_point(a):
                    a.data.lerp_(b.data, 1. - decay)
                else:
                    a.data.copy_(b.data)
        if self.update_model_with_ema_every is not None and (step % self.update_model_with_ema_every) == 0:
            self.update_model_with_ema(self.update_model_with_ema_beta)

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)
