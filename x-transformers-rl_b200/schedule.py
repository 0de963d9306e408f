"""Host-side schedules that decide WHAT the optimiser kernel does each step (no tensor math here).

EmaSchedule mirrors ema_pytorch.EMA.update as the reference configures it
(x_transformers_rl.py:687-689, :747, :753): first call copies, every
`update_every`-th step copies while step <= update_after_step and lerps afterwards with
decay = clamp(1 - (1 + step - after - 1)^-power, min_value, beta), and every
`update_model_with_ema_every`-th step the EMA weights are copied back into the online model.
"""
from __future__ import annotations

EMA_NONE, EMA_COPY, EMA_LERP = 0, 1, 2


class EmaSchedule:
    def __init__(self, beta=0.9, update_after_step=100, update_every=10, inv_gamma=1., power=2 / 3,
                 min_value=0., update_model_with_ema_every=1250):
        self.beta, self.after, self.every = beta, update_after_step, update_every
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.copy_back_every = update_model_with_ema_every
        self.step, self.initted = 0, False

    def current_decay(self) -> float:
        epoch = max(self.step - self.after - 1, 0.)
        if epoch <= 0:
            return 0.
        value = 1 - (1 + epoch / self.inv_gamma) ** -self.power
        return min(max(value, self.min_value), self.beta)

    def next(self):
        """-> (ema_mode, ema_decay, copy_ema_to_online) for the optimiser step about to run."""
        step = self.step
        self.step += 1
        if not self.initted:
            self.initted = True
            return EMA_COPY, 0., False
        should = step % self.every == 0
        if should and step <= self.after:
            return EMA_COPY, 0., False
        mode, decay = (EMA_LERP, self.current_decay()) if should else (EMA_NONE, 0.)
        back = self.copy_back_every is not None and step % self.copy_back_every == 0
        return mode, decay, back

    def state_dict(self):
        return dict(step=self.step, initted=self.initted)

    def load_state_dict(self, sd):
        self.step, self.initted = sd['step'], sd['initted']
