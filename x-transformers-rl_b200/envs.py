"""Synthetic environments of the BASELINE shapes (README.md:16-21 `Sim`, generalised and vectorised).

Additive to the reference's env protocol (x_transformers_rl.py:1232-1237, :1297-1305): an env that
exposes `num_envs` is stepped B-wide with batched tensors:
    reset(seed=None) -> state (B, S)
    step(action (B,) int32 tensor) -> (state (B,S), reward (B,), terminated (B,) bool)
A plain env (no `num_envs`) keeps the reference's one-episode-at-a-time semantics.
"""
from __future__ import annotations

import numpy as np
import torch


class Sim:
    """the README / tests Sim: randn state, randn(1) reward, never done."""

    def __init__(self, state_dim=5):
        self.state_dim = state_dim

    def reset(self, seed=None):
        return np.random.randn(self.state_dim)

    def step(self, actions):
        return np.random.randn(self.state_dim), np.random.randn(1), False


class SyntheticVecEnv:
    """B independent Sim-like envs resident on the GPU: state ~ N(0,1)^S, reward ~ N(0,1), never done.
    Draws a whole block of `horizon` steps per reset so a step is a slice, not a launch."""

    def __init__(self, num_envs, state_dim, horizon, device='cuda', seed=0):
        self.num_envs, self.state_dim, self.horizon, self.device = num_envs, state_dim, horizon, torch.device(device)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.never_terminates = True
        # step() only hands out views of tensors that keep their storage across reset(): a whole rollout over
        # this env can be captured in one CUDA graph and replayed (learner.Learner.use_cuda_graph)
        self.cuda_graph_safe = True
        self._t = 0
        B, S, H = num_envs, state_dim, horizon
        self.states = torch.empty(H + 1, B, S, device=self.device)
        self.rewards = torch.empty(H, B, device=self.device)
        self.done = torch.zeros(B, dtype=torch.bool, device=self.device)
        self._draw()

    def _draw(self):
        self.states.normal_(generator=self.gen)       # in place: storage (and any captured graph) stays valid
        self.rewards.normal_(generator=self.gen)

    def reset(self, seed=None):
        if seed is not None:
            self.gen.manual_seed(int(seed))
        self._draw()
        self._t = 0
        return self.states[0]

    def step(self, action):
        t = self._t
        self._t += 1
        return self.states[t + 1], self.rewards[t], self.done


class HostVecEnv:
    """the same synthetic env living in HOST memory (numpy in, numpy out): what a user's CPU simulator
    farm looks like.  The learner copies states host->device and actions device->host every step."""

    def __init__(self, num_envs, state_dim, horizon, seed=0):
        self.num_envs, self.state_dim, self.horizon = num_envs, state_dim, horizon
        self.rng = np.random.default_rng(seed)
        self.never_terminates = True
        self._t = 0
        self._draw()

    def _draw(self):
        B, S, H = self.num_envs, self.state_dim, self.horizon
        self.states = self.rng.standard_normal((H + 1, B, S), dtype=np.float32)
        self.rewards = self.rng.standard_normal((H, B), dtype=np.float32)
        self.done = np.zeros(B, dtype=bool)

    def reset(self, seed=None):
        """rewinds to the start of the pre-drawn block; a seed draws a new block (numpy draws 16 M normals in
        ~0.2 s at the c3 shape -- simulator cost, kept out of the per-update path unless asked for)."""
        if seed is not None:
            self.rng = np.random.default_rng(int(seed))
            self._draw()
        self._t = 0
        return self.states[0]

    def step(self, action):
        assert action.shape == (self.num_envs,)
        t = self._t
        self._t += 1
        return self.states[t + 1], self.rewards[t], self.done


class _HostSubEnv:
    """envs [lo, hi) of a HostVecEnv, steppable on their own (lets the learner pipeline groups of envs)."""

    def __init__(self, parent, lo, hi):
        self.parent, self.lo, self.hi = parent, lo, hi
        self.num_envs, self.state_dim, self.never_terminates = hi - lo, parent.state_dim, True
        self._t = 0

    def reset(self, seed=None):
        self._t = 0
        return self.parent.states[0, self.lo:self.hi]

    def step(self, action):
        assert action.shape == (self.num_envs,)
        t = self._t
        self._t += 1
        return self.parent.states[t + 1, self.lo:self.hi], self.parent.rewards[t, self.lo:self.hi], self.parent.done[self.lo:self.hi]


def _host_split(self, n):
    """n equally sized sub-envs (views).  Call reset() on the parent first to draw a new block."""
    assert self.num_envs % n == 0
    b = self.num_envs // n
    return [_HostSubEnv(self, i * b, (i + 1) * b) for i in range(n)]


HostVecEnv.split = _host_split
