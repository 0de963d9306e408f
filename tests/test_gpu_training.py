"""GPU: the whole loop learns.  A contextual-bandit-like vector env (reward +1 when the action matches the sign of
state[0], else -1) is solved by PPO through the kernels: mean reward per step rises well above chance.  Parity tests
pin every piece against the oracle; this one checks that the pieces compose into a working learner (signs of the
gradients, EMA hand-over to the rollout kernel, RSNorm updates, optimiser)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


class SignEnv:
    """B envs on the device; never terminates; step() is pure tensor ops (not graph-safe: rewards depend on actions)."""

    def __init__(self, num_envs, horizon, device, seed=0):
        self.num_envs, self.horizon, self.device = num_envs, horizon, device
        self.never_terminates = True
        self.gen = torch.Generator(device=device).manual_seed(seed)
        self.done = torch.zeros(num_envs, dtype=torch.bool, device=device)

    def reset(self, seed=None):
        self.state = torch.randn(self.num_envs, 4, device=self.device, generator=self.gen)
        return self.state

    def step(self, action):
        reward = torch.where(action == (self.state[:, 0] > 0).int(), 1., -1.)
        self.state = torch.randn(self.num_envs, 4, device=self.device, generator=self.gen)
        return self.state, reward, self.done


def test_ppo_learns_sign_task(tmp_path):
    import x_transformers_rl
    torch.manual_seed(0)
    B, T = 256, 8
    learner = x_transformers_rl.Learner(
        state_dim=4, num_actions=2, reward_range=(-10., 10.), max_timesteps=T, batch_size=64, num_episodes_per_update=B,
        epochs=2, lr=3e-3, world_model=dict(attn_dim_head=16, heads=4, depth=1),
        agent_kwargs=dict(save_path=str(tmp_path / 'ppo.pt'), dropout=0.))
    env = SignEnv(B, T, learner.device)
    gids = torch.zeros(B, dtype=torch.int32, device=learner.device)
    rewards = []
    for update in range(40):
        learner(env, 1)
        buf, _ = learner.rollout_vectorized(env, gids, T)
        rewards.append(buf.rewards.mean().item())
    first, last = sum(rewards[:3]) / 3, sum(rewards[-3:]) / 3
    assert abs(first) < 0.35, rewards[:3]                 # chance level at the start
    assert last > first + 0.3 and last > 0.4, (first, last, rewards[-6:])
