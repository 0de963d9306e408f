"""GPU: the drop-in Learner / Agent API (reference tests/test_x_transformers_rl.py:4-53), the learn-side
model + fused loss against the oracle's autograd, and the vectorised rollout against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import restated as R
from helpers import rel_err, xtrl, oracle_config

pytestmark = pytest.mark.gpu


def _learner_cls():
    import x_transformers_rl            # the drop-in import name
    return x_transformers_rl.Learner


class Sim:                                # README.md:16-21 / reference tests :10-15
    def reset(self, seed=None):
        return np.random.randn(5)

    def step(self, actions):
        return np.random.randn(5), np.random.randn(1), False


@pytest.mark.parametrize('evolutionary', (False, True))
@pytest.mark.parametrize('continuous_actions', (False, True))
def test_e2e_reference_smoke(evolutionary, continuous_actions, tmp_path):
    """the reference's own (only) test, all four parameter combinations (tests/test_x_transformers_rl.py:4-53):
    construct, learn one update, act twice."""
    Learner = _learner_cls()
    learner = Learner(
        state_dim=5, num_actions=2, reward_range=(-1., 1.), max_timesteps=10, batch_size=2, num_episodes_per_update=2,
        evolutionary=evolutionary, continuous_actions=continuous_actions,
        latent_gene_pool=dict(dim=32, num_genes_per_island=3, num_selected=2, tournament_size=2),
        world_model=dict(attn_dim_head=16, heads=4, depth=1),
        agent_kwargs=dict(save_path=str(tmp_path / 'ppo.pt')))
    learner(Sim(), 1)
    hiddens = None
    actions, hiddens = learner.agent(np.random.randn(5), hiddens=hiddens)
    actions, hiddens = learner.agent(np.random.randn(5), hiddens=hiddens)
    assert actions.shape == ((4,) if continuous_actions else (2,)) and torch.isfinite(actions).all()
    assert os.path.exists(tmp_path / 'ppo.pt')
    assert learner.timings['env_steps'] == 2 * 10 * (3 if evolutionary else 1)
    assert learner.accelerator.is_main_process and len(learner.episode_genes_for_process) == (6 if evolutionary else 2)
    sd = torch.load(tmp_path / 'ppo.pt', weights_only=True)['model']
    assert 'transformer.attn_layers.layers.0.1.to_q.weight' in sd and 'critic_head.2.bias' in sd


def test_continuous_vectorized_rollout_and_clamp(tmp_path):
    """continuous actions through the vectorised rollout: tanh-squashed samples, environment clamp (:1291-1295)."""
    from importlib import import_module
    envs = import_module('x-transformers-rl_b200.envs')
    learner = _make(tmp_path, 8, 12, num_episodes_per_update=32, continuous_actions=True, continuous_actions_clamp=(-0.5, 0.5))
    env = envs.SyntheticVecEnv(32, 8, 12, device=learner.device, seed=2)
    learner(env, 2, seed=0)
    gids = torch.zeros(32, dtype=torch.int32, device=learner.device)
    buf, _ = learner.rollout_vectorized(env, gids, 12)
    assert buf.actions.shape == (12, 32, 4) and buf.log_probs.shape == (12, 32, 4)
    assert buf.actions.abs().max() <= 0.5 and torch.isfinite(buf.log_probs).all()
    assert torch.isfinite(learner.agent.model.flat_params).all()


LANDER_WM = dict(attn_dim_head=16, heads=4, depth=2, attn_gate_values=True, add_value_residual=True,
                 learned_value_residual_mix=True)


def _make(tmp_path, B, T, **kw):
    Learner = _learner_cls()
    args = dict(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=T, batch_size=B,
                num_episodes_per_update=B, world_model=LANDER_WM, epochs=1,
                accelerate_kwargs=dict(mixed_precision='no'),          # the fp32 parity class (the default is the bf16 class)
                agent_kwargs=dict(save_path=str(tmp_path / 'ppo.pt'), dropout=0.))
    args.update(kw)
    return Learner(**args)


def test_vectorized_rollout_matches_oracle(tmp_path):
    """B envs in lock step through the kernel == the oracle stepping the same envs with the same uniforms."""
    pkg = xtrl()
    from importlib import import_module
    envs = import_module('x-transformers-rl_b200.envs')
    B, T = 24, 12
    torch.manual_seed(0)
    learner = _make(tmp_path, B, T)
    agent = learner.agent
    with torch.no_grad():                  # move the weights off their init so gates / mixes matter
        agent.ema_params.add_(torch.randn_like(agent.ema_params) * 0.1)
        agent.rsnorm.running_mean.normal_(); agent.rsnorm.running_variance.uniform_(0.5, 2.)
    env = envs.SyntheticVecEnv(B, 8, T, device=learner.device, seed=3)
    u = torch.rand(T, B, device=learner.device)
    gids = torch.zeros(B, dtype=torch.int32, device=learner.device)
    buf, cum = learner.rollout_vectorized(env, gids, T, uniforms=u)
    sd = {k: v.cpu() for k, v in agent.ema_state_dict().items()}
    cfg = oracle_config(dict(state_dim=8, num_actions=4, reward_range=(-5., 5.), world_model=LANDER_WM))
    mean, var = agent.rsnorm.running_mean.cpu(), agent.rsnorm.running_variance.cpu()
    states, rewards, uc = env.states[:T].cpu(), env.rewards.cpu(), u.cpu()
    cache, prev_a, prev_r = None, torch.full((B, 1), -1), torch.zeros(B, 1)
    n_safe = 0
    for t in range(T):
        packed = R.rsnorm_eval(torch.cat((states[t], prev_r), -1), mean, var)
        ra, va, _, _, cache = R.wm_forward(sd, cfg, packed[:, None, :-1], prev_a, packed[:, -1:], 1., cache=cache)
        probs, logp, _ = R.discrete_stats(ra[:, 0])
        got_a = buf.actions[t].cpu().long()
        margin = (probs.cumsum(-1) - uc[t][:, None]).abs().min(-1).values
        safe = margin > 1e-4
        assert torch.equal(got_a[safe], R.sample_inverse_cdf(probs, uc[t])[safe])
        n_safe += int(safe.sum())
        assert rel_err(buf.values[t].cpu(), R.hl_decode(va[:, 0], cfg)) < 5e-5
        assert rel_err(buf.log_probs[t].cpu(), logp.gather(-1, got_a[:, None])[:, 0]) < 5e-5
        prev_a, prev_r = got_a[:, None], rewards[t][:, None]       # follow the kernel's actions
    assert n_safe > 0.98 * B * T
    assert torch.equal(buf.episode_lens.cpu(), torch.full((B,), T, dtype=torch.int32))
    assert rel_err(cum.cpu(), rewards.sum(0)) < 1e-5


@pytest.mark.parametrize('evolutionary,continuous', ((False, False), (True, False), (False, True)))
def test_learn_gradients_match_oracle_autograd(tmp_path, evolutionary, continuous):
    """one minibatch of Agent.learn (dropout 0): GAE kernel + torch sequence model + fused loss kernel +
    autograd  ==  the oracle's learn_minibatch differentiated by autograd, gradient for gradient."""
    from importlib import import_module
    agent_mod = import_module('x-transformers-rl_b200.agent')
    B, T = 6, 14
    torch.manual_seed(1)
    kw = dict(evolutionary=True, latent_gene_pool=dict(dim=32, num_genes_per_island=3, num_selected=2, tournament_size=2),
              num_episodes_per_update=2, batch_size=2) if evolutionary else {}
    if continuous:
        kw['continuous_actions'] = True
    learner = _make(tmp_path, B, T, **kw)
    agent, dev = learner.agent, learner.device
    agent.batch_size = B
    with torch.no_grad():
        agent.model.flat_params.add_(torch.randn_like(agent.model.flat_params) * 0.1)
        agent.rsnorm.running_mean.normal_(); agent.rsnorm.running_variance.uniform_(0.5, 2.)
    g = torch.Generator().manual_seed(5)
    buf = agent_mod.RolloutBuffer.allocate(T, B, 8, dev, continuous_dim=4 if continuous else 0)
    lens = torch.tensor([T, T - 3, 5, T, 2, T - 1], dtype=torch.int32)
    tm_mask = (torch.arange(T)[:, None] < lens[None, :])
    acts = torch.tanh(torch.randn(T, B, 4, generator=g)) * 0.9 if continuous else torch.randint(0, 4, (T, B), generator=g).int()
    data = dict(states=torch.randn(T, B, 8, generator=g), actions=acts,
                log_probs=-torch.rand(*((T, B, 4) if continuous else (T, B)), generator=g) * 1.5, rewards=torch.randn(T, B, generator=g),
                values=torch.randn(T, B, generator=g), is_boundary=torch.rand(T, B, generator=g) < 0.05)
    for k, v in data.items():              # padding rows hold pad_sequence zeros, like the reference
        v = torch.where(tm_mask if v.ndim == 2 else tm_mask[..., None], v, torch.zeros_like(v))
        data[k] = v
        getattr(buf, k).copy_(v)
    buf.episode_lens.copy_(lens)
    gene_ids = torch.tensor([0, 1, 2, 0, 1, 2], dtype=torch.int32)
    buf.gene_ids.copy_(gene_ids)
    agent.keep_last_grads = True
    sd = {k: v.detach().cpu().clone().requires_grad_() for k, v in agent.model.state_dict().items() if v.is_floating_point()}
    mean, var = agent.rsnorm.running_mean.cpu().clone(), agent.rsnorm.running_variance.cpu().clone()
    genes = agent.gene_pool.genes.detach().cpu().clone() if evolutionary else None
    processed = agent.learn(buf, fitnesses=None, permutations=[torch.arange(B)])
    assert processed == int(lens.sum())
    # oracle
    wm = dict(LANDER_WM)
    cfg = oracle_config(dict(state_dim=8, num_actions=4, reward_range=(-5., 5.), world_model=wm, evolutionary=evolutionary,
                             latent_gene_pool=dict(dim=32), continuous_actions=continuous))
    em = lambda t: t.transpose(0, 1).contiguous()
    returns = R.calc_gae(em(data['rewards']), em(data['values']), ~em(data['is_boundary']))
    batch = dict(states=em(data['states']), actions=em(data['actions']) if continuous else em(data['actions']).long(), rewards=em(data['rewards']),
                 old_log_probs=em(data['log_probs']), returns=returns, old_values=em(data['values']),
                 dones=em(data['is_boundary']), gene_ids=gene_ids.long(), episode_lens=lens.long())
    loss, logs, packed_masked = R.learn_minibatch(sd, cfg, batch, mean, var, reward_flag=1., genes=genes)
    loss.backward()
    # reward dropout: the product drew its own Bernoulli flag; replay the oracle with the flag it used
    got_total = agent.last_logs[0].item()
    if abs(got_total - loss.item()) > 1e-5 * abs(loss.item()):
        for p in sd.values():
            p.grad = None
        loss, logs, packed_masked = R.learn_minibatch(sd, cfg, batch, mean, var, reward_flag=0., genes=genes)
        loss.backward()
    assert rel_err(torch.tensor(got_total), loss.detach()) < 1e-5
    names = [n for n, _ in agent.model.named_parameters()]
    offs = agent.model._offsets
    worst, worst_name = 0., ''
    for n, o0, o1 in zip(names, offs[:-1], offs[1:]):
        want = sd[n].grad if sd[n].grad is not None else torch.zeros_like(sd[n])
        got = agent.last_grads[o0:o1].view(want.shape).cpu()
        if want.abs().max() < 1e-7:
            assert got.abs().max() < 1e-6, n
            continue
        if rel_err(got, want) > worst:
            worst, worst_name = rel_err(got, want), n
        assert rel_err(got, want) < 1e-5, n          # north star: 1e-5 relative in fp32 (measured worst tensor: 2e-6)
    print('GRAD_WORST', evolutionary, continuous, worst_name, worst)
    # RSNorm copy was updated from the batch mean of the already-normalised rows
    m2, v2, _ = R.rsnorm_update(packed_masked.detach(), mean, var, 1)
    assert rel_err(agent.rsnorm.running_mean.cpu(), m2) < 1e-4
    assert agent.rsnorm.time == 2 and agent._step_host == 1 and agent.opt_steps == 1


def test_vectorized_learner_runs_updates(tmp_path):
    from importlib import import_module
    envs = import_module('x-transformers-rl_b200.envs')
    B, T = 32, 16
    learner = _make(tmp_path, 8, T, num_episodes_per_update=B, epochs=2, evolutionary=True, evolve_every=1,
                    evolve_after_step=0,
                    latent_gene_pool=dict(dim=32, num_genes_per_island=4, num_selected=2, tournament_size=2))
    B_total = B * 4
    env = envs.SyntheticVecEnv(B_total, 8, T, device=learner.device, seed=1)
    before = learner.agent.model.flat_params.clone()
    genes_before = learner.agent.gene_pool.genes.clone()
    learner(env, 3, seed=0)
    assert learner.timings['env_steps'] == 3 * B_total * T
    assert learner.timings['learn_timesteps'] == 3 * 2 * B_total * T
    assert torch.isfinite(learner.agent.model.flat_params).all()
    assert not torch.equal(before, learner.agent.model.flat_params)
    assert not torch.equal(genes_before, learner.agent.gene_pool.genes)          # evolve_ ran (step > evolve_after_step)
    assert rel_err(learner.agent.gene_pool.genes.norm(dim=-1).cpu(), torch.ones(4)) < 1e-5
    # host-memory env: same API, states cross PCIe every step
    henv = envs.HostVecEnv(B_total, 8, T, seed=1)
    learner(henv, 1)
    assert learner.timings['env_steps'] == 4 * B_total * T


def test_bf16_learn_step(tmp_path):
    from importlib import import_module
    envs = import_module('x-transformers-rl_b200.envs')
    learner = _make(tmp_path, 8, 16, num_episodes_per_update=16, accelerate_kwargs=dict(mixed_precision='bf16'))
    env = envs.SyntheticVecEnv(16, 8, 16, device=learner.device, seed=1)
    learner(env, 2, seed=0)
    assert torch.isfinite(learner.agent.model.flat_params).all()


def test_pipelined_host_env_rollout_equals_single_stream(tmp_path):
    """groups of envs pipelined on separate streams (host-memory env) == the one-stream rollout, bit for bit."""
    from importlib import import_module
    envs = import_module('x-transformers-rl_b200.envs')
    B, T = 512, 12
    learner = _make(tmp_path, 8, T, num_episodes_per_update=B)
    with torch.no_grad():
        learner.agent.ema_params.add_(torch.randn_like(learner.agent.ema_params) * 0.1)
    u = torch.rand(T, B, device=learner.device)
    gids = torch.zeros(B, dtype=torch.int32, device=learner.device)
    henv = envs.HostVecEnv(B, 8, T, seed=5)
    learner.rollout_groups = 1
    a, cum_a = learner.rollout_vectorized(henv, gids, T, seed=9, uniforms=u)
    learner.rollout_groups = 2
    b, cum_b = learner.rollout_vectorized(henv, gids, T, seed=9, uniforms=u)
    for name in ('states', 'actions', 'log_probs', 'rewards', 'values', 'episode_lens'):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.equal(cum_a, cum_b)


# ---- Agent.learn as a whole against the reference's own learn() (goldens: oracle/golden_learn.py) ----------------------

@pytest.mark.parametrize('name', ('c1', 'c2s', 'c4c'))
def test_agent_learn_matches_reference_golden(name, tmp_path):
    """the CUDA Agent.learn (fp32 parity class) through two learning updates == the reference's Agent.learn
    (x_transformers_rl.py:808-1023) on the same memories (list[list[Memory]] with value LOGITS, ragged, one terminated
    episode), the same shuffle, reward-dropout flags and evolve noise: parameters after clip + AdoptAtan2, the EMA copy
    (copy, lerp and copy-back all inside the 8 steps), RSNorm statistics, genes, and the logged losses."""
    import json
    from importlib import import_module
    from helpers import learn_golden, learn_golden_noises, split
    agent_mod = import_module('x-transformers-rl_b200.agent')
    z, meta, cfg, sd0, data = learn_golden(name)
    kw = dict(meta['learner'])
    Learner = _learner_cls()
    learner = Learner(batch_size=meta['batch_size'], num_episodes_per_update=2, epochs=meta['epochs'], **kw, **meta['extra'],
                      accelerate_kwargs=dict(mixed_precision='no'),
                      agent_kwargs=dict(save_path=str(tmp_path / 'ppo.pt'), dropout=0., ema_kwargs=dict(meta['ema_kwargs'])))
    agent, dev = learner.agent, learner.device
    evo = bool(kw.get('evolutionary'))
    i = split(z, 'in')
    agent.model.load_state_dict({k: v.to(dev) for k, v in split(z, 'sd0').items()}, strict=True)
    agent.ema_params.copy_(agent.model.flat_params)
    agent.rsnorm.running_mean.copy_(i['rs_mean']); agent.rsnorm.running_variance.copy_(i['rs_var'])
    agent.rsnorm.step.fill_(int(i['rs_step'])); agent.rsnorm.time = int(i['rs_step'])
    agent.step.fill_(meta['agent_step0']); agent._step_host = meta['agent_step0']
    if evo:
        agent.gene_pool.genes.data.copy_(i['genes'])
    lens = data['lens']
    memories = [[agent_mod.Memory(data['states'][e, t], data['actions'][e, t], data['log_probs'][e, t], data['rewards'][e, t],
                                  data['is_boundary'][e, t], data['value_logits'][e, t]) for t in range(int(lens[e]))]
                for e in range(len(lens))]
    fit = i['fitnesses'].to(dev) if evo else None
    sd_names = [n for n, _ in agent.model.named_parameters()]
    worst = {}
    for u in range(meta['updates']):
        g = split(z, f'u{u}')
        if evo:
            noises = iter(learn_golden_noises(z, u))
            agent.gene_pool.draw_noise = lambda device, _n=noises: tuple(t.to(device) for t in next(_n))
        agent.learn(memories, lens, data['gene_ids'], fit, permutations=[p for p in g['perms']], reward_flags=g['flags'].tolist())
        got_sd = {k: v.detach().cpu() for k, v in agent.model.state_dict().items()}
        got_ema = {k: v.cpu() for k, v in agent.ema_state_dict().items()}
        for k, v in split(z, f'u{u}/sd').items():
            if k in sd_names:
                worst[k] = max(worst.get(k, 0.), rel_err(got_sd[k], v))
        for k, v in split(z, f'u{u}/ema').items():
            if k in sd_names:
                worst['ema/' + k] = max(worst.get('ema/' + k, 0.), rel_err(got_ema[k], v))
        assert rel_err(agent.rsnorm.running_mean.cpu(), g['rs_mean']) < 1e-5
        assert rel_err(agent.rsnorm.running_variance.cpu(), g['rs_var']) < 1e-5
        assert agent.rsnorm.time == int(g['rs_step']) and agent._step_host == int(g['agent_step'])
        if evo:
            assert rel_err(agent.gene_pool.genes.data.cpu(), g['genes']) < 1e-5
        logs = agent.last_logs.cpu()                     # scalars: total, actor, critic, autoreg, done, ...
        for j, k in enumerate(('actor_loss', 'critic_loss', 'autoreg_loss', 'pred_done_loss')):
            assert rel_err(logs[1 + j], g[f'logs/{k}']) < 2e-5, (u, k, float(logs[1 + j]), float(g[f'logs/{k}']))
    bad = {k: v for k, v in worst.items() if v > 1e-5}        # measured worst tensor over both updates: 7e-7
    print('AGENT_LEARN_WORST', name, json.dumps(dict(sorted(worst.items(), key=lambda kv: -kv[1])[:5])))
    assert not bad, bad


def test_plain_env_padding_of_scalar_values_uses_the_support_midpoint(tmp_path):
    """ragged episodes through the reference-format path with SCALAR values (what rollout_episode stores) and an asymmetric
    reward range: the padding must decode like the reference's zero logits do -- (lo + hi) / 2 -- because the padded value
    feeds v_{t+1} of a shorter episode that did not terminate (:837, :843)."""
    from importlib import import_module
    agent_mod = import_module('x-transformers-rl_b200.agent')
    Learner = _learner_cls()
    learner = Learner(state_dim=4, num_actions=2, reward_range=(0., 500.), max_timesteps=6, batch_size=2, num_episodes_per_update=2,
                      world_model=dict(attn_dim_head=16, heads=4, depth=1), accelerate_kwargs=dict(mixed_precision='no'),
                      agent_kwargs=dict(save_path=str(tmp_path / 'ppo.pt'), dropout=0.))
    agent, dev = learner.agent, learner.device
    g = torch.Generator().manual_seed(0)
    lens = [6, 3]
    mk = lambda v: agent_mod.Memory(torch.randn(4, generator=g), torch.tensor(1), torch.tensor(-0.7), torch.randn((), generator=g),
                                    torch.tensor(False), v)
    scal = [[mk(torch.rand((), generator=g) * 400 + 50) for _ in range(n)] for n in lens]
    buf = agent._buffer_from_memories(scal, torch.tensor(lens), torch.zeros(2, dtype=torch.long))
    assert torch.equal(buf['old_values'][1, 3:].cpu(), torch.full((3,), 250.))
    logits = [[m._replace(value=torch.zeros(100)) for m in ep] for ep in scal]          # zero logits decode to the midpoint too
    buf2 = agent._buffer_from_memories(logits, torch.tensor(lens), torch.zeros(2, dtype=torch.long))
    assert rel_err(buf2['old_values'].cpu(), torch.full((2, 6), 250.)) < 1e-6
    ret = xtrl().ops.gae(buf['rewards'].contiguous(), buf['old_values'].contiguous(), buf['dones'].contiguous())
    want = R.calc_gae(buf['rewards'].cpu(), buf['old_values'].cpu(), ~buf['dones'].cpu())
    assert rel_err(ret.cpu(), want) < 1e-5
