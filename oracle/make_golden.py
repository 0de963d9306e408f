"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FILES (oracle/ref_loader.py) on
seeded inputs.  TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # rewrites tests/golden/

The vectors pin oracle/restated.py (tests/test_oracle_golden.py, CPU) and are re-used by the
-m gpu parity tests, which cannot read /root/reference on the GPU box.
Third-party arithmetic inside these vectors comes from oracle/shims (parity unpinned there).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')

MODEL_CASES = dict(
    # README Sim shape (README.md:13-42): S5 A2 T10 L1 H4
    c1=dict(state_dim=5, num_actions=2, reward_range=(-1., 1.), max_timesteps=10,
            world_model=dict(attn_dim_head=16, heads=4, depth=1), evolutionary=False),
    # train_lander.py:24-54 shape, shortened: gates + value residual + learned mix + genes
    c2s=dict(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=12,
             world_model=dict(attn_dim_head=16, heads=4, depth=3, attn_gate_values=True,
                              add_value_residual=True, learned_value_residual_mix=True),
             evolutionary=True,
             latent_gene_pool=dict(dim=32, num_genes_per_island=3, num_selected=2, tournament_size=2)),
    # value residual with the constant 0.5 mix, 8 heads
    # continuous actions (tests/test_x_transformers_rl.py:5 parametrises them): squashed (Learner default) and not
    c4c=dict(state_dim=6, num_actions=3, reward_range=(-2., 2.), max_timesteps=8, continuous_actions=True,
             world_model=dict(attn_dim_head=16, heads=4, depth=2, attn_gate_values=True), evolutionary=False),
    c5c=dict(state_dim=6, num_actions=3, reward_range=(-2., 2.), max_timesteps=8, continuous_actions=True, squash_continuous=False,
             world_model=dict(attn_dim_head=16, heads=4, depth=1), evolutionary=False),
    c3s=dict(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=9,
             world_model=dict(attn_dim_head=16, heads=8, depth=2, add_value_residual=True),
             evolutionary=False),
)


def _np(t):
    return t.detach().cpu().numpy().copy()   # copy: parameters are mutated in place later


def build_learner(ref, case, seed, **extra):
    torch.manual_seed(seed)
    kw = dict(MODEL_CASES[case])
    kw.update(extra)
    agent_kwargs = dict(kw.pop('agent_kwargs', {}))
    agent_kwargs.setdefault('save_path', os.path.join('/tmp', f'xtrl_golden_{case}.pt'))
    learner = ref.Learner(batch_size=2, num_episodes_per_update=2, agent_kwargs=agent_kwargs, **kw)
    # move every weight off its init so gates / mixes / biases all matter
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in learner.agent.model.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.15)
        if learner.agent.gene_pool is not None:
            pass
    learner.agent.ema_model.copy_params_from_model_to_ema()
    return learner


def model_case(ref, case, seed):
    learner = build_learner(ref, case, seed)
    agent = learner.agent
    model = agent.model.eval()
    kw = MODEL_CASES[case]
    S, A, T = kw['state_dim'], kw['num_actions'], kw['max_timesteps']
    g = torch.Generator().manual_seed(seed + 2)
    b = 3
    out = {}
    for k, v in model.state_dict().items():
        out[f'sd/{k}'] = _np(v)
    states = torch.randn(b, T, S, generator=g)
    prev_rewards = torch.randn(b, T, generator=g)
    cont = kw.get('continuous_actions', False)
    if cont:
        actions = torch.tanh(torch.randn(b, T, A, generator=g))
        prev_actions = torch.cat((torch.zeros(b, 1, A), actions[:, :-1]), dim=1)
    else:
        actions = torch.randint(0, A, (b, T), generator=g)
        prev_actions = torch.cat((torch.full((b, 1), -1), actions[:, :-1]), dim=1)
    lens = torch.tensor([T, max(T - 3, 2), max(T // 2, 2)])
    mask = torch.arange(T)[None] < lens[:, None]
    gene = None
    if kw['evolutionary']:
        gene_ids = torch.tensor([0, 2, 1])
        gene = agent.gene_pool[gene_ids]
        out['in/genes'] = _np(agent.gene_pool.genes)
        out['in/gene_ids'] = _np(gene_ids)
        out['in/gene'] = _np(gene)

    # (1) full-sequence forward as in Agent.learn :928-935, eval mode (no dropout)
    with torch.no_grad():
        ra, va, sp, dn, _ = model(states, rewards=prev_rewards, actions=prev_actions, latent_gene=gene,
                                  next_actions=actions, mask=mask)
    out.update({'in/states': _np(states), 'in/prev_rewards': _np(prev_rewards), 'in/actions': _np(actions),
                'in/prev_actions': _np(prev_actions), 'in/lens': _np(lens),
                'seq/raw_actions': _np(ra), 'seq/values': _np(va), 'seq/state_pred': _np(sp), 'seq/dones': _np(dn)})

    # (2) token-by-token rollout with the cache, exactly the call of :1269-1275 (batch 1, env 0)
    cache = None
    step_ra, step_va = [], []
    with torch.no_grad():
        for t in range(T):
            ra1, va1, _, _, cache = agent.ema_model.forward_eval(
                states[:1, t:t + 1], rewards=prev_rewards[0, t], latent_gene=None if gene is None else gene[:1],
                cache=cache, actions=prev_actions[:1, t:t + 1])
            step_ra.append(ra1[0, 0]); step_va.append(va1[0, 0])
    out['step/raw_actions'] = _np(torch.stack(step_ra))
    out['step/values'] = _np(torch.stack(step_va))

    # (3) the action distribution on the step outputs (:197-277)
    acts0 = actions[0]
    if cont:
        d = model.action_type_klass(torch.stack(step_ra))
        out['step/mean_variance'] = _np(d.mean_variance)
        out['step/log_prob'] = _np(d.log_prob(acts0))
    else:
        d = ref.Discrete(torch.stack(step_ra))
        out['step/probs'] = _np(d.probs)
        out['step/log_prob'] = _np(d.log_prob(acts0))
        out['step/entropy'] = _np(d.entropy())

    # (4) losses on the full-sequence outputs, with gradients w.r.t. every head output (:939-978)
    old_values = torch.randn(b, T, va.shape[-1], generator=g) * 0.5
    old_log_probs = -torch.rand(*((b, T, A) if cont else (b, T)), generator=g) * 1.5
    lo, hi = kw['reward_range']
    returns = torch.randn(b, T, generator=g) * (hi - lo) * 0.2
    dones = torch.rand(b, T, generator=g) < 0.1
    real = torch.cat((states, prev_rewards[..., None]), dim=-1)
    ra_, va_, sp_, dn_ = (t.clone().requires_grad_() for t in (ra, va, sp, dn))
    wm = model.compute_autoregressive_loss(sp_, real)[mask[:, :-1]]
    dl = model.compute_done_loss(dn_, dones)[mask]
    al = model.compute_actor_loss(ra_, actions, old_log_probs, returns, old_values, mask=mask)
    cl = model.compute_critic_loss(va_, returns, old_values)
    w_a, w_c, w_ar = 0.5, 1., 1.
    loss = (al * w_a + cl * w_c)[mask].mean() + (wm.mean() + dl.mean()) * w_ar
    loss.backward()
    out.update({
        'loss/old_values_logits': _np(old_values),
        'loss/old_values_scalar': _np(model.critic_hl_gauss_loss(old_values)),
        'loss/old_log_probs': _np(old_log_probs), 'loss/returns': _np(returns), 'loss/dones': _np(dones),
        'loss/actor': _np(al), 'loss/critic': _np(cl), 'loss/autoreg_masked': _np(wm), 'loss/done_masked': _np(dl),
        'loss/total': _np(loss), 'loss/weights': np.array([w_a, w_c, w_ar], dtype=np.float32),
        'loss/g_raw_actions': _np(ra_.grad), 'loss/g_values': _np(va_.grad),
        'loss/g_state_pred': _np(sp_.grad), 'loss/g_dones': _np(dn_.grad),
        'loss/scalar_values': _np(model.critic_hl_gauss_loss(va)),
    })
    cfg = dict(case=case, seed=seed, **{k: v for k, v in kw.items()})
    out['config_json'] = np.array(json.dumps(cfg))
    np.savez(os.path.join(OUT, f'model_{case}.npz'), **out)


def gae_case(ref, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, (E, T) in dict(a=(7, 33), b=(3, 1), c=(5, 500)).items():
        rewards = torch.randn(E, T, generator=g)
        values = torch.randn(E, T, generator=g)
        boundary = torch.rand(E, T, generator=g) < 0.07
        returns = ref.calc_gae(rewards, values, ~boundary, gamma=0.99, lam=0.95, use_accelerated=False)
        out.update({f'{name}/rewards': _np(rewards), f'{name}/values': _np(values),
                    f'{name}/boundary': _np(boundary), f'{name}/returns': _np(returns)})
    np.savez(os.path.join(OUT, 'gae.npz'), **out)


POOLS = dict(
    lander=dict(dim=32, num_genes_per_island=3, num_selected=2, tournament_size=2),
    one_island_64=dict(dim=128, num_genes_per_island=64, num_selected=16, tournament_size=4),
    eight_islands=dict(dim=128, num_genes_per_island=8, num_selected=4, tournament_size=2, num_islands=8),
    no_elite=dict(dim=16, num_genes_per_island=6, num_selected=3, tournament_size=3, num_elites=0, num_islands=2),
)


def evolve_case(ref_evo, seed):
    out = {}
    for name, kw in POOLS.items():
        torch.manual_seed(seed)
        pool = ref_evo.LatentGenePool(**kw)
        I, g, s, G = pool.num_islands, pool.num_genes_per_island, pool.num_selected, pool.dim_gene
        c = pool.num_children
        n_steps = 10 if I > 1 else 2          # step 10 exercises the migration branch (:137-160)
        for it in range(n_steps):
            fitness = torch.randn(I * g)
            before = pool.genes.detach().clone()
            torch.manual_seed(seed + 100 + it)
            # the reference's draws, in its order and shapes (evolution.py:113,133,173)
            nt = torch.randn((I, c, s)); nc = torch.randn((I, c, G))
            nm = torch.randn((I, g - 1 if pool.num_elites > 0 else g, G))
            torch.manual_seed(seed + 100 + it)
            sel = pool.evolve_(fitness)
            if it in (0, n_steps - 1):
                tag = f'{name}/{it}'
                out.update({f'{tag}/genes_in': _np(before), f'{tag}/fitness': _np(fitness),
                            f'{tag}/noise_tourn': _np(nt), f'{tag}/noise_cross': _np(nc), f'{tag}/noise_mut': _np(nm),
                            f'{tag}/genes_out': _np(pool.genes), f'{tag}/selected': _np(sel),
                            f'{tag}/step_before': np.array(it)})
        out[f'{name}/config_json'] = np.array(json.dumps(kw))
    np.savez(os.path.join(OUT, 'evolve.npz'), **out)


def rsnorm_case(ref, seed):
    g = torch.Generator().manual_seed(seed)
    norm = ref.RSNorm(6)
    norm.train()
    out = {}
    for it in range(3):
        x = torch.randn(11, 6, generator=g) * 2 + 1
        normed = norm(x)
        out.update({f'{it}/x': _np(x), f'{it}/normed': _np(normed), f'{it}/mean': _np(norm.running_mean),
                    f'{it}/var': _np(norm.running_variance), f'{it}/step': _np(norm.step)})
    np.savez(os.path.join(OUT, 'rsnorm.npz'), **out)


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    for i, case in enumerate(MODEL_CASES):
        model_case(ref, case, seed=100 + i)
    gae_case(ref, 7)
    evolve_case(ref_loader.evolution(), 11)
    rsnorm_case(ref, 13)
    try:
        from oracle import golden_learn
    except ImportError:
        golden_learn = None
    if golden_learn is not None:
        golden_learn.make(ref, OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
