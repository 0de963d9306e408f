"""shared test helpers: golden loading, reference-style config -> oracle Config."""
import json
import os

import numpy as np
import torch

from oracle import restated as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def t(a):
    return torch.from_numpy(np.asarray(a))


def split(z, prefix):
    p = prefix + '/'
    return {k[len(p):]: t(v) for k, v in z.items() if k.startswith(p)}


def oracle_config(kw, **over) -> R.Config:
    """Learner kwargs (as stored in config_json) -> oracle Config."""
    wm = kw['world_model']
    pool = kw.get('latent_gene_pool') or {}
    cfg = R.Config(
        state_dim=kw['state_dim'], num_actions=kw['num_actions'], reward_range=tuple(kw['reward_range']),
        heads=wm['heads'], dim_head=wm['attn_dim_head'], depth=wm['depth'],
        gate_values=wm.get('attn_gate_values', False), value_residual=wm.get('add_value_residual', False),
        learned_mix=wm.get('learned_value_residual_mix', False),
        evolutionary=kw.get('evolutionary', False), dim_gene=pool.get('dim', 0) if kw.get('evolutionary') else 0,
        continuous=kw.get('continuous_actions', False), squash=kw.get('squash_continuous', True),
    )
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def golden_model(case):
    z = load_golden(f'model_{case}.npz')
    kw = json.loads(str(z['config_json']))
    return z, kw, oracle_config(kw), split(z, 'sd')


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


# ---- product package + random models (GPU tests) --------------------------------------------

def xtrl():
    """the product package (directory name has a hyphen -> importlib)."""
    import importlib
    return importlib.import_module('x-transformers-rl_b200')


def model_config(cfg: R.Config, max_timesteps, **extra):
    """oracle Config -> product ModelConfig (same shapes)."""
    ops = xtrl().ops
    return ops.ModelConfig(state_dim=cfg.state_dim, num_actions=cfg.num_actions, reward_range=cfg.reward_range,
                           max_timesteps=max_timesteps, dim=cfg.dim, heads=cfg.heads, dim_head=cfg.dim_head,
                           depth=cfg.depth, num_bins=cfg.num_bins, dim_gene=cfg.dim_gene,
                           gate_values=cfg.gate_values, value_residual=cfg.value_residual,
                           learned_mix=cfg.learned_mix, continuous=cfg.continuous, squash=cfg.squash, **extra)


def random_state_dict(cfg: R.Config, seed=0, scale=0.2):
    """reference-named state_dict with random weights of the right shapes (no reference needed)."""
    g = torch.Generator().manual_seed(seed)
    D, I, H, S, A, K, G = cfg.dim, cfg.inner, cfg.heads, cfg.state_dim, cfg.num_actions, cfg.num_bins, cfg.dim_gene
    rn = lambda *shape: torch.randn(*shape, generator=g) * scale
    hin = (3 if cfg.evolutionary else 2) * D
    sd = {
        'reward_embed': rn(D),
        'to_state_embed.weight': rn(D, S), 'to_state_embed.bias': rn(D),
        'to_pred_done.0.weight': rn(1, 2 * D), 'to_pred_done.0.bias': rn(1),
        'to_pred.0.weight': rn(D, 2 * D), 'to_pred.0.bias': rn(D),
        'to_pred.2.weight': rn(2 * (S + 1), D), 'to_pred.2.bias': rn(2 * (S + 1)),
        'critic_head.0.weight': rn(2 * D, hin), 'critic_head.0.bias': rn(2 * D),
        'critic_head.2.weight': rn(K, 2 * D), 'critic_head.2.bias': rn(K),
        'action_head.0.weight': rn(2 * D, hin), 'action_head.0.bias': rn(2 * D),
        'action_head.2.weight': rn(2 * A if cfg.continuous else A, 2 * D), 'action_head.2.bias': rn(2 * A if cfg.continuous else A),
        'transformer.project_in.weight': rn(D, S),
        'transformer.attn_layers.final_norm.gamma': 1 + rn(D),
    }
    if cfg.continuous:
        sd['action_embeds.weight'], sd['action_embeds.bias'] = rn(D, A), rn(D)
    else:
        sd['action_embeds.embed.weight'] = rn(A, D)
    if cfg.evolutionary:
        sd['latent_to_embed.weight'] = rn(D, G)
        sd['latent_to_embed.bias'] = rn(D)
    for l in range(cfg.depth):
        pa, pf = R._L(l)
        sd[f'{pa}.0.gamma'] = 1 + rn(D)
        for n in ('to_q', 'to_k', 'to_v'):
            sd[f'{pa}.1.{n}.weight'] = rn(I, D)
        sd[f'{pa}.1.to_out.weight'] = rn(D, I)
        if cfg.gate_values:
            sd[f'{pa}.1.to_v_gate.weight'] = rn(I, D)
            sd[f'{pa}.1.to_v_gate.bias'] = rn(I) + 1.
        if cfg.learned_mix and l > 0:
            sd[f'{pa}.1.to_value_residual_mix.0.weight'] = rn(H, D)
            sd[f'{pa}.1.to_value_residual_mix.0.bias'] = rn(H)
        sd[f'{pf}.0.gamma'] = 1 + rn(D)
        sd[f'{pf}.1.ff.0.0.weight'] = rn(4 * D, D)
        sd[f'{pf}.1.ff.0.0.bias'] = rn(4 * D)
        sd[f'{pf}.1.ff.2.weight'] = rn(D, 4 * D)
        sd[f'{pf}.1.ff.2.bias'] = rn(D)
    return sd


# ---- goldens of the reference's own Agent.learn (oracle/golden_learn.py) ----------------------

def learn_golden(name):
    """-> (z, meta, oracle Config, sd0 (dict of float tensors), data dict) of tests/golden/learn_<name>.npz"""
    z = load_golden(f'learn_{name}.npz')
    meta = json.loads(str(z['config_json']))
    cfg = oracle_config(meta['learner'])
    sd0 = {k: v for k, v in split(z, 'sd0').items() if v.is_floating_point() and 'inv_freq' not in k}
    i = split(z, 'in')
    data = dict(states=i['states'], actions=i['actions'], log_probs=i['log_probs'], rewards=i['rewards'],
                is_boundary=i['is_boundary'].bool(), value_logits=i['value_logits'], lens=i['lens'], gene_ids=i['gene_ids'])
    return z, meta, cfg, sd0, data


def learn_golden_noises(z, u):
    """the evolve_ noise triples recorded for update u, in call order"""
    out, j = [], 0
    while f'u{u}/noise{j}/tourn' in z:
        out.append(tuple(t(z[f'u{u}/noise{j}/{k}']) for k in ('tourn', 'cross', 'mut')))
        j += 1
    return out
