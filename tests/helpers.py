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
