"""Golden vectors of the reference's OWN `Agent.learn` (x_transformers_rl.py:808-1023), executed verbatim from
/root/reference under the shims (oracle/ref_loader.py).  TEST INFRASTRUCTURE; container only; called by
`python -m oracle.make_golden`, writes tests/golden/learn_*.npz.

What is pinned: minibatch prep :903-926, the four losses and their assembly :939-978, backward, clip :987,
optimizer.step + the EMA post-step hook :989-992 (:747-753), the RSNorm copy update :1005, evolve_ inside the minibatch
loop :1009-1015, rsnorm write-back and step :1021-1023 -- for two consecutive learning updates on the same memories.

The reference draws three things from torch's global RNG inside `learn`; each is recorded so that the restatement
(oracle/restated.py agent_learn) and the CUDA Agent can replay them:
  * the DataLoader shuffle (:878)            -> `DataLoader` in the reference module's namespace is replaced by a loader
                                                that walks a recorded permutation per epoch (same batching, drop_last=False);
  * the reward-conditioning Bernoulli (:501) -> read back through a forward hook on `model.reward_dropout`;
  * the three randn draws of evolve_         -> the RNG state is rewound after each call and the draws repeated in the
                                                reference's order and shapes (evolution.py:113, :133, :173).
Dropout inside the Decoder is 0 (`agent_kwargs=dict(dropout=0.)`, Agent kwarg :684): learn-side parity is defined at
dropout 0.  Third-party arithmetic (Decoder, HL-Gauss, AdoptAtan2, EMA) comes from oracle/shims: parity
unpinned there.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from .make_golden import MODEL_CASES, build_learner, _np

# (golden name, model case, evolutionary on/off comes with the case, extra Learner kwargs)
LEARN_CASES = dict(
    learn_c1=dict(case='c1', extra=dict()),                                           # discrete, no genes (README shape)
    learn_c2s=dict(case='c2s', extra=dict(evolve_every=1, evolve_after_step=0)),      # evolutionary, gates + value residual
    learn_c4c=dict(case='c4c', extra=dict()),                                         # continuous actions, squashed
)
EPOCHS, BATCH, EPISODES, UPDATES = 2, 2, 4, 2
EMA_KW = dict(update_after_step=2, update_every=2, update_model_with_ema_every=5)      # copy, lerp and copy-back all inside 8 steps


class _FixedOrderLoader:
    """stands in for DataLoader(dataset, batch_size, shuffle=True): one recorded permutation per epoch."""

    def __init__(self, dataset, perms, batch_size):
        self.dataset, self.perms, self.batch_size, self.epoch = dataset, perms, batch_size, 0

    def __iter__(self):
        perm = self.perms[self.epoch]
        self.epoch += 1
        for lo in range(0, len(perm), self.batch_size):
            idx = perm[lo:lo + self.batch_size]
            yield tuple(t[idx] for t in self.dataset.tensors)


def _memories(ref, kw, seed):
    S, A, T = kw['state_dim'], kw['num_actions'], kw['max_timesteps']
    cont = kw.get('continuous_actions', False)
    g = torch.Generator().manual_seed(seed)
    lens = [T, max(T - 3, 2), max(T // 2, 2), T][:EPISODES]
    lo, hi = kw['reward_range']
    episodes = []
    for e, n in enumerate(lens):
        ep = []
        for t in range(n):
            action = torch.tanh(torch.randn(A, generator=g)) * 0.9 if cont else torch.randint(0, A, (), generator=g)
            logp = -torch.rand((A,) if cont else (), generator=g) * 1.5
            ep.append(ref.Memory(torch.randn(S, generator=g), action, logp, torch.randn((), generator=g) * (hi - lo) * 0.1,
                                 torch.tensor(bool(e == 1 and t == n - 1)),            # one terminated episode
                                 torch.randn(100, generator=g) * 0.5))
        episodes.append(ep)
    return episodes, torch.tensor(lens)


def make_case(ref, name, spec, seed):
    kw = dict(MODEL_CASES[spec['case']])
    learner = build_learner(ref, spec['case'], seed, epochs=EPOCHS, agent_kwargs=dict(dropout=0., ema_kwargs=dict(EMA_KW)), **spec['extra'])
    agent = learner.agent
    agent.batch_size = BATCH
    evo = bool(kw['evolutionary'])
    g = torch.Generator().manual_seed(seed + 7)
    with torch.no_grad():                                           # non-trivial normaliser statistics
        agent.rsnorm.running_mean.copy_(torch.randn(agent.rsnorm.running_mean.shape, generator=g) * 0.3)
        agent.rsnorm.running_variance.copy_(torch.rand(agent.rsnorm.running_variance.shape, generator=g) + 0.5)
        agent.rsnorm.step.fill_(3)
        if evo:
            agent.step.fill_(1)                                     # evolve_ fires when step > evolve_after_step (:1009-1013)
    memories, lens = _memories(ref, kw, seed + 3)
    gene_ids = torch.tensor([0, 1, 2, 0][:EPISODES]) if evo else torch.zeros(EPISODES, dtype=torch.long)
    fitnesses = torch.randn(agent.gene_pool.num_genes, generator=g) if evo else None

    out = {'config_json': np.array(json.dumps(dict(case=spec['case'], learner=kw, extra=spec['extra'], epochs=EPOCHS, batch_size=BATCH,
                                                   ema_kwargs=EMA_KW, updates=UPDATES, agent_step0=int(agent.step.item()))))}
    for k, v in agent.model.state_dict().items():
        out[f'sd0/{k}'] = _np(v)
    out.update({'in/rs_mean': _np(agent.rsnorm.running_mean), 'in/rs_var': _np(agent.rsnorm.running_variance), 'in/rs_step': _np(agent.rsnorm.step),
                'in/lens': _np(lens), 'in/gene_ids': _np(gene_ids)})
    if evo:
        out['in/genes'] = _np(agent.gene_pool.genes)
        out['in/fitnesses'] = _np(fitnesses)
    # the episodes as padded tensors (what pad_sequence makes of them, :822-837)
    pad = lambda i: torch.nn.utils.rnn.pad_sequence([torch.stack([m[i] for m in ep]) for ep in memories], batch_first=True)
    for i, field in enumerate(('states', 'actions', 'log_probs', 'rewards', 'is_boundary', 'value_logits')):
        out[f'in/{field}'] = _np(pad(i))

    # recorders
    flags, noises = [], []
    hook = agent.model.reward_dropout.register_forward_hook(lambda mod, inp, res: flags.append(float(res > 0.)))
    if evo:
        pool, orig_evolve = agent.gene_pool, agent.gene_pool.evolve_
        I, gpi, s, G = pool.num_islands, pool.num_genes_per_island, pool.num_selected, pool.dim_gene
        c = gpi - s

        def recording_evolve(fit, *a, **k):
            before = torch.get_rng_state()
            res = orig_evolve(fit, *a, **k)
            after = torch.get_rng_state()
            torch.set_rng_state(before)                             # repeat the reference's draws (evolution.py:113,133,173)
            noises.append((torch.randn((I, c, s)), torch.randn((I, c, G)), torch.randn((I, gpi - 1, G))))
            torch.set_rng_state(after)
            return res
        pool.evolve_ = recording_evolve
    real_loader = ref.DataLoader
    try:
        for u in range(UPDATES):
            perms = [torch.randperm(EPISODES, generator=g) for _ in range(EPOCHS)]
            ref.DataLoader = lambda ds, batch_size, shuffle, _p=perms: _FixedOrderLoader(ds, _p, batch_size)
            n_flags, n_noise = len(flags), len(noises)
            torch.manual_seed(seed + 100 + u)
            agent.learn(memories, lens, gene_ids, fitnesses)
            tag = f'u{u}'
            out[f'{tag}/perms'] = np.stack([_np(p) for p in perms])
            out[f'{tag}/flags'] = np.array(flags[n_flags:], dtype=np.float32)
            for j, (nt, nc, nm) in enumerate(noises[n_noise:]):
                out[f'{tag}/noise{j}/tourn'], out[f'{tag}/noise{j}/cross'], out[f'{tag}/noise{j}/mut'] = _np(nt), _np(nc), _np(nm)
            out[f'{tag}/num_evolves'] = np.array(len(noises) - n_noise)
            last = u == UPDATES - 1                                   # fixtures stay small: every tensor after the last update,
            for k, v in agent.model.state_dict().items():             # the small ones (norm gains, biases, embeddings) after each
                if last or v.numel() <= 512:
                    out[f'{tag}/sd/{k}'] = _np(v)
            for k, v in agent.ema_model.ema_model.state_dict().items():
                if last or v.numel() <= 512:
                    out[f'{tag}/ema/{k}'] = _np(v)
            out[f'{tag}/rs_mean'], out[f'{tag}/rs_var'] = _np(agent.rsnorm.running_mean), _np(agent.rsnorm.running_variance)
            out[f'{tag}/rs_step'], out[f'{tag}/agent_step'] = _np(agent.rsnorm.step), _np(agent.step)
            if evo:
                out[f'{tag}/genes'] = _np(agent.gene_pool.genes)
            logs = agent.accelerator.last_logs                        # the last minibatch's logged means (:996-1019)
            for k in ('actor_loss', 'critic_loss', 'autoreg_loss', 'pred_done_loss'):
                out[f'{tag}/logs/{k}'] = _np(logs[k])
    finally:
        ref.DataLoader = real_loader
        hook.remove()
    return out


def make(ref, out_dir):
    for i, (name, spec) in enumerate(LEARN_CASES.items()):
        np.savez(os.path.join(out_dir, f'{name}.npz'), **make_case(ref, name, spec, seed=500 + i))
