"""CPU: pin oracle/restated.py against the vectors produced by the reference's own files
(oracle/make_golden.py -> tests/golden).  Tolerances are fp32 re-association level."""
import json

import pytest
import torch

from oracle import restated as R
from helpers import golden_model, load_golden, split, rel_err

CASES = ('c1', 'c2s', 'c3s', 'c4c', 'c5c')
TOL = 2e-5


@pytest.mark.parametrize('case', CASES)
def test_sequence_forward(case):
    z, kw, cfg, sd = golden_model(case)
    i, seq = split(z, 'in'), split(z, 'seq')
    T = i['states'].shape[1]
    mask = torch.arange(T)[None] < i['lens'][:, None]
    ra, va, sp, dn, _ = R.wm_forward(sd, cfg, i['states'], i['prev_actions'], i['prev_rewards'], 1.,
                                     gene=i.get('gene'), next_actions=i['actions'], mask=mask)
    m3 = mask[..., None]
    assert rel_err(ra * m3, seq['raw_actions'] * m3) < TOL
    assert rel_err(va * m3, seq['values'] * m3) < TOL
    assert rel_err(sp * m3, seq['state_pred'] * m3) < TOL
    assert rel_err(dn * mask, seq['dones'] * mask) < TOL


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('batched', (False, True))
def test_cached_rollout_steps(case, batched):
    """token-by-token with a KV cache == the reference's forward_eval loop (:1269-1275)."""
    z, kw, cfg, sd = golden_model(case)
    i, step = split(z, 'in'), split(z, 'step')
    T = i['states'].shape[1]
    sl = slice(0, 3) if batched else slice(0, 1)
    cache, ras, vas = None, [], []
    for tt in range(T):
        ra, va, _, _, cache = R.wm_forward(sd, cfg, i['states'][sl, tt:tt + 1], i['prev_actions'][sl, tt:tt + 1],
                                           i['prev_rewards'][sl, tt:tt + 1], 1.,
                                           gene=None if 'gene' not in i else i['gene'][sl], cache=cache)
        ras.append(ra[0, 0]); vas.append(va[0, 0])
    assert rel_err(torch.stack(ras), step['raw_actions']) < TOL
    assert rel_err(torch.stack(vas), step['values']) < TOL
    if cfg.continuous:
        mean, var, _ = R.continuous_stats(step['raw_actions'])
        assert rel_err(torch.stack((mean, var)), step['mean_variance']) < 1e-6
        assert rel_err(R.continuous_log_prob(step['raw_actions'], i['actions'][0], cfg.squash), step['log_prob']) < 1e-5
        return
    probs, logp, ent = R.discrete_stats(step['raw_actions'])
    assert rel_err(probs, step['probs']) < 1e-6
    assert rel_err(logp.gather(-1, i['actions'][0][:, None])[:, 0], step['log_prob']) < 1e-6
    assert rel_err(ent, step['entropy']) < 1e-6


def test_rollout_positions_true_matches_sequence():
    """with true positions the cached path must equal the full-sequence path (self-consistency)."""
    z, kw, cfg, sd = golden_model('c2s')
    cfg.rollout_positions = 'true'
    i, seq = split(z, 'in'), split(z, 'seq')
    T = i['states'].shape[1]
    cache, ras = None, []
    for tt in range(T):
        ra, *_, cache = R.wm_forward(sd, cfg, i['states'][:1, tt:tt + 1], i['prev_actions'][:1, tt:tt + 1],
                                     i['prev_rewards'][:1, tt:tt + 1], 1., gene=i['gene'][:1], cache=cache)
        ras.append(ra[0, 0])
    assert rel_err(torch.stack(ras), seq['raw_actions'][0]) < TOL


@pytest.mark.parametrize('case', CASES)
def test_losses_and_head_gradients(case):
    z, kw, cfg, sd = golden_model(case)
    i, seq, L = split(z, 'in'), split(z, 'seq'), split(z, 'loss')
    T = i['states'].shape[1]
    mask = torch.arange(T)[None] < i['lens'][:, None]
    cfg.actor_loss_weight, cfg.critic_loss_weight, cfg.autoregressive_loss_weight = (float(w) for w in L['weights'])
    assert rel_err(R.hl_decode(L['old_values_logits'], cfg), L['old_values_scalar']) < 1e-6
    assert rel_err(R.hl_decode(seq['values'], cfg), L['scalar_values']) < 1e-6
    ra, va, sp, dn = (seq[k].clone().requires_grad_() for k in ('raw_actions', 'values', 'state_pred', 'dones'))
    real = torch.cat((i['states'], i['prev_rewards'][..., None]), dim=-1)
    al = R.actor_loss(ra, i['actions'], L['old_log_probs'], L['returns'], L['old_values_scalar'], mask, cfg)
    cl = R.critic_loss(va, L['returns'], L['old_values_scalar'], cfg)
    assert rel_err(al, L['actor']) < TOL
    assert rel_err(cl, L['critic']) < TOL
    loss, logs = R.total_loss(ra, va, sp, dn, real, i['actions'], L['old_log_probs'], L['returns'],
                              L['old_values_scalar'], L['dones'], mask, cfg)
    assert rel_err(loss, L['total']) < TOL
    loss.backward()
    assert rel_err(ra.grad, L['g_raw_actions']) < TOL
    assert rel_err(va.grad, L['g_values']) < TOL
    assert rel_err(sp.grad, L['g_state_pred']) < TOL
    assert rel_err(dn.grad, L['g_dones']) < TOL


def test_gae():
    z = load_golden('gae.npz')
    for name in 'abc':
        g = split(z, name)
        out = R.calc_gae(g['rewards'], g['values'], ~g['boundary'])
        assert rel_err(out, g['returns']) < 1e-6


def test_evolve_selection_exact_genes_1ulp():
    z = load_golden('evolve.npz')
    names = sorted({k.split('/')[0] for k in z})
    assert len(names) == 4
    for name in names:
        pc = R.PoolConfig(**json.loads(str(z[f'{name}/config_json'])))
        its = sorted({k.split('/')[1] for k in z if k.startswith(name + '/') and k.split('/')[1].isdigit()})
        for it in its:
            g = split(z, f'{name}/{it}')
            genes, sel = R.evolve(g['genes_in'], g['fitness'], g['noise_tourn'], g['noise_cross'], g['noise_mut'],
                                  pc, step=int(g['step_before']))
            assert torch.equal(sel, g['selected']), (name, it)
            # unit-norm rows: 1 ulp of the lerp's fused/unfused multiply-add (see restated._torch_lerp)
            assert (genes - g['genes_out']).abs().max() < 1e-7, (name, it)


def test_evolve_migration_type_error_is_kept():
    """(num_selected-1)*0.1 > 1 with >1 island raises in the reference; so do we."""
    pc = R.PoolConfig(dim=4, num_genes_per_island=40, num_selected=31, tournament_size=2, num_islands=2)
    I, g, s, G, c = 2, 40, 31, 4, 9
    with pytest.raises(TypeError):
        R.evolve(torch.randn(80, 4), torch.randn(80), torch.randn(I, c, s), torch.randn(I, c, G),
                 torch.randn(I, g - 1, G), pc, step=9)


def test_rsnorm():
    z = load_golden('rsnorm.npz')
    mean, var, step = torch.zeros(6), torch.ones(6), 1
    for it in range(3):
        g = split(z, str(it))
        assert rel_err(R.rsnorm_eval(g['x'], mean, var), g['normed']) < 1e-6
        mean, var, step = R.rsnorm_update(g['x'], mean, var, step)
        assert rel_err(mean, g['mean']) < 1e-6 and rel_err(var + 1, g['var'] + 1) < 1e-6 and step == int(g['step'])


def test_sample_inverse_cdf_edges():
    p = torch.tensor([[0.1, 0.2, 0.3, 0.4]])
    for u, a in ((0.0, 0), (0.0999, 0), (0.1001, 1), (0.5999, 2), (0.61, 3), (0.99999, 3)):
        assert R.sample_inverse_cdf(p, torch.tensor([u])).item() == a
    # u beyond the accumulated mass (rounding) falls into the last action
    assert R.sample_inverse_cdf(torch.tensor([[0.5, 0.49999]]), torch.tensor([0.999999])).item() == 1


# ---- Agent.learn as a whole (x_transformers_rl.py:808-1023), goldens from the reference's own learn() --------------------

@pytest.mark.parametrize('name', ('c1', 'c2s', 'c4c'))
def test_agent_learn_restatement_matches_reference(name):
    """two consecutive learning updates (2 epochs x 2 minibatches each, ragged episodes, one terminated): parameters, EMA copy,
    RSNorm statistics, genes and the logged losses of oracle/restated.agent_learn == the reference's Agent.learn, fed the
    shuffle / reward-dropout flags / evolve noise the reference drew."""
    from helpers import learn_golden, learn_golden_noises
    z, meta, cfg, sd0, data = learn_golden(name)
    kw = meta['learner']
    sd = {k: v.clone().requires_grad_() for k, v in sd0.items()}
    ema_sd = {k: v.clone() for k, v in sd0.items()}
    i = split(z, 'in')
    rs = dict(mean=i['rs_mean'].clone(), var=i['rs_var'].clone(), step=int(i['rs_step']))
    evo = bool(kw.get('evolutionary'))
    genes = i['genes'].clone() if evo else None
    pool = R.PoolConfig(**kw['latent_gene_pool']) if evo else None
    fit = i['fitnesses'] if evo else None
    opt_state, ema_state = {}, R.EmaState(beta=0.9, **meta['ema_kwargs'])
    ema_state.initted = False
    agent_step, pool_step = meta['agent_step0'], 0
    extra = meta['extra']
    for u in range(meta['updates']):
        g = split(z, f'u{u}')
        logs, _, pool_step = R.agent_learn(
            sd, ema_sd, cfg, data, rs, opt_state, ema_state, [p for p in g['perms']], g['flags'].tolist(),
            batch_size=meta['batch_size'], genes=genes, pool=pool, pool_step=pool_step, fitnesses=fit,
            noises=learn_golden_noises(z, u), agent_step=agent_step, evolve_every=extra.get('evolve_every', 10),
            evolve_after_step=extra.get('evolve_after_step', 20))
        agent_step += 1
        assert int(g['num_evolves']) == (4 if evo else 0)
        want_sd, want_ema = split(z, f'u{u}/sd'), split(z, f'u{u}/ema')
        assert len(want_sd) >= 5
        for k, v in want_sd.items():
            if k in sd:
                assert rel_err(sd[k].detach(), v) < 5e-5, (u, k)
        for k, v in want_ema.items():
            if k in ema_sd:
                assert rel_err(ema_sd[k], v) < 5e-5, (u, k)
        assert rel_err(rs['mean'], g['rs_mean']) < 1e-5 and rel_err(rs['var'], g['rs_var']) < 1e-5
        assert rs['step'] == int(g['rs_step']) and agent_step == int(g['agent_step'])
        if evo:
            assert rel_err(genes, g['genes']) < 1e-5
        for k in ('actor_loss', 'critic_loss', 'autoreg_loss', 'pred_done_loss'):
            assert rel_err(logs[k], g[f'logs/{k}']) < 5e-5, (u, k)
    # the update moved the weights by more than the tolerance asks of them
    moved = max(rel_err(sd[k].detach(), sd0[k]) for k in sd0)
    assert moved > 1e-3


def test_hl_gauss_reduction_u3_scalar_alternative_is_different():
    """the reference calls hl_gauss(values, returns) without `reduction=` and then treats the result
    per element (torch.min / torch.where with (b, n) conditions, `[mask]`, :461-475, :973).  Some versions of hl-gauss-pytorch
    return the MEAN scalar by default, which would still broadcast.  The oracle and the kernels implement the per-element
    reading; this test documents what the scalar reading would compute instead and that the two differ."""
    cfg = R.Config(state_dim=4, num_actions=2, reward_range=(-3., 3.))
    g = torch.Generator().manual_seed(0)
    b, n = 3, 7
    values, returns, old = torch.randn(b, n, cfg.num_bins, generator=g), torch.randn(b, n, generator=g), torch.randn(b, n, generator=g) * 0.3
    per_elem = R.critic_loss(values, returns, old, cfg)                                   # (b, n): the implemented reading
    assert per_elem.shape == (b, n)
    # the scalar alternative: both cross entropies collapse to their batch means before the clip logic
    ce = R.hl_cross_entropy(values, returns, cfg).mean()
    ce_clip = R.hl_cross_entropy(values, returns.clamp(-cfg.value_clip, cfg.value_clip), cfg).mean()
    v, lo, hi = R.hl_decode(values, cfg), old - cfg.value_clip, old + cfg.value_clip
    zero = ((returns < v) & (v < lo)) | ((hi < v) & (v < returns))
    scalar_alt = torch.where(zero, torch.zeros(()), torch.minimum(ce, ce_clip).expand(b, n))
    assert scalar_alt.shape == (b, n)
    assert scalar_alt[~zero].unique().numel() <= 1                                        # one number wherever the clip does not zero it
    assert per_elem[~zero].unique().numel() > 1 and not torch.allclose(per_elem, scalar_alt)
    # both readings agree on WHICH elements the clip zeroes, and on the mean of the unclipped cross entropy
    assert torch.equal(per_elem == 0, scalar_alt == 0) or zero.sum() == 0
    assert rel_err(R.hl_cross_entropy(values, returns, cfg).mean(), ce) < 1e-6
