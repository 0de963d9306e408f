"""GPU parity: every CUDA entry point (through the C ABI) against the oracle on the same inputs and
against the committed golden vectors.  Tolerances follow BASELINE.json north_star: GAE / returns /
losses 1e-5 relative in fp32 (1e-2 with bf16 I/O); actions and gene selection bit-exact."""
import json

import pytest
import torch

from oracle import restated as R
from helpers import golden_model, load_golden, split, rel_err, xtrl, model_config, random_state_dict

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-5


def cu(t):
    return t.to(DEV).contiguous()


# ---------------------------------------------------------------------------------------------
# (2) GAE
# ---------------------------------------------------------------------------------------------

def test_gae_golden_both_layouts():
    ops = xtrl().ops
    z = load_golden('gae.npz')
    for name in 'abc':
        g = split(z, name)
        em = ops.gae(cu(g['rewards']), cu(g['values']), cu(g['boundary'])).cpu()
        tm = ops.gae(cu(g['rewards'].T), cu(g['values'].T), cu(g['boundary'].T), time_major=True).cpu().T
        assert rel_err(em, g['returns']) < TOL, name
        assert rel_err(tm, g['returns']) < TOL, name


@pytest.mark.parametrize('E,T', [(1, 1), (3, 31), (5, 32), (7, 33), (33, 129), (64, 500), (300, 513), (4096, 500)])
def test_gae_vs_oracle(E, T):
    ops = xtrl().ops
    g = torch.Generator().manual_seed(E * 1000 + T)
    r, v = torch.randn(E, T, generator=g), torch.randn(E, T, generator=g)
    bd = torch.rand(E, T, generator=g) < 0.05
    want = R.calc_gae(r.double(), v.double(), ~bd).float()
    em = ops.gae(cu(r), cu(v), cu(bd)).cpu()
    tm = ops.gae(cu(r.T), cu(v.T), cu(bd.T), time_major=True).cpu().T
    assert rel_err(em, want) < TOL
    assert rel_err(tm, want) < TOL


def test_gae_properties_full_size():
    """BASELINE c3 size: linearity in (rewards, values) and the no-discount telescoping identity."""
    ops = xtrl().ops
    E, T = 4096, 500
    g = torch.Generator(device=DEV).manual_seed(5)
    r1, r2 = torch.randn(E, T, generator=g, device=DEV), torch.randn(E, T, generator=g, device=DEV)
    v1, v2 = torch.randn(E, T, generator=g, device=DEV), torch.randn(E, T, generator=g, device=DEV)
    bd = torch.rand(E, T, generator=g, device=DEV) < 0.02
    a, b = ops.gae(r1, v1, bd), ops.gae(r2, v2, bd)
    ab = ops.gae(r1 + 2 * r2, v1 + 2 * v2, bd)
    assert rel_err(ab.cpu(), (a + 2 * b).cpu()) < 5e-5
    # gamma = lam = 1, no boundaries: returns_t = sum_{s>=t} r_s (values telescope away)
    nb = torch.zeros_like(bd)
    ret = ops.gae(r1, v1, nb, gamma=1., lam=1.)
    want = r1.double().flip(1).cumsum(1).flip(1).float()
    assert rel_err(ret.cpu(), want.cpu()) < 5e-5
    # empty inputs are fine
    assert ops.gae(r1[:0], v1[:0], bd[:0]).shape == (0, T)


# ---------------------------------------------------------------------------------------------
# (3) fused PPO loss
# ---------------------------------------------------------------------------------------------

def _loss_case(case):
    z, kw, cfg, sd = golden_model(case)
    i, seq, Lg = split(z, 'in'), split(z, 'seq'), split(z, 'loss')
    T = i['states'].shape[1]
    mask = torch.arange(T)[None] < i['lens'][:, None]
    cfg.actor_loss_weight, cfg.critic_loss_weight, cfg.autoregressive_loss_weight = (float(w) for w in Lg['weights'])
    raw = {}
    ra, va, sp, dn, _ = R.wm_forward(sd, cfg, i['states'], i['prev_actions'], i['prev_rewards'], 1.,
                                     gene=i.get('gene'), next_actions=i['actions'], mask=mask, raw_out=raw)
    return cfg, i, seq, Lg, mask, ra, va, raw


@pytest.mark.parametrize('case', ('c1', 'c2s', 'c3s', 'c4c', 'c5c'))
def test_ppo_loss_golden(case):
    ops = xtrl().ops
    cfg, i, seq, Lg, mask, ra, va, raw = _loss_case(case)
    real = torch.cat((i['states'], i['prev_rewards'][..., None]), dim=-1)
    lc = ops.LossConfig(num_actions=cfg.num_actions, num_bins=cfg.num_bins, state_dim=cfg.state_dim,
                        reward_range=cfg.reward_range, eps_clip=cfg.eps_clip, value_clip=cfg.value_clip,
                        entropy_weight=cfg.entropy_weight, actor_weight=cfg.actor_loss_weight,
                        critic_weight=cfg.critic_loss_weight, autoreg_weight=cfg.autoregressive_loss_weight,
                        continuous=cfg.continuous, squash=cfg.squash)
    # the head outputs of the reference run (golden) feed the kernel; raw world-model heads from the oracle
    scal, grads, per = ops.ppo_loss_fwd_bwd(
        lc, cu(seq['raw_actions']), cu(seq['values']), cu(raw['state_pred_raw'].detach()), cu(raw['done_logit'].detach()),
        cu(i['actions'].float() if cfg.continuous else i['actions'].int()), cu(Lg['old_log_probs']), cu(Lg['returns']),
        cu(Lg['old_values_scalar']), cu(real), cu(Lg['dones']), cu(i['lens'].int()), per_step=True)
    scal, per = scal.cpu(), per.cpu()
    assert rel_err(per[..., 0], Lg['actor']) < TOL
    assert rel_err(per[..., 1], Lg['critic']) < TOL
    assert rel_err(scal[0], Lg['total']) < TOL
    assert rel_err(scal[1], Lg['actor'].mean()) < TOL and rel_err(scal[2], Lg['critic'].mean()) < TOL
    assert rel_err(scal[3], Lg['autoreg_masked'].mean()) < TOL and rel_err(scal[4], Lg['done_masked'].mean()) < TOL
    assert scal[5].item() == mask.sum().item()
    # gradients w.r.t. the actor / critic logits are pinned by the reference's own autograd
    assert rel_err(grads['raw_actions'].cpu(), Lg['g_raw_actions']) < TOL
    assert rel_err(grads['value_logits'].cpu(), Lg['g_values']) < TOL
    # world-model heads: chain the reference's gradients (w.r.t. mean/var and sigmoid) to the raw outputs
    spr = raw['state_pred_raw'].detach().unflatten(-1, (-1, 2))
    th = torch.tanh(spr[..., 1] / 3)
    var = torch.exp(th * 3)
    g_mean, g_var = Lg['g_state_pred'][0], Lg['g_state_pred'][1]
    want_sp = torch.stack((g_mean, g_var * var * (1 - th * th)), dim=-1).flatten(-2)
    assert rel_err(grads['state_pred_raw'].cpu(), want_sp) < TOL
    p = torch.sigmoid(raw['done_logit'].detach())
    assert rel_err(grads['done_logit'].cpu(), Lg['g_dones'] * p * (1 - p)) < TOL


@pytest.mark.parametrize('b,T,A,K,S', [(2, 7, 2, 100, 5), (8, 500, 4, 100, 8), (3, 40, 18, 51, 3), (4, 33, 6, 128, 12)])
@pytest.mark.parametrize('bf16', (False, True))
def test_ppo_loss_vs_oracle_autograd(b, T, A, K, S, bf16):
    ops = xtrl().ops
    g = torch.Generator().manual_seed(b * 31 + T)
    cfg = R.Config(state_dim=S, num_actions=A, reward_range=(-3., 5.), num_bins=K, entropy_weight=0.02,
                   actor_loss_weight=0.7, critic_loss_weight=1.3, autoregressive_loss_weight=0.9)
    rn = lambda *s: torch.randn(*s, generator=g)
    ra, va, spr, dl = rn(b, T, A), rn(b, T, K), rn(b, T, 2 * (S + 1)) * 2, rn(b, T) * 2
    if bf16:
        ra, va, spr, dl = (t.bfloat16().float() for t in (ra, va, spr, dl))
    actions = torch.randint(0, A, (b, T), generator=g)
    old_lp, returns, old_v = -torch.rand(b, T, generator=g) * 2, rn(b, T) * 2, rn(b, T)
    real, dones = rn(b, T, S + 1), torch.rand(b, T, generator=g) < 0.1
    lens = torch.randint(1, T + 1, (b,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None] < lens[:, None]
    leaves = [t.clone().double().requires_grad_() for t in (ra, va, spr, dl)]
    pr = leaves[2].unflatten(-1, (-1, 2))
    sp = torch.stack((pr[..., 0], torch.exp(torch.tanh(pr[..., 1] / 3) * 3)))
    loss, logs = R.total_loss(leaves[0], leaves[1], sp, torch.sigmoid(leaves[3]), real.double(), actions, old_lp.double(),
                              returns.double(), old_v.double(), dones, mask, cfg)
    loss.backward()
    lc = ops.LossConfig(num_actions=A, num_bins=K, state_dim=S, reward_range=cfg.reward_range,
                        entropy_weight=cfg.entropy_weight, actor_weight=0.7, critic_weight=1.3, autoreg_weight=0.9)
    dt = torch.bfloat16 if bf16 else torch.float32
    scal, grads = ops.ppo_loss_fwd_bwd(lc, cu(ra.to(dt)), cu(va.to(dt)), cu(spr.to(dt)), cu(dl.to(dt)), cu(actions.int()),
                                       cu(old_lp), cu(returns), cu(old_v), cu(real), cu(dones), cu(lens.int()))
    tol = 1e-2 if bf16 else 2e-5      # bf16: only the gradient STORE is rounded (inputs were made bf16-exact)
    assert rel_err(scal[0].cpu(), loss.detach()) < 2e-5
    for name, leaf in zip(('raw_actions', 'value_logits', 'state_pred_raw', 'done_logit'), leaves):
        assert rel_err(grads[name].float().cpu(), leaf.grad) < tol, name
    assert rel_err(scal[1].cpu(), logs['actor_loss'].detach()) < 2e-5
    assert rel_err(scal[2].cpu(), logs['critic_loss'].detach()) < 2e-5


# ---------------------------------------------------------------------------------------------
# (4) gene pool
# ---------------------------------------------------------------------------------------------

def test_evolve_golden_selection_bit_exact():
    ops = xtrl().ops
    z = load_golden('evolve.npz')
    names = sorted({k.split('/')[0] for k in z})
    for name in names:
        kw = json.loads(str(z[f'{name}/config_json']))
        pc = R.PoolConfig(**kw)
        its = sorted({k.split('/')[1] for k in z if k.startswith(name + '/') and k.split('/')[1].isdigit()})
        for it in its:
            g = split(z, f'{name}/{it}')
            step = int(g['step_before'])
            migrate = (step + 1) % pc.migrate_genes_every == 0 and pc.num_islands > 1 and pc.num_frac_migrate > 0
            n_can = pc.num_selected - (1 if pc.num_elites > 0 else 0)
            num_migrate = max(1, n_can * pc.num_frac_migrate)
            genes = cu(g['genes_in'].clone())
            sel = ops.evolve_(genes, cu(g['fitness']), cu(g['noise_tourn']), cu(g['noise_cross']), cu(g['noise_mut']),
                              num_islands=pc.num_islands, num_genes_per_island=pc.num_genes_per_island,
                              num_selected=pc.num_selected, tournament_size=pc.tournament_size,
                              num_elites=pc.num_elites, mutation_std_dev=pc.mutation_std_dev, migrate=migrate,
                              num_migrate=int(num_migrate))
            assert torch.equal(sel.cpu(), g['selected']), (name, it)
            assert (genes.cpu() - g['genes_out']).abs().max() < 2e-7, (name, it)    # unit rows: 1-2 ulp


def test_evolve_scaled_pool_vs_oracle():
    """bandwidth-sized pool (4096 islands x 16 genes x 128)."""
    ops = xtrl().ops
    pc = R.PoolConfig(dim=128, num_genes_per_island=16, num_selected=4, tournament_size=2, num_islands=4096)
    I, g, s, G, c = 4096, 16, 4, 128, 12
    gen = torch.Generator().manual_seed(9)
    genes = torch.nn.functional.normalize(torch.randn(I * g, G, generator=gen), dim=-1)
    fit = torch.randn(I * g, generator=gen)
    nt, nc, nm = torch.randn(I, c, s, generator=gen), torch.randn(I, c, G, generator=gen), torch.randn(I, g - 1, G, generator=gen)
    want, want_sel = R.evolve(genes, fit, nt, nc, nm, pc, step=0)
    got = cu(genes.clone())
    sel = ops.evolve_(got, cu(fit), cu(nt), cu(nc), cu(nm), num_islands=I, num_genes_per_island=g, num_selected=s,
                      tournament_size=2)
    assert torch.equal(sel.cpu(), want_sel)
    assert (got.cpu() - want).abs().max() < 2e-7
    assert rel_err(got.norm(dim=-1).cpu(), torch.ones(I * g)) < 1e-6     # idempotent normalisation


# ---------------------------------------------------------------------------------------------
# (1) rollout step
# ---------------------------------------------------------------------------------------------

def _rollout_oracle(sd, cfg, states, prev_actions, prev_rewards, genes_n):
    """token-by-token oracle rollout on given inputs: returns per-step logits (T,B,A), (T,B,K)."""
    T = states.shape[1]
    cache, ras, vas = None, [], []
    for t in range(T):
        ra, va, _, _, cache = R.wm_forward(sd, cfg, states[:, t:t + 1], prev_actions[:, t:t + 1],
                                           prev_rewards[:, t:t + 1], 1., gene=genes_n, cache=cache)
        ras.append(ra[:, 0]); vas.append(va[:, 0])
    return torch.stack(ras), torch.stack(vas)


def _rollout_kernel(sd, cfg, states, prev_actions, prev_rewards, genes=None, gene_ids=None, **mc_extra):
    ops = xtrl().ops
    B, T, S = states.shape
    mc = model_config(cfg, T, **mc_extra)
    st = ops.WorldModelStepper(mc, B, DEV)
    st.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    mean, var = torch.zeros(S + 1, device=DEV), torch.ones(S + 1, device=DEV)
    ras, vas, vals = [], [], []
    for t in range(T):
        pa = cu(prev_actions[:, t].float()) if cfg.continuous else cu(prev_actions[:, t].int())
        fa = cu(torch.zeros(B, cfg.num_actions)) if cfg.continuous else cu(torch.zeros(B, dtype=torch.int32))
        out = st.step(cu(states[:, t]), pa, cu(prev_rewards[:, t]), mean, var, forced_action=fa, want_logits=True,
                      genes=None if genes is None else cu(genes), gene_ids=None if gene_ids is None else cu(gene_ids.int()))
        ras.append(out['raw_actions'].clone()); vas.append(out['value_logits'].clone()); vals.append(out['value'].clone())
    return torch.stack(ras).cpu(), torch.stack(vas).cpu(), torch.stack(vals).cpu()


@pytest.mark.parametrize('case', ('c1', 'c2s', 'c3s', 'c4c', 'c5c'))
def test_wm_step_golden_rollout(case):
    """the reference's own forward_eval loop (golden `step/*`, env 0) and the oracle for all 3 envs."""
    z, kw, cfg, sd = golden_model(case)
    i, step = split(z, 'in'), split(z, 'step')
    genes = i.get('genes')
    ra, va, val = _rollout_kernel(sd, cfg, i['states'], i['prev_actions'], i['prev_rewards'],
                                  genes=genes, gene_ids=i.get('gene_ids'))
    assert rel_err(ra[:, 0], step['raw_actions']) < 2e-5
    assert rel_err(va[:, 0], step['values']) < 2e-5
    want_ra, want_va = _rollout_oracle(sd, cfg, i['states'], i['prev_actions'], i['prev_rewards'], i.get('gene'))
    assert rel_err(ra, want_ra) < 2e-5 and rel_err(va, want_va) < 2e-5
    assert rel_err(val, R.hl_decode(want_va, cfg)) < 2e-5


@pytest.mark.parametrize('kw', [
    dict(heads=8, depth=4, state_dim=8, num_actions=4),                                      # BASELINE c3 shape
    dict(heads=4, depth=4, state_dim=8, num_actions=4, gate_values=True, value_residual=True,
         learned_mix=True, evolutionary=True, dim_gene=32),                                  # c2 (train_lander.py)
    dict(heads=8, depth=6, state_dim=8, num_actions=4),                                      # c5 shape
    dict(heads=3, depth=2, state_dim=5, num_actions=3, num_bins=51, value_residual=True),
])
@pytest.mark.parametrize('true_pos', (False, True))
def test_wm_step_vs_oracle_random_model(kw, true_pos):
    cfg = R.Config(reward_range=(-5., 5.), **kw)
    cfg.rollout_positions = 'true' if true_pos else 'reference'
    sd = random_state_dict(cfg, seed=1)
    B, T = 37, 70                       # B not a multiple of the CTA tile; T crosses two 32-key blocks
    g = torch.Generator().manual_seed(2)
    states, prev_r = torch.randn(B, T, cfg.state_dim, generator=g), torch.randn(B, T, generator=g)
    prev_a = torch.randint(-1, cfg.num_actions, (B, T), generator=g)
    genes = gene_ids = gene_n = None
    if cfg.evolutionary:
        genes = torch.randn(5, cfg.dim_gene, generator=g)
        gene_ids = torch.randint(0, 5, (B,), generator=g)
        gene_n = torch.nn.functional.normalize(genes[gene_ids], dim=-1)
    ra, va, val = _rollout_kernel(sd, cfg, states, prev_a, prev_r, genes=genes, gene_ids=gene_ids, rotary_true_pos=true_pos)
    want_ra, want_va = _rollout_oracle(sd, cfg, states, prev_a, prev_r, gene_n)
    assert rel_err(ra, want_ra) < 3e-5
    assert rel_err(va, want_va) < 3e-5
    assert rel_err(val, R.hl_decode(want_va, cfg)) < 3e-5


def test_wm_step_sampling_and_rsnorm():
    """inverse-CDF sampling is bit-exact outside the tolerance margin; log-prob / value match; RSNorm applied."""
    ops = xtrl().ops
    cfg = R.Config(state_dim=8, num_actions=4, heads=4, depth=1, reward_range=(-5., 5.))
    sd = random_state_dict(cfg, seed=4, scale=0.5)
    B = 4096
    g = torch.Generator().manual_seed(6)
    state, prev_r = torch.randn(B, 8, generator=g) * 3 + 1, torch.randn(B, generator=g)
    prev_a = torch.randint(-1, 4, (B,), generator=g)
    mean, var = torch.randn(9, generator=g), torch.rand(9, generator=g) + 0.1
    var[3] = 0.                                       # variance 0 -> the 1e-5 clamp divides
    u = torch.rand(B, generator=g)
    st = ops.WorldModelStepper(model_config(cfg, 4), B, DEV)
    st.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    out = st.step(cu(state), cu(prev_a.int()), cu(prev_r), cu(mean), cu(var), uniform=cu(u), want_logits=True)
    packed = R.rsnorm_eval(torch.cat((state, prev_r[:, None]), -1), mean, var)
    ra, va, *_ = R.wm_forward(sd, cfg, packed[:, None, :-1], prev_a[:, None], packed[:, None, -1], 1.)
    ra, va = ra[:, 0], va[:, 0]
    assert rel_err(out['raw_actions'].cpu(), ra) < 3e-5
    probs, logp, _ = R.discrete_stats(ra)
    want_a = R.sample_inverse_cdf(probs, u)
    got_a = out['action'].cpu().long()
    cdf = probs.cumsum(-1)
    margin = (cdf - u[:, None]).abs().min(dim=-1).values
    safe = margin > 1e-5
    assert safe.float().mean() > 0.99
    assert torch.equal(got_a[safe], want_a[safe])
    assert rel_err(out['log_prob'].cpu()[safe], logp.gather(-1, want_a[:, None])[:, 0][safe]) < 3e-5
    assert rel_err(out['value'].cpu(), R.hl_decode(va, cfg)) < 3e-5


@pytest.mark.parametrize('kw', [
    dict(heads=8, depth=4, state_dim=8, num_actions=4),                                      # c3
    dict(heads=4, depth=4, state_dim=8, num_actions=4, gate_values=True, value_residual=True,
         learned_mix=True, evolutionary=True, dim_gene=32),                                  # c2: NQ=260 -> 3 tiles, HIN=144
    dict(heads=8, depth=2, state_dim=8, num_actions=4, gate_values=True, value_residual=True, learned_mix=True),  # NQ=520: 5 tiles, 2 weight groups
    dict(heads=6, depth=2, state_dim=5, num_actions=3, dim=32, gate_values=True),              # hidden_dim 32: K=32, tile groups capped by TMEM
])
@pytest.mark.parametrize('kv_bf16', (False, True))
def test_wm_step_tensor_core_dense(kw, kv_bf16):
    """tcgen05 path (bf16 operands, fp32 accumulation in TMEM) against the fp32 oracle: the 1e-2 class."""
    cfg = R.Config(reward_range=(-5., 5.), **kw)
    sd = random_state_dict(cfg, seed=1)
    B, T = 37, 40
    g = torch.Generator().manual_seed(2)
    states, prev_r = torch.randn(B, T, cfg.state_dim, generator=g), torch.randn(B, T, generator=g)
    prev_a = torch.randint(-1, cfg.num_actions, (B, T), generator=g)
    genes = gene_ids = gene_n = None
    if cfg.evolutionary:
        genes = torch.randn(5, cfg.dim_gene, generator=g)
        gene_ids = torch.randint(0, 5, (B,), generator=g)
        gene_n = torch.nn.functional.normalize(genes[gene_ids], dim=-1)
    ra, va, val = _rollout_kernel(sd, cfg, states, prev_a, prev_r, genes=genes, gene_ids=gene_ids, dense_bf16=True, kv_bf16=kv_bf16)
    want_ra, want_va = _rollout_oracle(sd, cfg, states, prev_a, prev_r, gene_n)
    # measured on B200 (scripts/tc_error_probe.py): logits 0.7-1.4 % of their max, decoded value <= 0.9 % of the
    # reward range -- the bf16-operand class; the fp32 path is the 1e-5 class
    assert rel_err(ra, want_ra) < 2e-2
    assert rel_err(va, want_va) < 2e-2
    lo, hi = cfg.reward_range
    assert ((val - R.hl_decode(want_va, cfg)).abs().max() / (hi - lo)).item() < 1.5e-2
    # and it is NOT the fp32 path in disguise: bf16 rounding must be visible
    assert rel_err(ra, want_ra) > 1e-5


def test_wm_step_full_size_c3():
    """BASELINE c3 at full size (4096 envs, T=500, depth 4, heads 8).  Envs are independent, so (a) a few
    envs of the 4096-wide kernel rollout equal the oracle stepping just those envs, (b) at the LAST timestep
    (500 cached keys) the fp32 kernel, the tensor-core kernel and the bf16-KV kernel agree within their classes,
    (c) every env's cache row count advanced identically (idempotent re-run gives bit-identical outputs)."""
    ops = xtrl().ops
    cfg = R.Config(state_dim=8, num_actions=4, heads=8, depth=4, reward_range=(-5., 5.))
    sd = random_state_dict(cfg, seed=3)
    B, T = 4096, 500
    g = torch.Generator(device=DEV).manual_seed(11)
    states = torch.randn(T, B, 8, device=DEV, generator=g)
    rewards = torch.randn(T, B, device=DEV, generator=g)
    acts = torch.randint(-1, 4, (T, B), device=DEV, generator=g, dtype=torch.int32)
    mean, var = torch.zeros(9, device=DEV), torch.ones(9, device=DEV)
    pick = torch.tensor([0, 7, 2049, 4095])

    def run(**extra):
        st = ops.WorldModelStepper(model_config(cfg, T, **extra), B, DEV)
        st.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
        early, last = [], None
        for t in range(T):
            out = st.step(states[t], acts[t], rewards[t], mean, var, forced_action=acts[t].clamp(min=0), want_logits=True,
                          pos_uniform=t)
            if t < 48:
                early.append(out['raw_actions'][pick.to(DEV)].clone())
            last = out
        res = dict(early=torch.stack(early).cpu(), ra=last['raw_actions'].cpu(), vl=last['value_logits'].cpu(),
                   val=last['value'].cpu(), lp=last['log_prob'].cpu())
        del st
        torch.cuda.empty_cache()
        return res

    fp32 = run()
    # (a) against the oracle on the picked envs, first 48 steps
    cache, want = None, []
    s_cpu, r_cpu, a_cpu = states[:48, pick.to(DEV)].cpu(), rewards[:48, pick.to(DEV)].cpu(), acts[:48, pick.to(DEV)].cpu().long()
    for t in range(48):
        ra, va, _, _, cache = R.wm_forward(sd, cfg, s_cpu[t][:, None], a_cpu[t][:, None], r_cpu[t][:, None], 1., cache=cache)
        want.append(ra[:, 0])
    assert rel_err(fp32['early'], torch.stack(want)) < 3e-5
    # (c) determinism of the whole 500-step rollout
    again = run()
    assert torch.equal(again['ra'], fp32['ra']) and torch.equal(again['val'], fp32['val'])
    # (b) precision classes at t = 499
    tc = run(dense_bf16=True)
    assert rel_err(tc['ra'], fp32['ra']) < 2e-2 and ((tc['val'] - fp32['val']).abs().max() / 10.).item() < 1.5e-2
    lo = run(dense_bf16=True, kv_bf16=True)
    assert rel_err(lo['ra'], fp32['ra']) < 3e-2 and ((lo['val'] - fp32['val']).abs().max() / 10.).item() < 1.5e-2
    assert torch.isfinite(fp32['vl']).all() and torch.isfinite(lo['lp']).all()


def test_wm_step_ragged_positions_and_empty_batch():
    """per-env cache lengths: envs that restart (pos -> 0) mid-rollout while their neighbours continue, as a
    vector env with auto-reset needs; and a zero-env call is a no-op."""
    ops = xtrl().ops
    cfg = R.Config(state_dim=8, num_actions=4, heads=4, depth=2, reward_range=(-5., 5.), value_residual=True)
    sd = random_state_dict(cfg, seed=7)
    B, T = 21, 40
    g = torch.Generator().manual_seed(3)
    states, prev_r = torch.randn(B, T, 8, generator=g), torch.randn(B, T, generator=g)
    prev_a = torch.randint(-1, 4, (B, T), generator=g)
    restart_at = torch.randint(1, T, (B,), generator=g)
    restart_at[::3] = T + 1                                        # a third of the envs never restart
    st = ops.WorldModelStepper(model_config(cfg, T), B, DEV)
    st.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    mean, var = torch.zeros(9, device=DEV), torch.ones(9, device=DEV)
    got = []
    for t in range(T):
        st.pos[(restart_at == t).to(DEV)] = 0                      # episode ended: empty cache for those envs
        out = st.step(cu(states[:, t]), cu(prev_a[:, t].int()), cu(prev_r[:, t]), mean, var,
                      forced_action=cu(torch.zeros(B, dtype=torch.int32)), want_logits=True)
        got.append(out['raw_actions'].cpu().clone())
    got = torch.stack(got, 1)                                      # (B, T, A)
    for b in range(B):
        cache = None
        for t in range(T):
            if t == int(restart_at[b]):
                cache = None
            ra, _, _, _, cache = R.wm_forward(sd, cfg, states[b:b + 1, t:t + 1], prev_a[b:b + 1, t:t + 1], prev_r[b:b + 1, t:t + 1], 1., cache=cache)
            assert rel_err(got[b, t], ra[0, 0]) < 5e-5, (b, t)
    empty = ops.WorldModelStepper(model_config(cfg, T), 0, DEV)
    empty.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    out = empty.step(torch.empty(0, 8, device=DEV), torch.empty(0, dtype=torch.int32, device=DEV), torch.empty(0, device=DEV), mean, var)
    assert out['action'].shape == (0,)


def test_ppo_loss_empty_and_single_row():
    ops = xtrl().ops
    lc = ops.LossConfig(num_actions=3, num_bins=51, state_dim=4, reward_range=(-1., 1.))
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=DEV, dtype=dt)
    scal, grads = ops.ppo_loss_fwd_bwd(lc, z(0, 5, 3), z(0, 5, 51), z(0, 5, 10), z(0, 5), z(0, 5, dt=torch.int32), z(0, 5), z(0, 5),
                                       z(0, 5), z(0, 5, 5), z(0, 5, dt=torch.bool), z(0, dt=torch.int32))
    assert grads['raw_actions'].shape == (0, 5, 3) and torch.all(scal.cpu() == 0)
    # one episode of one step: the unbiased variance of a single advantage is NaN in the reference too
    g = torch.Generator().manual_seed(0)
    ra, va = torch.randn(1, 1, 3, generator=g), torch.randn(1, 1, 51, generator=g)
    scal, grads = ops.ppo_loss_fwd_bwd(lc, cu(ra), cu(va), z(1, 1, 10), z(1, 1), z(1, 1, dt=torch.int32), z(1, 1), z(1, 1) + 0.3,
                                       z(1, 1), z(1, 1, 5), z(1, 1, dt=torch.bool), torch.ones(1, device=DEV, dtype=torch.int32))
    assert scal[5].item() == 1 and torch.isnan(scal[7]).item()
    cfg = R.Config(state_dim=4, num_actions=3, reward_range=(-1., 1.), num_bins=51)
    want = R.critic_loss(va, torch.full((1, 1), 0.3), torch.zeros(1, 1), cfg)
    assert rel_err(scal[2].cpu(), want.mean()) < 1e-5


@pytest.mark.parametrize('squash', (True, False))
def test_wm_step_continuous_sampling(squash):
    """Continuous head (:224-277): sample = mean + std * normal (tanh-squashed), per-dim log-prob, on given normals."""
    ops = xtrl().ops
    cfg = R.Config(state_dim=6, num_actions=3, heads=4, depth=2, reward_range=(-2., 2.), continuous=True, squash=squash)
    sd = random_state_dict(cfg, seed=9, scale=0.4)
    B = 300
    g = torch.Generator().manual_seed(1)
    state, prev_r = torch.randn(B, 6, generator=g), torch.randn(B, generator=g)
    prev_a, normal = torch.tanh(torch.randn(B, 3, generator=g)), torch.randn(B, 3, generator=g)
    st = ops.WorldModelStepper(model_config(cfg, 4), B, DEV)
    st.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    mean, var = torch.zeros(7, device=DEV), torch.ones(7, device=DEV)
    out = st.step(cu(state), cu(prev_a), cu(prev_r), mean, var, normal=cu(normal), want_logits=True)
    ra, va, *_ = R.wm_forward(sd, cfg, state[:, None], prev_a[:, None], prev_r[:, None], 1.)
    ra = ra[:, 0]
    assert out['raw_actions'].shape == (B, 6) and rel_err(out['raw_actions'].cpu(), ra) < 3e-5
    want_a = R.continuous_sample(ra, normal, squash)
    assert rel_err(out['action'].cpu(), want_a) < 3e-5
    got_lp = out['log_prob'].cpu()
    want_lp = R.continuous_log_prob(ra, out['action'].cpu(), squash)          # at the kernel's own action
    assert got_lp.shape == (B, 3) and rel_err(got_lp, want_lp) < 1e-4


def test_wm_step_bf16_kv_cache():
    cfg = R.Config(state_dim=8, num_actions=4, heads=8, depth=4, reward_range=(-5., 5.))
    sd = random_state_dict(cfg, seed=1)
    B, T = 20, 40
    g = torch.Generator().manual_seed(2)
    states, prev_r = torch.randn(B, T, 8, generator=g), torch.randn(B, T, generator=g)
    prev_a = torch.randint(-1, 4, (B, T), generator=g)
    ra, va, _ = _rollout_kernel(sd, cfg, states, prev_a, prev_r, kv_bf16=True)
    want_ra, want_va = _rollout_oracle(sd, cfg, states, prev_a, prev_r, None)
    assert rel_err(ra, want_ra) < 1e-2 and rel_err(va, want_va) < 1e-2


# ---------------------------------------------------------------------------------------------
# (5) optimiser
# ---------------------------------------------------------------------------------------------

def test_adopt_atan2_ema_flat_vs_oracle():
    ops = xtrl().ops
    g = torch.Generator().manual_seed(8)
    shapes = [(48, 8), (48,), (192, 48), (100, 96), (1,), (4, 96)]
    sizes = [int(torch.tensor(s).prod()) for s in shapes]
    offs = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
    n = int(offs[-1])
    p0 = torch.randn(n, generator=g) * 0.3
    params = [p0[offs[i]:offs[i + 1]].clone() for i in range(len(shapes))]
    ema_o = [torch.zeros_like(p) for p in params]
    states = [dict() for _ in shapes]
    ema_sched = R.EmaState(beta=0.9, update_after_step=2, update_every=2, update_model_with_ema_every=5)
    from importlib import import_module
    sched = import_module('x-transformers-rl_b200.schedule').EmaSchedule(beta=0.9, update_after_step=2, update_every=2,
                                                                          update_model_with_ema_every=5)
    P, M, V, PI, EMA = cu(p0.clone()), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    lr = 8e-4
    for it in range(9):
        grad = torch.randn(n, generator=g) * (3. if it % 2 else 0.05)     # some steps clip, some do not
        gl = [grad[offs[i]:offs[i + 1]] for i in range(len(shapes))]
        gl, _ = R.clip_grad_norm(gl, 0.5)
        for p, gg, st in zip(params, gl, states):
            R.adopt_atan2_step(p, gg, st, lr, lr, betas=(0.9, 0.99), regen=1e-4, cautious=0.1)
        mode, decay, back = sched.next()          # host-side schedule of the product (schedule.py)
        ema_sched.update(dict(enumerate(ema_o)), dict(enumerate(params)))
        ops.adopt_atan2_step(P, cu(grad), M, V, PI, EMA, cu(offs), lr=lr, init_lr=lr, betas=(0.9, 0.99), regen_rate=1e-4,
                             cautious_factor=0.1, max_grad_norm=0.5, first_step=(it == 0), ema_mode=mode,
                             ema_decay=decay, copy_ema_to_online=back)
        assert rel_err(P.cpu(), torch.cat(params)) < 2e-5, it
        assert rel_err(EMA.cpu(), torch.cat(ema_o)) < 2e-5, it
