"""GPU parity of the learn-side fused sub-layer kernels against a plain PyTorch fp32 reference of
the same sub-layer (the x-transformers Decoder block of x_transformers_rl.py:721-734; App. A.1 -- the same torch
expression `model.py` uses, which tests/test_model_cpu.py pins to the reference's goldens).

These are floating-point tensor-core kernels with bf16 operands and fp32 accumulation: the tolerance is the
north star's bf16 bound, 1e-2 relative (to the largest magnitude of each compared tensor), written below."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import xtrl

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL_BF16 = 1e-2


def maxrel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


def ff_params(seed, D=48, H=192):
    g = torch.Generator().manual_seed(seed)
    p = dict(gamma=1. + 0.2 * torch.randn(D, generator=g), w1=torch.randn(H, D, generator=g) / math.sqrt(D),
             b1=0.3 * torch.randn(H, generator=g), w2=torch.randn(D, H, generator=g) / math.sqrt(H), b2=0.3 * torch.randn(D, generator=g))
    return {k: v.to(DEV).requires_grad_() for k, v in p.items()}


def ff_reference(x, p):
    xn = F.layer_norm(x, x.shape[-1:]) * p['gamma']
    return x + F.linear(F.gelu(F.linear(xn, p['w1'], p['b1'])), p['w2'], p['b2'])


@pytest.mark.parametrize('N', [1, 127, 128, 129, 128 * 5 + 37, 128 * 300 + 5])
def test_ff_forward_backward_vs_torch(N):
    ops = xtrl().ops
    p = ff_params(N)
    g = torch.Generator().manual_seed(N + 1)
    x = (torch.randn(N, 48, generator=g) * 1.5 + 0.3).to(DEV).requires_grad_()
    dy = torch.randn(N, 48, generator=g).to(DEV) * 1e-3
    y_ref = ff_reference(x, p)
    y_ref.backward(dy)
    img = ops.ff_pack(p['w1'].detach(), p['w2'].detach())
    y = ops.ff_fwd(x.detach(), p['gamma'].detach(), p['b1'].detach(), p['b2'].detach(), img)
    assert maxrel(y - x.detach(), (y_ref - x).detach()) < TOL_BF16
    grads = {k: torch.zeros_like(v) for k, v in p.items()}
    dx = ops.ff_bwd(x.detach(), dy, p['gamma'].detach(), p['b1'].detach(), img, grads['gamma'], grads['w1'], grads['b1'], grads['w2'], grads['b2'])
    assert maxrel(dx, x.grad) < TOL_BF16
    for k in p:
        assert maxrel(grads[k], p[k].grad) < 2 * TOL_BF16, (k, maxrel(grads[k], p[k].grad))
    # gradients accumulate (+=): a second call doubles them
    ops.ff_bwd(x.detach(), dy, p['gamma'].detach(), p['b1'].detach(), img, grads['gamma'], grads['w1'], grads['b1'], grads['w2'], grads['b2'])
    for k in p:
        assert maxrel(grads[k], 2 * p[k].grad) < 2 * TOL_BF16, k


@pytest.mark.parametrize('N', [1, 129, 40000])
def test_final_layernorm_vs_torch(N):
    ops = xtrl().ops
    g = torch.Generator().manual_seed(N)
    x = (torch.randn(N, 48, generator=g) * 2 + 0.5).to(DEV).requires_grad_()
    gamma = (1 + 0.3 * torch.randn(48, generator=g)).to(DEV).requires_grad_()
    dy = torch.randn(N, 48, generator=g).to(DEV)
    y_ref = F.layer_norm(x, (48,)) * gamma
    y_ref.backward(dy)
    assert maxrel(ops.ln_fwd(x.detach(), gamma.detach()), y_ref.detach()) < 1e-5          # fp32 kernel: the fp32 tolerance
    dgam = torch.zeros(48, device=DEV)
    dx = ops.ln_bwd(x.detach(), dy, gamma.detach(), dgam)
    assert maxrel(dx, x.grad) < 1e-4 and maxrel(dgam, gamma.grad) < 1e-4


@pytest.mark.parametrize('N,K', [(1, 4), (333, 8), (70000, 8), (70000, 4)])
def test_skinny_weight_gradient_vs_torch(N, K):
    ops = xtrl().ops
    g = torch.Generator().manual_seed(N + K)
    dy = torch.randn(N, 48, generator=g).to(DEV)
    feats = torch.randn(N, K, generator=g).to(DEV)
    dw, db = torch.zeros(48, K, device=DEV), torch.zeros(48, device=DEV)
    ops.small_dw(dy, dw, feats=feats, d_b=db)
    assert maxrel(dw, dy.double().T @ feats.double()) < 1e-4 and maxrel(db, dy.double().sum(0)) < 1e-4          # fp32 sums in another order
    ids = torch.randint(-1, K, (N,), generator=g).to(DEV)
    de = torch.zeros(K, 48, device=DEV)
    ops.small_dw(dy, de, ids=ids, out_kj=True)
    ref = torch.zeros(K + 1, 48, device=DEV, dtype=torch.float64).index_add_(0, ids + 1, dy.double())[1:]
    assert maxrel(de, ref) < 1e-4


def test_ff_dropout_mask_is_shared_by_forward_and_backward():
    """W1 = 0, b1 = c, W2 = 1: the forward output counts the kept hidden units of a token and dW2 of a one-hot dy
    lists them -- the two must agree, the keep rate must be 1-p, and the mask must follow the rng counter."""
    ops = xtrl().ops
    N, D, H, p_drop, c = 128 * 9 + 3, 48, 192, 0.25, 0.7
    x = torch.randn(N, D, device=DEV)
    gamma, b1, b2 = torch.ones(D, device=DEV), torch.full((H,), c, device=DEV), torch.zeros(D, device=DEV)
    w1, w2 = torch.zeros(H, D, device=DEV), torch.ones(D, H, device=DEV)
    img = ops.ff_pack(w1, w2)
    step = torch.tensor(5, dtype=torch.int64, device=DEV)
    unit = float((F.gelu(torch.tensor(c)) / (1 - p_drop)).bfloat16())             # one kept unit, as the bf16 operand
    y = ops.ff_fwd(x, gamma, b1, b2, img, dropout=p_drop, rng_step=step, rng_key=3)
    kept = ((y - x)[:, 0] / unit)
    assert (kept - kept.round()).abs().max() < 0.51 and abs(float(kept.mean()) / H - (1 - p_drop)) < 0.01
    y2 = ops.ff_fwd(x, gamma, b1, b2, img, dropout=p_drop, rng_step=step, rng_key=3)
    assert torch.equal(y, y2)
    assert not torch.equal(y, ops.ff_fwd(x, gamma, b1, b2, img, dropout=p_drop, rng_step=step + 1, rng_key=3))
    assert not torch.equal(y, ops.ff_fwd(x, gamma, b1, b2, img, dropout=p_drop, rng_step=step, rng_key=4))
    for t0 in (0, 127, 128, N - 1):
        dy = torch.zeros(N, D, device=DEV)
        dy[t0, 0] = 1.
        gr = [torch.zeros_like(t) for t in (gamma, w1, b1, w2, b2)]
        ops.ff_bwd(x, dy, gamma, b1, img, *gr, dropout=p_drop, rng_step=step, rng_key=3)
        mask_row = gr[3][0] / unit                                   # dW2[0][j] = h[t0][j]
        assert (mask_row - mask_row.round()).abs().max() < 0.02
        assert abs(float(mask_row.sum()) - float(kept[t0])) < 0.6
        assert float(gr[4][0]) == 1.                                 # db2 = sum of dy


# ---------------------------------------------------------------------------------------------
# whole learn step: the bf16 class with fused sub-layers against the fp32 class (torch expression, itself checked
# gradient for gradient against the oracle in test_gpu_learner.py) on the same weights and the same minibatch
# ---------------------------------------------------------------------------------------------

def _learn_once(tmp_path, tag, mixed, fused, wm, T=40, B=12, evolutionary=False, continuous=False):
    from importlib import import_module
    L = import_module('x-transformers-rl_b200.learner')
    agent_mod = import_module('x-transformers-rl_b200.agent')
    torch.manual_seed(3)
    evo = dict(evolutionary=True, latent_gene_pool=dict(dim=32, num_genes_per_island=3, num_selected=2, tournament_size=2)) if evolutionary else {}
    learner = L.Learner(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=T, batch_size=B // 3 if evolutionary else B,
                        num_episodes_per_update=B // 3 if evolutionary else B, world_model=wm, epochs=1, continuous_actions=continuous, **evo,
                        agent_kwargs=dict(save_path=str(tmp_path / f'{tag}.pt'), dropout=0.),
                        accelerate_kwargs=dict(mixed_precision='bf16' if mixed else 'no'))
    agent, dev = learner.agent, learner.device
    agent.fused_learn = fused
    agent.batch_size = B
    with torch.no_grad():
        g0 = torch.Generator(device=dev).manual_seed(11)
        agent.model.flat_params.add_(torch.randn(agent.model.flat_params.shape, generator=g0, device=dev) * 0.1)
    g = torch.Generator().manual_seed(5)
    buf = agent_mod.RolloutBuffer.allocate(T, B, 8, dev, continuous_dim=4 if continuous else 0)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g).int()
    tm_mask = torch.arange(T)[:, None] < lens[None, :]
    data = dict(states=torch.randn(T, B, 8, generator=g), actions=torch.randint(0, 4, (T, B), generator=g).int(),
                log_probs=-torch.rand(T, B, generator=g) * 1.5, rewards=torch.randn(T, B, generator=g),
                values=torch.randn(T, B, generator=g), is_boundary=torch.rand(T, B, generator=g) < 0.05)
    if continuous:                                         # tanh-squashed float actions, per-dimension log-probs
        data['actions'] = torch.tanh(torch.randn(T, B, 4, generator=g)) * 0.9
        data['log_probs'] = -torch.rand(T, B, 4, generator=g) * 1.5
    for k, v in data.items():
        getattr(buf, k).copy_(torch.where(tm_mask if v.ndim == 2 else tm_mask[..., None], v, torch.zeros_like(v)))
    buf.episode_lens.copy_(lens)
    if evolutionary:
        buf.gene_ids.copy_(torch.arange(B, dtype=torch.int32) % 3)
    agent.keep_last_grads = True
    torch.manual_seed(0)                                   # same reward-dropout flag in every variant
    agent.learn(buf, fitnesses=None, permutations=[torch.arange(B)])
    return agent.last_grads.clone(), agent.last_logs.clone(), agent.model._offsets, [n for n, _ in agent.model.named_parameters()]


WM_LANDER = dict(attn_dim_head=16, heads=4, depth=2, attn_gate_values=True, add_value_residual=True, learned_value_residual_mix=True)
WM_PLAIN8 = dict(attn_dim_head=16, heads=8, depth=2)        # the c3 block: fused attention projections + fused feed-forward
WM_PLAIN4 = dict(attn_dim_head=16, heads=4, depth=1)        # the README block


@pytest.mark.parametrize('wm', (WM_LANDER, WM_PLAIN8, WM_PLAIN4), ids=('lander', 'plain8', 'plain4'))
def test_fused_bf16_learn_step_tracks_fp32(tmp_path, wm):
    g32, s32, offs, names = _learn_once(tmp_path, 'f32', mixed=False, fused=False, wm=wm)
    gfu, sfu, _, _ = _learn_once(tmp_path, 'fused', mixed=True, fused=True, wm=wm)
    gun, sun, _, _ = _learn_once(tmp_path, 'unfused', mixed=True, fused=False, wm=wm)
    assert abs(float(sfu[0] - s32[0])) < TOL_BF16 * abs(float(s32[0]))          # total loss, 1e-2 relative (bf16 class)
    cos = lambda a, b: float(torch.dot(a, b) / (a.norm() * b.norm()))
    assert cos(gfu, g32) > 0.999 and cos(gun, g32) > 0.999
    err_fused, err_unfused = float((gfu - g32).norm() / g32.norm()), float((gun - g32).norm() / g32.norm())
    assert err_fused < 3e-2, (err_fused, err_unfused)
    assert err_fused < 1.5 * err_unfused + 5e-3, (err_fused, err_unfused)       # no worse than torch's own bf16 autocast path
    for n, o0, o1 in zip(names, offs[:-1], offs[1:]):                            # and tensor by tensor on the fused sub-layers
        fused_attn = ('.to_q.' in n or '.to_k.' in n or '.to_v.' in n or '.to_out.' in n or '.to_v_gate.' in n or
                      '.to_value_residual_mix.' in n or ('layers.0.' in n or 'layers.2.' in n) and n.endswith('gamma'))
        if '.ff.' in n or ('layers.1.' in n or 'layers.3.' in n) and n.endswith('gamma') or fused_attn:
            ref = g32[o0:o1]
            if ref.abs().max() < 1e-3 * g32.abs().max():      # e.g. the value gate at its bias-10 init: sigmoid'(10) rounds to 0 in bf16
                continue
            assert float((gfu[o0:o1] - ref).abs().max() / ref.abs().max().clamp(min=1e-12)) < 5e-2, n


# ---------------------------------------------------------------------------------------------
# attention sub-layer around the core: pre (LN, q/k/v, rotary, head layout) and post (to_out, mask, residual)
# ---------------------------------------------------------------------------------------------

def _rotate_ref(t, cos, sin):
    x, rest = t[..., :8], t[..., 8:]
    pairs = x.unflatten(-1, (-1, 2))
    half = torch.stack((-pairs[..., 1], pairs[..., 0]), dim=-1).flatten(-2)
    return torch.cat((x * cos + half * sin, rest), dim=-1)


def _tables(n):
    inv_freq = 1. / (10000 ** (torch.arange(0, 8, 2).float() / 8))
    ang = torch.arange(n, dtype=torch.float32)[:, None] * inv_freq[None]
    return ang.cos().to(DEV).contiguous(), ang.sin().to(DEV).contiguous()


@pytest.mark.parametrize('H,b,n', [(8, 3, 50), (4, 2, 77), (8, 5, 128), (8, 40, 500), (4, 1, 1)])
def test_attn_pre_post_vs_torch(H, b, n):
    ops = xtrl().ops
    g = torch.Generator().manual_seed(H * 1000 + b * 10 + n)
    I, D = 16 * H, 48
    r = lambda *s, sc=1.: (torch.randn(*s, generator=g) * sc).to(DEV).requires_grad_()
    gamma, wq, wk, wv, wo = r(D, sc=0.2), r(I, D, sc=D ** -0.5), r(I, D, sc=D ** -0.5), r(I, D, sc=D ** -0.5), r(D, I, sc=I ** -0.5)
    with torch.no_grad():
        gamma += 1.
    x = (torch.randn(b, n, D, generator=g) * 1.3 + 0.2).to(DEV).requires_grad_()
    cos, sin = _tables(n)
    cos8, sin8 = cos.repeat_interleave(2, -1), sin.repeat_interleave(2, -1)
    img = ops.attn_pack(wq.detach(), wk.detach(), wv.detach(), wo.detach())
    # ---- pre ----
    xn = F.layer_norm(x, (D,)) * gamma
    split = lambda t: t.view(b, n, H, 16).transpose(1, 2)
    q_ref, k_ref, v_ref = _rotate_ref(split(xn @ wq.T), cos8, sin8), _rotate_ref(split(xn @ wk.T), cos8, sin8), split(xn @ wv.T)
    q, k, v = ops.attn_pre_fwd(x.detach(), gamma.detach(), img, cos, sin, H)
    for got, ref in ((q, q_ref), (k, k_ref), (v, v_ref)):
        assert maxrel(got.float(), ref.detach()) < TOL_BF16
    dq, dk, dv = (torch.randn(b, H, n, 16, generator=g).to(DEV).bfloat16() for _ in range(3))
    (q_ref * dq.float()).sum().add((k_ref * dk.float()).sum()).add((v_ref * dv.float()).sum()).backward()
    gr = {k_: torch.zeros_like(t) for k_, t in dict(gamma=gamma, wq=wq, wk=wk, wv=wv).items()}
    dx = ops.attn_pre_bwd(x.detach(), dq, dk, dv, gamma.detach(), img, cos, sin, gr['gamma'], gr['wq'], gr['wk'], gr['wv'])
    assert maxrel(dx, x.grad) < TOL_BF16
    for k_, t in dict(gamma=gamma, wq=wq, wk=wk, wv=wv).items():
        assert maxrel(gr[k_], t.grad) < 2 * TOL_BF16, k_
    # the block's residual-path gradient added inside the kernel: exactly dx + d_res (same arithmetic, one more fp32 add)
    d_res = torch.randn(b, n, D, generator=g).to(DEV)
    scratch = {k_: torch.zeros_like(t) for k_, t in dict(gamma=gamma, wq=wq, wk=wk, wv=wv).items()}
    dx2 = ops.attn_pre_bwd(x.detach(), dq, dk, dv, gamma.detach(), img, cos, sin, scratch['gamma'], scratch['wq'], scratch['wk'],
                           scratch['wv'], d_res=d_res)
    assert torch.equal(dx2, dx + d_res)
    # ---- post ----
    x.grad = None
    lens = torch.randint(1, n + 1, (b,), generator=g).int().to(DEV)
    mask = torch.arange(n, device=DEV)[None, :] < lens[:, None]
    o = torch.randn(b, H, n, 16, generator=g).to(DEV).bfloat16()
    o32 = o.float().requires_grad_()
    y_ref = x + (o32.transpose(1, 2).reshape(b, n, I) @ wo.T).masked_fill(~mask[..., None], 0.)
    y = ops.attn_post_fwd(x.detach(), o, lens, img)
    assert maxrel(y - x.detach(), (y_ref - x).detach()) < TOL_BF16
    dy = torch.randn(b, n, D, generator=g).to(DEV)
    y_ref.backward(dy)
    d_wo = torch.zeros_like(wo)
    d_o = ops.attn_post_bwd(dy, o, lens, img, d_wo)
    assert maxrel(d_o.float(), o32.grad) < TOL_BF16
    assert maxrel(d_wo, wo.grad) < 2 * TOL_BF16


@pytest.mark.parametrize('H,gate,mix', [(4, True, True), (4, True, False), (4, False, True), (8, True, False), (8, False, True)])
def test_attn_pre_gate_and_mix_blocks_vs_torch(H, gate, mix):
    """the optional extra projection blocks: value-gate pre-activations and value-residual mix logits (+ their biases),
    forward and every gradient."""
    ops = xtrl().ops
    b, n, D, I = 3, 150, 48, 16 * H
    g = torch.Generator().manual_seed(H * 100 + gate * 10 + mix)
    r = lambda *s, sc=1.: (torch.randn(*s, generator=g) * sc).to(DEV).requires_grad_()
    gamma, wq, wk, wv, wo = r(D, sc=0.2), r(I, D, sc=D ** -0.5), r(I, D, sc=D ** -0.5), r(I, D, sc=D ** -0.5), r(D, I, sc=I ** -0.5)
    wg, bg, wm, bm = r(I, D, sc=D ** -0.5), r(I, sc=0.5), r(H, D, sc=D ** -0.5), r(H, sc=0.5)
    with torch.no_grad():
        gamma += 1.
    x = (torch.randn(b, n, D, generator=g) * 1.3 + 0.2).to(DEV).requires_grad_()
    cos, sin = _tables(n)
    cos8, sin8 = cos.repeat_interleave(2, -1), sin.repeat_interleave(2, -1)
    img = ops.attn_pack(wq.detach(), wk.detach(), wv.detach(), wo.detach(), wg=wg.detach() if gate else None, wm=wm.detach() if mix else None)
    xn = F.layer_norm(x, (D,)) * gamma
    split = lambda t: t.view(b, n, H, 16).transpose(1, 2)
    refs = [_rotate_ref(split(xn @ wq.T), cos8, sin8), _rotate_ref(split(xn @ wk.T), cos8, sin8), split(xn @ wv.T)]
    if gate:
        refs.append(split(xn @ wg.T + bg))
    if mix:
        refs.append(xn @ wm.T + bm)                                   # (b, n, H)
    outs = ops.attn_pre_fwd(x.detach(), gamma.detach(), img, cos, sin, H, gate_bias=bg.detach() if gate else None,
                            mix_bias=bm.detach() if mix else None)
    assert len(outs) == len(refs)
    for i, (got, ref) in enumerate(zip(outs, refs)):
        got = got[..., :H] if (mix and i == len(refs) - 1) else got
        assert maxrel(got.float(), ref.detach()) < TOL_BF16, i
    douts = [torch.randn(b, H, n, 16, generator=g).to(DEV).bfloat16() for _ in range(3 + gate)]
    dmix = None
    if mix:
        dmix = torch.zeros(b, n, 16, device=DEV, dtype=torch.bfloat16)
        dmix[..., :H] = torch.randn(b, n, H, generator=g).to(DEV).bfloat16()
    loss = sum((ref * d.float()).sum() for ref, d in zip(refs[:3 + gate], douts))
    if mix:
        loss = loss + (refs[-1] * dmix[..., :H].float()).sum()
    loss.backward()
    names = dict(gamma=gamma, wq=wq, wk=wk, wv=wv, **(dict(wg=wg, bg=bg) if gate else {}), **(dict(wm=wm, bm=bm) if mix else {}))
    gr = {k_: torch.zeros_like(t) for k_, t in names.items()}
    dx = ops.attn_pre_bwd(x.detach(), douts[0], douts[1], douts[2], gamma.detach(), img, cos, sin, gr['gamma'], gr['wq'], gr['wk'], gr['wv'],
                          dg=douts[3] if gate else None, dmix=dmix, d_wg=gr.get('wg'), d_bg=gr.get('bg'), d_wm=gr.get('wm'), d_bm=gr.get('bm'))
    assert maxrel(dx, x.grad) < TOL_BF16
    for k_, t in names.items():
        assert maxrel(gr[k_], t.grad) < 2 * TOL_BF16, (k_, maxrel(gr[k_], t.grad))
    # the output projection still reads its image from the front of the same buffer
    lens = torch.full((b,), n, dtype=torch.int32, device=DEV)
    o = torch.randn(b, H, n, 16, generator=g).to(DEV).bfloat16()
    y = ops.attn_post_fwd(x.detach(), o, lens, img)
    assert maxrel(y - x.detach(), (o.float().transpose(1, 2).reshape(b, n, I) @ wo.T).detach()) < TOL_BF16


# ---------------------------------------------------------------------------------------------
# attention core (dim_head 16, causal, n <= 512)
# ---------------------------------------------------------------------------------------------

def _core_reference(q, k, v, keep=None, p=0.):
    n = q.shape[-2]
    s = (q @ k.transpose(-1, -2)) * 0.25
    s = s.masked_fill(torch.ones(n, n, dtype=torch.bool, device=q.device).triu(1), float('-inf'))
    a = s.softmax(-1)
    if keep is not None:
        a = a * keep / (1. - p)
    return a @ v


@pytest.mark.parametrize('b,H,n', [(1, 1, 1), (2, 4, 10), (1, 2, 127), (2, 1, 128), (1, 3, 129), (2, 8, 300), (3, 8, 500), (1, 2, 512)])
def test_attn_core_vs_torch(b, H, n):
    ops = xtrl().ops
    g = torch.Generator().manual_seed(b * 100000 + H * 1000 + n)
    q, k, v, d_o = ((torch.randn(b, H, n, 16, generator=g) * sc).to(DEV).bfloat16() for sc in (1.5, 1.5, 1., 1.))
    q32, k32, v32 = (t.float().requires_grad_() for t in (q, k, v))
    o_ref = _core_reference(q32, k32, v32)
    o_ref.backward(d_o.float())
    o, lse = ops.attn_core_fwd(q, k, v)
    assert maxrel(o.float(), o_ref.detach()) < TOL_BF16
    s = (q32 @ k32.transpose(-1, -2) * 0.25).masked_fill(torch.ones(n, n, dtype=torch.bool, device=DEV).triu(1), float('-inf'))
    assert (lse - s.logsumexp(-1)).abs().max() < 2e-3
    dq, dk, dv = ops.attn_core_bwd(q, k, v, o, lse, d_o)
    # P is rounded to bf16 for the P.V product while the normaliser is fp32, so a row with a single key (n = 1: exact
    # gradient 0) carries a 2^-9 relative residue of |dO||V|: the error is measured against max(|ref|, 0.5), inputs are O(1)
    for got, ref, name in ((dq, q32.grad, 'dq'), (dk, k32.grad, 'dk'), (dv, v32.grad, 'dv')):
        err = float((got.float() - ref).abs().max() / ref.abs().max().clamp(min=0.5))
        assert err < 2 * TOL_BF16, (name, err)


def test_attn_core_dropout_mask_is_shared_and_unbiased():
    """a one-hot dO on query q0 makes dV[:, 0] the dropped attention row of q0: that recovers the mask the backward
    used; the forward and both gradients must then agree with a torch reference that applies the same mask."""
    ops = xtrl().ops
    b, H, n, p = 1, 2, 70, 0.25
    g = torch.Generator().manual_seed(9)
    q, k, v, d_o = ((torch.randn(b, H, n, 16, generator=g) * sc).to(DEV).bfloat16() for sc in (1., 1., 1., 1.))
    step = torch.tensor(3, dtype=torch.int64, device=DEV)
    kw = dict(dropout=p, rng_step=step, rng_key=2)
    o, lse = ops.attn_core_fwd(q, k, v, **kw)
    assert torch.equal(o, ops.attn_core_fwd(q, k, v, **kw)[0])
    assert not torch.equal(o, ops.attn_core_fwd(q, k, v, dropout=p, rng_step=step + 1, rng_key=2)[0])
    o_plain, lse_plain = ops.attn_core_fwd(q, k, v)
    assert (lse - lse_plain).abs().max() < 1e-5                      # the normaliser ignores the mask
    keep = torch.zeros(b, H, n, n, device=DEV)
    a_plain = (q.float() @ k.float().transpose(-1, -2) * 0.25).masked_fill(torch.ones(n, n, dtype=torch.bool, device=DEV).triu(1), float('-inf')).softmax(-1)
    for q0 in range(n):
        one = torch.zeros(b, H, n, 16, device=DEV)
        one[:, :, q0, 0] = 1.
        _, _, dv = ops.attn_core_bwd(q, k, v, o, lse, one.bfloat16(), **kw)
        keep[:, :, q0] = (dv[..., 0].float() > 0.5 * a_plain[:, :, q0] / (1 - p)).float() * (a_plain[:, :, q0] > 0)
    causal = torch.ones(n, n, device=DEV).tril()
    rate = float((keep * causal).sum() / (causal.sum() * b * H))
    assert abs(rate - (1 - p)) < 0.03, rate
    assert not torch.equal(keep[0, 0], keep[0, 1])                   # heads draw different masks
    q32, k32, v32 = (t.float().requires_grad_() for t in (q, k, v))
    o_ref = _core_reference(q32, k32, v32, keep=keep, p=p)
    o_ref.backward(d_o.float())
    assert maxrel(o.float(), o_ref.detach()) < TOL_BF16
    dq, dk, dv = ops.attn_core_bwd(q, k, v, o, lse, d_o, **kw)
    for got, ref, name in ((dq, q32.grad, 'dq'), (dk, k32.grad, 'dk'), (dv, v32.grad, 'dv')):
        assert maxrel(got.float(), ref) < 2 * TOL_BF16, (name, maxrel(got.float(), ref))


def test_fused_bf16_learn_step_evolutionary_tracks_fp32(tmp_path):
    """latent genes: the heads take a third, per-episode source (K_in 144) and hand its gradient back summed over time."""
    g32, s32, offs, names = _learn_once(tmp_path, 'e32', mixed=False, fused=False, wm=WM_LANDER, evolutionary=True)
    gfu, sfu, _, _ = _learn_once(tmp_path, 'efu', mixed=True, fused=True, wm=WM_LANDER, evolutionary=True)
    assert abs(float(sfu[0] - s32[0])) < TOL_BF16 * abs(float(s32[0]))
    assert float(torch.dot(gfu, g32) / (gfu.norm() * g32.norm())) > 0.999
    for n, o0, o1 in zip(names, offs[:-1], offs[1:]):
        if 'latent_to_embed' in n or 'critic_head' in n or 'action_head' in n:
            ref = g32[o0:o1]
            assert float((gfu[o0:o1] - ref).abs().max() / ref.abs().max().clamp(min=1e-12)) < 5e-2, n


def test_unbuilt_shapes_are_errors_in_the_bf16_class(tmp_path):
    """no multi-backend dispatch: sequences longer than the attention core's 512, or a hidden size other than 48, raise in
    the bf16 (tensor-core) class and point at the explicit fp32 parity class -- which runs them."""
    pkg = xtrl()
    with pytest.raises(pkg.XtrlError, match='sequences up to 512'):
        _learn_once(tmp_path, 'long', mixed=True, fused=True, wm=WM_PLAIN8, T=520, B=2)
    g_ref, s_ref, _, _ = _learn_once(tmp_path, 'long32', mixed=False, fused=False, wm=WM_PLAIN8, T=520, B=2)
    assert torch.isfinite(g_ref).all()
    from importlib import import_module
    model_mod = import_module('x-transformers-rl_b200.model')
    ops = pkg.ops
    cfg = ops.ModelConfig(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=16, dim=64, heads=4, dim_head=16, depth=1)
    m = model_mod.WorldModelActorCritic(cfg).to(DEV)
    with pytest.raises(pkg.XtrlError, match='hidden_dim=64'):
        m.fused_state()


def test_fused_bf16_learn_step_continuous_tracks_fp32(tmp_path):
    """continuous actions in the bf16 class: Gaussian actor head (2A outputs) and the nn.Linear action embedding feed the
    fused head kernels."""
    g32, s32, offs, names = _learn_once(tmp_path, 'c32', mixed=False, fused=False, wm=WM_PLAIN4, continuous=True)
    gfu, sfu, _, _ = _learn_once(tmp_path, 'cfu', mixed=True, fused=True, wm=WM_PLAIN4, continuous=True)
    assert abs(float(sfu[0] - s32[0])) < TOL_BF16 * max(1., abs(float(s32[0])))
    assert float(torch.dot(gfu, g32) / (gfu.norm() * g32.norm())) > 0.995


# ---------------------------------------------------------------------------------------------
# head MLPs (critic / actor / world-model + done), zero-padded dense weights
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize('nsrc,hid,outp,cols,n_silu', [(2, 96, 112, 100, 96), (2, 96, 16, 4, 96), (2, 64, 32, 19, 48), (3, 96, 112, 100, 96), (3, 96, 16, 8, 96)])
def test_head_mlp_forward_backward_vs_torch(nsrc, hid, outp, cols, n_silu):
    ops = xtrl().ops
    b, n = 5, 77                                                    # 385 tokens: three full tiles and a ragged one
    N, kin = b * n, 48 * nsrc
    g = torch.Generator().manual_seed(nsrc * 1000 + hid + cols)
    r = lambda *s, sc=1.: (torch.randn(*s, generator=g) * sc).to(DEV)
    w1, b1, w2, b2 = r(hid, kin, sc=kin ** -0.5), r(hid, sc=0.3), r(outp, hid, sc=hid ** -0.5), r(outp, sc=0.3)
    w1[n_silu + 1:] = 0; b1[n_silu + 1:] = 0; w2[cols:] = 0; b2[cols:] = 0; w2[:, n_silu + 1:] = 0        # padding rows / columns are zero
    params = [t.clone().requires_grad_() for t in (w1, b1, w2, b2)]
    srcs = [r(N, 48).requires_grad_(), r(N, 48).requires_grad_()] + ([r(b, 48).requires_grad_()] if nsrc == 3 else [])
    full = list(srcs[:2]) + ([srcs[2][:, None].expand(b, n, 48).reshape(N, 48)] if nsrc == 3 else [])
    pre = torch.cat(full, -1) @ params[0].T + params[1]
    h = torch.cat((F.silu(pre[:, :n_silu]), pre[:, n_silu:]), -1)
    y_ref = (h @ params[2].T + params[3])[:, :cols]
    gy = (torch.randn(N, cols, generator=g) * 1e-2).to(DEV).bfloat16()
    y_ref.backward(gy.float())
    img = ops.mlp_pack(w1, w2)
    det = [s.detach() for s in srcs]
    y = ops.mlp_fwd(det, n, b1, b2, img, cols, n_silu, per_episode2=nsrc == 3)
    assert y.shape == (N, cols) and maxrel(y.float(), y_ref.detach()) < TOL_BF16
    gr = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
    dsrc = ops.mlp_bwd(det, n, gy, b1, img, *gr, n_silu, per_episode2=nsrc == 3)
    for i in range(2):
        assert maxrel(dsrc[i], srcs[i].grad) < 2 * TOL_BF16, i
    if nsrc == 3:
        assert maxrel(dsrc[2].view(b, n, 48).sum(1), srcs[2].grad) < 2 * TOL_BF16
    names = ('w1', 'b1', 'w2', 'b2')
    for name, got, p in zip(names, gr, params):
        ref = p.grad
        if name == 'w1':
            got, ref = got[:n_silu + 1], ref[:n_silu + 1]             # rows of zero-padded hidden units carry no meaning
        if name == 'b1':
            got, ref = got[:n_silu + 1], ref[:n_silu + 1]
        if name == 'w2':
            got, ref = got[:cols, :n_silu + 1], ref[:cols, :n_silu + 1]
        if name == 'b2':
            got, ref = got[:cols], ref[:cols]
        assert maxrel(got, ref) < 2 * TOL_BF16, (name, maxrel(got, ref))


def test_attn_core_exact_max_path_for_large_norms():
    """when scale * max|q| * max|k| exceeds the bound threshold the forward takes its exact-row-max path: same result."""
    ops = xtrl().ops
    b, H, n = 1, 2, 300
    g = torch.Generator().manual_seed(77)
    q, k = ((torch.randn(b, H, n, 16, generator=g) * 5.).to(DEV).bfloat16() for _ in range(2))          # |q||k| / 4 ~ 100 > 40
    v, d_o = ((torch.randn(b, H, n, 16, generator=g)).to(DEV).bfloat16() for _ in range(2))
    assert float(0.25 * q.float().norm(dim=-1).max() * k.float().norm(dim=-1).max()) > 40.
    q32, k32, v32 = (t.float().requires_grad_() for t in (q, k, v))
    o_ref = _core_reference(q32, k32, v32)
    o_ref.backward(d_o.float())
    o, lse = ops.attn_core_fwd(q, k, v)
    assert maxrel(o.float(), o_ref.detach()) < TOL_BF16
    dq, dk, dv = ops.attn_core_bwd(q, k, v, o, lse, d_o)
    for got, ref, name in ((dq, q32.grad, 'dq'), (dk, k32.grad, 'dk'), (dv, v32.grad, 'dv')):
        err = float((got.float() - ref).abs().max() / ref.abs().max().clamp(min=0.5))
        assert err < 2 * TOL_BF16, (name, err)


# ---------------------------------------------------------------------------------------------
# minibatch prologue (gather + shifts + RSNorm + input embeddings) and its backward, csrc/learn_prep.cu
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize('continuous', (False, True))
@pytest.mark.parametrize('S,A,E,b,T', ((8, 4, 9, 5, 37), (5, 2, 4, 4, 10), (16, 16, 3, 2, 130)))
def test_learn_prologue_forward_backward_vs_torch(continuous, S, A, E, b, T):
    """xtrl_learn_prep_fwd / _bwd against the torch expression of Agent.learn's minibatch prep (:903-926) and the input side of
    WorldModelActorCritic.forward (:492-503, :516-529): exact fp32 arithmetic, so 1e-5."""
    ops = xtrl().ops
    g = torch.Generator().manual_seed(S * 100 + A + T)
    r = lambda *s, sc=1.: (torch.randn(*s, generator=g) * sc).to(DEV)
    lens = torch.randint(1, T + 1, (E,), generator=g).int().to(DEV)
    src = dict(states=r(E, T, S), actions=(torch.tanh(r(E, T, A)) if continuous else torch.randint(0, A, (E, T), generator=g).int().to(DEV)),
               rewards=r(E, T), old_log_probs=-r(E, T, A).abs() if continuous else -r(E, T).abs(), returns=r(E, T), old_values=r(E, T),
               dones=(torch.rand(E, T, generator=g) < 0.1).to(DEV), lens=lens, gene_ids=torch.arange(E, dtype=torch.int32, device=DEV) % 3)
    idx = torch.randperm(E, generator=g)[:b].to(DEV)
    mean, var = r(S + 1, sc=0.3), (torch.rand(S + 1, generator=g) + 0.5).to(DEV)
    var[1] = 0.                                                       # the clamp (variance 0 after the first RSNorm update)
    w_in, w_se, b_se, re = r(48, S, sc=0.3).requires_grad_(), r(48, S, sc=0.3).requires_grad_(), r(48, sc=0.1).requires_grad_(), r(48, sc=0.1).requires_grad_()
    w_act = (r(48, A, sc=0.3) if continuous else r(A, 48, sc=0.3)).requires_grad_()
    b_act = r(48, sc=0.1).requires_grad_() if continuous else None
    flag = torch.tensor(1., device=DEV)
    out = ops.learn_prep_fwd(src, idx, mean, var, w_in.detach(), w_se.detach(), b_se.detach(), w_act.detach(),
                             b_act.detach() if continuous else None, re.detach(), flag, continuous=continuous)
    # torch reference
    st, ac, rw = src['states'][idx], src['actions'][idx], src['rewards'][idx]
    prev_r = F.pad(rw, (1, -1), value=0.)
    packed = (torch.cat((st, prev_r[..., None]), -1) - mean) / var.sqrt().clamp(min=1e-5)
    if continuous:
        prev_a = F.pad(ac, (0, 0, 1, -1), value=0.)
        ae_prev, ae_next = F.linear(prev_a, w_act, b_act), F.linear(ac, w_act, b_act)
    else:
        prev_a = F.pad(ac, (1, -1), value=-1)
        emb = lambda ids: F.embedding(ids.clamp(min=0).long(), w_act) * (ids >= 0)[..., None]
        ae_prev, ae_next = emb(prev_a), emb(ac)
    x0 = F.linear(packed[..., :S], w_in) + ae_prev + packed[..., S:] * re * flag
    se = F.linear(packed[..., :S], w_se, b_se)
    assert maxrel(out['packed'], packed) < 1e-5 and maxrel(out['x0'], x0.detach()) < 1e-5
    assert maxrel(out['se'], se.detach()) < 1e-5 and maxrel(out['ae_next'], ae_next.detach()) < 1e-5
    for k in ('returns', 'old_values', 'old_log_probs'):
        assert torch.equal(out[k], src[k][idx]), k
    assert torch.equal(out['actions'], ac) and torch.equal(out['dones'].bool(), src['dones'][idx])
    assert torch.equal(out['lens'], lens[idx]) and torch.equal(out['gene_ids'], src['gene_ids'][idx])
    mask = torch.arange(T, device=DEV)[None] < lens[idx][:, None]
    want_sum = (packed * mask[..., None]).sum((0, 1))
    assert maxrel(out['rs_sum'][:S + 1], want_sum) < 1e-4 and float(out['rs_sum'][S + 1]) == float(mask.sum())
    # backward
    dx0, dse, dae = r(b, T, 48, sc=1e-2), r(b, T, 48, sc=1e-2), r(b, T, 48, sc=1e-2)
    ((x0 * dx0).sum() + (se * dse).sum() + (ae_next * dae).sum()).backward()
    grads = dict(w_in=torch.zeros_like(w_in), w_se=torch.zeros_like(w_se), b_se=torch.zeros_like(b_se), w_act=torch.zeros_like(w_act),
                 b_act=torch.zeros(48, device=DEV) if continuous else None, re=torch.zeros_like(re))
    ops.learn_prep_bwd(dx0, dse, dae, out['packed'], out['actions'], flag, grads['w_in'], grads['w_se'], grads['b_se'], grads['w_act'],
                       grads['b_act'], grads['re'], continuous=continuous)
    for name, p in (('w_in', w_in), ('w_se', w_se), ('b_se', b_se), ('w_act', w_act), ('re', re)) + ((('b_act', b_act),) if continuous else ()):
        assert maxrel(grads[name], p.grad) < 1e-4, (name, maxrel(grads[name], p.grad))


# ---------------------------------------------------------------------------------------------
# the launch-count helpers of the captured learn graph
# ---------------------------------------------------------------------------------------------

def test_seg2d_copy_and_accumulate_vs_torch():
    """xtrl_seg2d_copy: rectangular blocks between parameter tensors and padded dense buffers (head refresh / scatter)"""
    ops = xtrl().ops
    g = torch.Generator().manual_seed(3)
    r = lambda *s: torch.randn(*s, generator=g).to(DEV)
    pad_w, pad_b = torch.zeros(64, 96, device=DEV), torch.zeros(64, device=DEV)
    w0, b0, w1, b1 = r(48, 96), r(48), r(1, 96), r(1)
    pad_w2, w2 = torch.zeros(32, 64, device=DEV), r(19, 48)
    pairs = [(w0, pad_w[0:48]), (b0, pad_b[0:48]), (w1, pad_w[48:49]), (b1, pad_b[48:49]), (w2, pad_w2[0:19, :48])]
    ops.SegTable(pairs).run()
    assert torch.equal(pad_w[:48], w0) and torch.equal(pad_w[48], w1[0]) and torch.equal(pad_b[:48], b0) and pad_b[48] == b1[0]
    assert torch.equal(pad_w2[:19, :48], w2) and float(pad_w2[:, 48:].abs().sum()) == 0 and float(pad_w[49:].abs().sum()) == 0
    # nine blocks (two launches of <= 8) accumulating the other way, as the gradient scatter does
    grads = [torch.ones_like(t) for t, _ in pairs] + [torch.ones(5, device=DEV) for _ in range(4)]
    srcs = [d for _, d in pairs] + [r(5) for _ in range(4)]
    want = [gd + s_ for gd, s_ in zip(grads, srcs)]
    ops.SegTable(list(zip(srcs, grads))).run(accumulate=True)
    for got, ref in zip(grads, want):
        assert torch.equal(got, ref)


def test_attention_block_node_equals_the_three_nodes():
    """_FusedAttnBlock (one autograd node, residual gradient added inside xtrl_attn_pre_bwd) against _FusedAttnPre ->
    _FusedAttnCore -> _FusedAttnPost with autograd's own add: same forward bits, same dx up to the order of one fp32 add,
    parameter gradients equal up to the atomicAdd order of the weight-gradient flush."""
    from importlib import import_module
    M, ops = import_module('x-transformers-rl_b200.model'), xtrl().ops
    torch.manual_seed(0)
    cfg = ops.ModelConfig(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=200, depth=1, heads=8, dim_head=16)
    model = M.WorldModelActorCritic(cfg, dropout=0.25).to(DEV)
    dec = model.transformer.attn_layers
    norm, attn = dec.layers[0]
    fs = model.fused_state()
    ops.attn_pack(attn.to_q.weight.data, attn.to_k.weight.data, attn.to_v.weight.data, attn.to_out.weight.data, out=fs['attn_img'][0])
    b, n = 3, 200
    cos4, sin4 = model._rotary_tables(n)
    assert model._rotary_tables(n)[0] is cos4                                  # cached per sequence length
    lens = torch.tensor([200, 37, 150], dtype=torch.int32, device=DEV)
    x0 = torch.randn(b, n, 48, device=DEV)
    dy = torch.randn(b, n, 48, device=DEV)
    outs = {}
    for mode in ('block', 'nodes'):
        model.flat_grads.zero_()
        x = x0.clone().requires_grad_()
        fs['rng_step'].fill_(7)                                                # the same dropout mask in both runs
        if mode == 'block':
            y = M._FusedAttnBlock.apply(x, norm, attn, fs['attn_img'][0], cos4, sin4, 8, lens, fs['rng_step'], 1000, 0.25)
        else:
            q, k, v = M._FusedAttnPre.apply(x, norm, attn, fs['attn_img'][0], cos4, sin4, 8, False, False)[:3]
            o = M._FusedAttnCore.apply(q, k, v, fs['rng_step'], 1000, 0.25)
            y = M._FusedAttnPost.apply(x, o, attn, fs['attn_img'][0], lens)
        y.backward(dy)
        outs[mode] = (y.detach().clone(), x.grad.clone(), model.flat_grads.clone())
    (yb, dxb, gb), (yn, dxn, gn) = outs['block'], outs['nodes']
    assert torch.equal(yb, yn)
    assert maxrel(dxb, dxn) < 1e-6
    assert maxrel(gb, gn) < 1e-5


def test_three_heads_node_equals_separate_heads():
    """_FusedHeads3 (one autograd node; the three xtrl_mlp_bwd launches add into shared d_embed / d_se buffers with
    frac_gradient's factor, xtrl_mlp_io.d_src_scale / d_src_accumulate) against three _FusedHead nodes + _FracGrad + autograd's
    adds: same outputs, the same input and parameter gradients up to fp32 summation order."""
    from importlib import import_module
    M, ops = import_module('x-transformers-rl_b200.model'), xtrl().ops
    torch.manual_seed(1)
    cfg = ops.ModelConfig(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=150, depth=1, heads=8, dim_head=16)
    model = M.WorldModelActorCritic(cfg, dropout=0.).to(DEV)
    hd = model.fused_state()['heads']
    for h in hd.values():
        h.refresh()
    b, n = 3, 150                                                              # 450 tokens: three full tiles and a partial one
    src = [torch.randn(b, n, 48, device=DEV) for _ in range(3)]
    gs = None
    outs = {}
    for mode in ('one', 'three'):
        model.flat_grads.zero_()
        embed, se, ae = (t.clone().requires_grad_() for t in src)
        if mode == 'one':
            ra, vl, sp18, dl = M._FusedHeads3.apply(hd, n, model.frac, embed, se, ae)
            sp = torch.cat([sp18, dl[..., None]], dim=-1)            # the node hands the loss its two dense inputs
        else:
            sp = M._FusedHead.apply(hd['pred'], n, embed, ae, None)
            emb_ac = M._FracGrad.apply(embed, model.frac)
            ra, vl = M._FusedHead.apply(hd['actor'], n, emb_ac, se, None), M._FusedHead.apply(hd['critic'], n, emb_ac, se, None)
        if gs is None:
            gs = [torch.randn_like(t.float()).to(t.dtype) for t in (ra, vl, sp)]
        torch.autograd.backward([ra, vl, sp], gs)
        outs[mode] = ([t.detach().clone() for t in (ra, vl, sp)], [t.grad.clone() for t in (embed, se, ae)], model.flat_grads.clone())
    (ya, ga, pa), (yb, gb, pb) = outs['one'], outs['three']
    for u, v in zip(ya, yb):
        assert torch.equal(u, v)
    for u, v, name in zip(ga, gb, ('d_embed', 'd_se', 'd_ae')):
        assert maxrel(u, v) < 1e-5, (name, maxrel(u, v))
    assert maxrel(pa, pb) < 1e-5
