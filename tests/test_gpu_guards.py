"""GPU: memory safety without compute-sanitizer (closed on this pool, profiles/r02_sanitizer_memcheck.log).  Every output
tensor the ops allocate is carved out of a larger buffer with sentinel-filled guard zones on both sides; after each kernel the
guards must be untouched.  Sizes are ragged on purpose (tails of tiles, rows, chunks)."""
import contextlib
import importlib

import pytest
import torch

from helpers import xtrl

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GUARD = 4096          # bytes on each side


class _Guarded:
    def __init__(self):
        self.records = []

    def empty(self, *shape, dtype=torch.float32, device=None, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        n = 1
        for s_ in shape:
            n *= int(s_)
        item = self._real_empty((), dtype=dtype).element_size()
        g = GUARD // item
        raw = self._real_empty(n + 2 * g, dtype=dtype, device=device)
        sent = raw.view(torch.uint8)
        sent[:GUARD] = 0xA5
        sent[GUARD + n * item:] = 0xA5
        self.records.append((raw, n, item))
        return raw[g:g + n].view(*shape)

    def check(self):
        for raw, n, item in self.records:
            sent = raw.view(torch.uint8)
            assert bool((sent[:GUARD] == 0xA5).all()), 'write before an output buffer'
            assert bool((sent[GUARD + n * item:] == 0xA5).all()), 'write past an output buffer'
        return len(self.records)


@contextlib.contextmanager
def guarded_outputs():
    g = _Guarded()
    g._real_empty = torch.empty
    real_like = torch.empty_like
    torch.empty = g.empty
    torch.empty_like = lambda t, **kw: g.empty(*t.shape, dtype=kw.get('dtype', t.dtype), device=t.device)
    try:
        yield g
    finally:
        torch.empty, torch.empty_like = g._real_empty, real_like
    torch.cuda.synchronize()
    assert g.check() > 0


def test_outputs_of_the_learn_side_kernels_stay_inside_their_buffers():
    ops = xtrl().ops
    g = torch.Generator().manual_seed(0)
    r = lambda *s, sc=1.: (torch.randn(*s, generator=g) * sc).to(DEV)
    b, n, H, D = 3, 131, 8, 48                      # 393 tokens: three full 128-token tiles and a tail of 9; n crosses one 128-key tile
    N, I = b * n, 16 * H
    x, dy = r(b, n, D), r(b, n, D, sc=1e-2)
    gamma, w1, b1, w2, b2 = 1 + r(D, sc=.1), r(4 * D, D, sc=.1), r(4 * D, sc=.1), r(D, 4 * D, sc=.1), r(D, sc=.1)
    wq, wk, wv, wo = r(I, D, sc=.1), r(I, D, sc=.1), r(I, D, sc=.1), r(D, I, sc=.1)
    step = torch.zeros((), dtype=torch.int64, device=DEV)
    inv_freq = 1. / (10000 ** (torch.arange(0, 8, 2, device=DEV).float() / 8))
    ang = torch.arange(n, device=DEV, dtype=torch.float32)[:, None] * inv_freq[None]
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    lens = torch.tensor([n, n - 7, 5], dtype=torch.int32, device=DEV)
    z = torch.zeros_like
    with guarded_outputs():
        ff_img, at_img = ops.ff_pack(w1, w2), ops.attn_pack(wq, wk, wv, wo)
        y = ops.ff_fwd(x, gamma, b1, b2, ff_img, dropout=0.25, rng_step=step, rng_key=1)
        ops.ff_bwd(x, dy, gamma, b1, ff_img, z(gamma), z(w1), z(b1), z(w2), z(b2), dropout=0.25, rng_step=step, rng_key=1)
        q, k, v = ops.attn_pre_fwd(x, gamma, at_img, cos, sin, H)
        o, lse = ops.attn_core_fwd(q, k, v, dropout=0.25, rng_step=step, rng_key=2)
        dq, dk, dv = ops.attn_core_bwd(q, k, v, o, lse, (o * 0.01).contiguous(), dropout=0.25, rng_step=step, rng_key=2)
        ops.attn_pre_bwd(x, dq, dk, dv, gamma, at_img, cos, sin, z(gamma), z(wq), z(wk), z(wv))
        y2 = ops.attn_post_fwd(x, o, lens, at_img)
        ops.attn_post_bwd(dy, o, lens, at_img, z(wo))
        ops.ln_bwd(x.view(N, D), dy.view(N, D), gamma, z(gamma)); ops.ln_fwd(x.view(N, D), gamma)
        w1p, b1p, w2p, b2p = r(96, 96, sc=.1), r(96, sc=.1), r(112, 96, sc=.1), r(112, sc=.1)
        w2p[100:] = 0; b2p[100:] = 0
        img = ops.mlp_pack(w1p, w2p)
        srcs = [r(N, 48), r(N, 48)]
        yy = ops.mlp_fwd(srcs, n, b1p, b2p, img, 100, 96)
        ops.mlp_bwd(srcs, n, (yy * 0.01).contiguous(), b1p, img, z(w1p), z(b1p), z(w2p), z(b2p), 96)
    assert torch.isfinite(y).all() and torch.isfinite(y2).all() and torch.isfinite(dq.float()).all()


def test_outputs_of_loss_gae_prologue_evolve_optimiser_stay_inside_their_buffers():
    ops = xtrl().ops
    g = torch.Generator().manual_seed(1)
    r = lambda *s, sc=1.: (torch.randn(*s, generator=g) * sc).to(DEV)
    E, T, S, A, K = 7, 133, 8, 4, 100                       # 931 rows: seven 128-row loss tiles and a tail of 35
    lens = torch.randint(1, T + 1, (E,), generator=g).int().to(DEV)
    src = dict(states=r(E, T, S), actions=torch.randint(0, A, (E, T), generator=g).int().to(DEV), rewards=r(E, T), old_log_probs=-r(E, T).abs(),
               returns=r(E, T), old_values=r(E, T), dones=(torch.rand(E, T, generator=g) < 0.1).to(DEV), lens=lens,
               gene_ids=torch.zeros(E, dtype=torch.int32, device=DEV))
    idx = torch.randperm(E, generator=g)[:5].to(DEV)
    flag = torch.ones((), device=DEV)
    cfg = ops.LossConfig(num_actions=A, num_bins=K, state_dim=S, reward_range=(-5., 5.))
    with guarded_outputs():
        out = ops.learn_prep_fwd(src, idx, r(S + 1, sc=.1), r(S + 1).abs() + 0.5, r(48, S, sc=.1), r(48, S, sc=.1), r(48, sc=.1), r(A, 48, sc=.1),
                                 None, r(48, sc=.1), flag)
        b = idx.numel()
        for dt in (torch.float32, torch.bfloat16):
            sc, gr = ops.ppo_loss_fwd_bwd(cfg, r(b, T, A).to(dt), r(b, T, K).to(dt), r(b, T, 2 * (S + 1)).to(dt), r(b, T).to(dt), out['actions'],
                                          out['old_log_probs'], out['returns'], out['old_values'], out['packed'], out['dones'], out['lens'])
            assert torch.isfinite(sc).all()
        ret = ops.gae(src['rewards'], src['old_values'], src['dones'])
        ret_tm = ops.gae(src['rewards'].T.contiguous(), src['old_values'].T.contiguous(), src['dones'].T.contiguous(), time_major=True)
        genes = torch.nn.functional.normalize(r(3 * 8, 32), dim=-1)
        sel = ops.evolve_(genes, r(24), r(3, 4, 4), r(3, 4, 32), r(3, 7, 32), num_islands=3, num_genes_per_island=8, num_selected=4, tournament_size=2)
        hv = ops.hl_decode(r(37, K), (-5., 5.))
    assert torch.allclose(ret, ret_tm.T, atol=1e-5) and sel.shape == (3, 4) and torch.isfinite(hv).all()
    n = 1003                                                # not a multiple of 4: the optimiser's 16-byte units have a tail
    offs = torch.tensor([0, 10, 333, 700, n], device=DEV)
    P, G_, M, V, PI, EM = (r(n, sc=.1) for _ in range(6))
    before = [t.clone() for t in (P, M, V, EM)]
    ops.adopt_atan2_step(P, G_, M, V.abs_(), PI, EM, offs, lr=8e-4, init_lr=8e-4, betas=(0.9, 0.99), regen_rate=1e-4, cautious_factor=0.1,
                         max_grad_norm=0.5, first_step=False, ema_mode=2, ema_decay=0.9)
    assert torch.isfinite(P).all() and not torch.equal(P, before[0])
