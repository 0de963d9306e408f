"""torch-tensor front of the C ABI (include/xtrl.h).  Every function launches hand-written sm_100a
kernels from libxtrl.so on torch's current CUDA stream; tensors only provide device memory.

Each function mirrors one operator seam of the reference:
    gae                    <- calc_gae                     x_transformers_rl.py:616-640
    hl_decode              <- HLGaussLoss(logits)          :843, :427
    ppo_loss_fwd_bwd       <- compute_*_loss + assembly    :398-477, :939-983
    evolve                 <- LatentGenePool.evolve_       evolution.py:76-184
    adopt_atan2_step       <- clip_grad_norm_/AdoptAtan2/EMA  :987-992, :753
    WorldModelStepper.step <- state_to_pred_action_and_value + sample  :1250-1289
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, name='tensor'):
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise L.XtrlError(f'{name} must be a CUDA tensor (there is no CPU path)')
    if dtype is not None and t.dtype != dtype:
        raise L.XtrlError(f'{name} must be {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise L.XtrlError(f'{name} must be contiguous')
    return C.c_void_p(t.data_ptr())


# --------------------------------------------------------------------------------------------
# (2) GAE
# --------------------------------------------------------------------------------------------

def gae(rewards, values, is_boundary, gamma=0.99, lam=0.95, time_major=False, out=None):
    """returns = gae + values.  rewards/values fp32, is_boundary bool/uint8; (E,T) episode-major
    (the reference's layout) or, with time_major=True, (T,B) as the rollout buffer stores them."""
    lib = L.load()
    if rewards.shape != values.shape or rewards.shape != is_boundary.shape or rewards.ndim != 2:
        raise L.XtrlError(f'gae: shapes {tuple(rewards.shape)}, {tuple(values.shape)}, {tuple(is_boundary.shape)}')
    bd = is_boundary.view(torch.uint8) if is_boundary.dtype == torch.bool else is_boundary
    out = torch.empty_like(rewards) if out is None else out
    if time_major:
        T, E = rewards.shape
        se, st = 1, E
    else:
        E, T = rewards.shape
        se, st = T, 1
    L.check(lib.xtrl_gae(_ptr(rewards, torch.float32, 'rewards'), _ptr(values, torch.float32, 'values'),
                         _ptr(bd, torch.uint8, 'is_boundary'), _ptr(out, torch.float32, 'out'),
                         E, T, se, st, gamma, lam, _stream()), 'xtrl_gae')
    return out


def hl_decode(logits, reward_range):
    lib = L.load()
    K = logits.shape[-1]
    out = torch.empty(logits.shape[:-1], device=logits.device, dtype=torch.float32)
    L.check(lib.xtrl_hl_decode(_ptr(logits, torch.float32, 'logits'), _ptr(out), out.numel(), K,
                               float(reward_range[0]), float(reward_range[1]), _stream()), 'xtrl_hl_decode')
    return out


# --------------------------------------------------------------------------------------------
# (3) fused PPO loss forward + backward w.r.t. the head outputs
# --------------------------------------------------------------------------------------------

@dataclass
class LossConfig:
    num_actions: int
    num_bins: int
    state_dim: int
    reward_range: tuple
    eps_clip: float = 0.2
    value_clip: float = 0.4
    entropy_weight: float = 0.01
    actor_weight: float = 1.
    critic_weight: float = 1.
    autoreg_weight: float = 1.
    normalize_advantages: bool = True
    continuous: bool = False
    squash: bool = True

    def c_struct(self, bf16):
        return L.LossCfg(self.num_actions, self.num_bins, self.state_dim, float(self.reward_range[0]),
                         float(self.reward_range[1]), self.eps_clip, self.value_clip, self.entropy_weight,
                         self.actor_weight, self.critic_weight, self.autoreg_weight,
                         int(self.normalize_advantages), int(bf16), int(self.continuous), int(self.squash))


SCALAR_NAMES = ('total', 'actor_loss', 'critic_loss', 'autoreg_loss', 'pred_done_loss', 'n_masked', 'adv_mean', 'adv_var')


def ppo_loss_fwd_bwd(cfg: LossConfig, raw_actions, value_logits, state_pred_raw, done_logit, actions,
                     old_log_probs, returns, old_values, target, dones, episode_lens, per_step=False):
    """All rows of a (b, T) minibatch.  Head outputs (b,T,A), (b,T,K), (b,T,2(S+1)) raw (mean, logvar
    interleaved), (b,T) pre-sigmoid; fp32 or bf16.  Returns (scalars[8], grads dict[, per_step (b,T,2)])."""
    lib = L.load()
    b, T = actions.shape[:2]
    io_dtype = raw_actions.dtype
    if io_dtype not in (torch.float32, torch.bfloat16):
        raise L.XtrlError(f'ppo_loss: head outputs must be fp32 or bf16, got {io_dtype}')
    for n, t_ in (('value_logits', value_logits), ('state_pred_raw', state_pred_raw), ('done_logit', done_logit)):
        if t_ is not None and t_.dtype != io_dtype:
            raise L.XtrlError(f'ppo_loss: {n} dtype {t_.dtype} != {io_dtype}')
    dev = raw_actions.device
    bd = None if dones is None else (dones.view(torch.uint8) if dones.dtype == torch.bool else dones)
    g_ra, g_vl = torch.empty_like(raw_actions), torch.empty_like(value_logits)
    g_sp = None if state_pred_raw is None else torch.empty_like(state_pred_raw)
    g_dl = None if done_logit is None else torch.empty_like(done_logit)
    ps = torch.empty(b, T, 2, device=dev, dtype=torch.float32) if per_step else None
    scalars = torch.empty(8, device=dev, dtype=torch.float32)
    ws = torch.empty(lib.xtrl_ppo_loss_workspace_bytes() // 8, device=dev, dtype=torch.float64)
    io = L.LossIO(_ptr(raw_actions, None, 'raw_actions'), _ptr(value_logits, None, 'value_logits'),
                  _ptr(state_pred_raw, None, 'state_pred_raw'), _ptr(done_logit, None, 'done_logit'),
                  _ptr(None if cfg.continuous else actions, torch.int32, 'actions'), _ptr(old_log_probs, torch.float32, 'old_log_probs'),
                  _ptr(returns, torch.float32, 'returns'), _ptr(old_values, torch.float32, 'old_values'),
                  _ptr(target, torch.float32, 'target'), _ptr(bd, torch.uint8, 'dones'),
                  _ptr(episode_lens, torch.int32, 'episode_lens'),
                  _ptr(g_ra), _ptr(g_vl), _ptr(g_sp), _ptr(g_dl), _ptr(ps), _ptr(scalars), _ptr(ws),
                  _ptr(actions if cfg.continuous else None, torch.float32, 'actions (continuous)'))
    cs = cfg.c_struct(io_dtype == torch.bfloat16)
    L.check(lib.xtrl_ppo_loss_fwd_bwd(C.byref(cs), C.byref(io), b, T, _stream()), 'xtrl_ppo_loss_fwd_bwd')
    grads = dict(raw_actions=g_ra, value_logits=g_vl, state_pred_raw=g_sp, done_logit=g_dl)
    return (scalars, grads, ps) if per_step else (scalars, grads)


# --------------------------------------------------------------------------------------------
# (4) gene pool
# --------------------------------------------------------------------------------------------

def evolve_(genes, fitness, noise_tourn, noise_cross, noise_mut, *, num_islands, num_genes_per_island,
            num_selected, tournament_size, num_elites=1, mutation_std_dev=0.1, migrate=False,
            num_migrate=0, temperature=1.5):
    """in-place evolution step; returns selected gene ids (I, s) int64 like the reference."""
    lib = L.load()
    G = genes.shape[-1]
    sel = torch.empty(num_islands, num_selected, device=genes.device, dtype=torch.int64)
    pd = L.PoolDesc(G, num_islands, num_genes_per_island, num_selected, tournament_size, num_elites,
                    int(migrate), int(num_migrate), mutation_std_dev, temperature)
    L.check(lib.xtrl_evolve(_ptr(genes, torch.float32, 'genes'), _ptr(fitness, torch.float32, 'fitness'),
                            _ptr(noise_tourn, torch.float32, 'noise_tourn'), _ptr(noise_cross, torch.float32, 'noise_cross'),
                            _ptr(noise_mut, torch.float32, 'noise_mut'), _ptr(sel), _ptr(None), C.byref(pd),
                            _stream()), 'xtrl_evolve')
    return sel


# --------------------------------------------------------------------------------------------
# (5) optimiser over flat buffers
# --------------------------------------------------------------------------------------------

def adopt_atan2_step(params, grads, m, v, param_init, ema, seg_offsets, *, lr, init_lr, betas, regen_rate,
                     cautious_factor, max_grad_norm, first_step, ema_mode=0, ema_decay=0., copy_ema_to_online=False,
                     a=1.27, b=1., workspace=None, grad_norm_out=None, grad_scale=1., peers=None):
    """one cooperative launch: global-norm clip of grads * grad_scale, AdoptAtan2, EMA (csrc/optim.cu).  grad_scale = 1 / world
    after a SUM all-reduce of the flat gradient buffer -- or, with peers (a distributed.PeerGroup whose `local` holds grads),
    the sum over the ranks' buffers is the first phase of the same kernel (peers.reduced receives it)."""
    lib = L.load()
    nt = seg_offsets.numel() - 1
    n = params.numel()
    if workspace is None:
        workspace = torch.empty(lib.xtrl_opt_workspace_bytes(nt) // 8, device=params.device, dtype=torch.float64)
    cfg = L.OptCfg(lr, init_lr, betas[0], betas[1], regen_rate, cautious_factor, a, b, max_grad_norm,
                   int(first_step), float(ema_mode), ema_decay, int(copy_ema_to_online), float(grad_scale))
    if peers is not None:
        if grads.data_ptr() != peers.local.data_ptr():
            raise L.XtrlError('adopt_atan2_step: with peers the gradients must be the head of peers.local')
        pc = peers.ctx()
        L.check(lib.xtrl_adopt_atan2_step_peers(_ptr(params, torch.float32, 'params'), _ptr(grads, torch.float32, 'grads'),
                                                _ptr(m, torch.float32), _ptr(v, torch.float32), _ptr(param_init, torch.float32),
                                                _ptr(ema, torch.float32), _ptr(seg_offsets, torch.int64, 'seg_offsets'), nt, n,
                                                C.byref(cfg), _ptr(workspace), _ptr(grad_norm_out), C.byref(pc), _stream()),
                'xtrl_adopt_atan2_step_peers')
        return
    L.check(lib.xtrl_adopt_atan2_step(_ptr(params, torch.float32, 'params'), _ptr(grads, torch.float32, 'grads'),
                                      _ptr(m, torch.float32), _ptr(v, torch.float32), _ptr(param_init, torch.float32),
                                      _ptr(ema, torch.float32), _ptr(seg_offsets, torch.int64, 'seg_offsets'), nt, n,
                                      C.byref(cfg), _ptr(workspace), _ptr(grad_norm_out), _stream()),
            'xtrl_adopt_atan2_step')


# --------------------------------------------------------------------------------------------
# (6) learn-side fused sub-layers (tcgen05)
# --------------------------------------------------------------------------------------------

def ff_supported(dim, hidden) -> bool:
    return L.load().xtrl_ff_img_bytes(dim, hidden) > 0


def ff_pack(w1, w2, out=None):
    """bf16 operand images of the feed-forward weights (w1 (hidden, dim), w2 (dim, hidden)); once per weight update."""
    lib = L.load()
    hidden, dim = w1.shape
    nbytes = lib.xtrl_ff_img_bytes(dim, hidden)
    if nbytes < 0:
        raise L.XtrlError(f'ff_pack: dim={dim}, hidden={hidden} is not built')
    out = torch.empty(nbytes, dtype=torch.uint8, device=w1.device) if out is None else out
    L.check(lib.xtrl_ff_pack(_ptr(w1, torch.float32, 'w1'), _ptr(w2, torch.float32, 'w2'), _ptr(out, torch.uint8, 'img'),
                             dim, hidden, _stream()), 'xtrl_ff_pack')
    return out


def ff_fwd(x, ln_gamma, b1, b2, img, *, dropout=0., rng_step=None, rng_key=0, out=None):
    """y = x + FF(LayerNorm(x) * gamma), x (..., dim) fp32."""
    lib = L.load()
    dim, hidden = x.shape[-1], b1.numel()
    y = torch.empty_like(x) if out is None else out
    io = L.FfIO(x=_ptr(x, torch.float32, 'x'), ln_gamma=_ptr(ln_gamma, torch.float32, 'ln_gamma'), b1=_ptr(b1, torch.float32, 'b1'),
                b2=_ptr(b2, torch.float32, 'b2'), w_img=_ptr(img, torch.uint8, 'img'), y=_ptr(y, torch.float32, 'y'),
                rng_step=_ptr(rng_step, torch.int64, 'rng_step'), rng_key=rng_key, dropout=dropout)
    if dropout > 0. and rng_step is None:
        raise L.XtrlError('ff_fwd: dropout needs the rng_step counter tensor')
    L.check(lib.xtrl_ff_fwd(C.byref(io), x.numel() // dim, dim, hidden, _stream()), 'xtrl_ff_fwd')
    return y


def ff_bwd(x, dy, ln_gamma, b1, img, d_ln_gamma, d_w1, d_b1, d_w2, d_b2, *, dropout=0., rng_step=None, rng_key=0):
    """dx of the feed-forward sub-layer; parameter gradients are accumulated (+=) into the d_* tensors."""
    lib = L.load()
    dim, hidden = x.shape[-1], b1.numel()
    dx = torch.empty_like(x)
    f32 = torch.float32
    io = L.FfIO(x=_ptr(x, f32, 'x'), dy=_ptr(dy, f32, 'dy'), ln_gamma=_ptr(ln_gamma, f32, 'ln_gamma'), b1=_ptr(b1, f32, 'b1'),
                w_img=_ptr(img, torch.uint8, 'img'), dx=_ptr(dx, f32, 'dx'), d_ln_gamma=_ptr(d_ln_gamma, f32, 'd_ln_gamma'),
                d_w1=_ptr(d_w1, f32, 'd_w1'), d_b1=_ptr(d_b1, f32, 'd_b1'), d_w2=_ptr(d_w2, f32, 'd_w2'), d_b2=_ptr(d_b2, f32, 'd_b2'),
                rng_step=_ptr(rng_step, torch.int64, 'rng_step'), rng_key=rng_key, dropout=dropout)
    L.check(lib.xtrl_ff_bwd(C.byref(io), x.numel() // dim, dim, hidden, _stream()), 'xtrl_ff_bwd')
    return dx


def attn_supported(dim, heads, dim_head, gate=False, mix=False) -> bool:
    return L.load().xtrl_attn_img_bytes(dim, heads, dim_head, int(gate), int(mix)) > 0


def attn_pack(wq, wk, wv, wo, wg=None, wm=None, out=None):
    """bf16 operand images of to_q/to_k/to_v (H*16, dim), to_out (dim, H*16) and, optionally, to_v_gate (H*16, dim) and
    to_value_residual_mix (H, dim); once per weight update."""
    lib = L.load()
    inner, dim = wq.shape
    heads = inner // 16
    nbytes = lib.xtrl_attn_img_bytes(dim, heads, 16, int(wg is not None), int(wm is not None))
    if nbytes < 0:
        raise L.XtrlError(f'attn_pack: dim={dim}, heads={heads}, gate={wg is not None}, mix={wm is not None} is not built')
    out = torch.empty(nbytes, dtype=torch.uint8, device=wq.device) if out is None else out
    f32 = torch.float32
    L.check(lib.xtrl_attn_pack(_ptr(wq, f32, 'wq'), _ptr(wk, f32, 'wk'), _ptr(wv, f32, 'wv'), _ptr(wo, f32, 'wo'),
                               _ptr(wg, f32, 'wg'), _ptr(wm, f32, 'wm'), _ptr(out, torch.uint8, 'img'), dim, heads, 16, _stream()),
            'xtrl_attn_pack')
    return out


def _ptr_array(tensors, dtype, what):
    return (C.c_void_p * len(tensors))(*[_ptr(t, dtype, what) for t in tensors])


def ff_pack_layers(w1s, w2s, imgs):
    """ff_pack for several layers (<= 8 per launch) at once"""
    lib = L.load()
    hidden, dim = w1s[0].shape
    for i in range(0, len(w1s), 8):
        sl = slice(i, i + 8)
        a1, a2, ai = _ptr_array(w1s[sl], torch.float32, 'w1'), _ptr_array(w2s[sl], torch.float32, 'w2'), _ptr_array(imgs[sl], torch.uint8, 'img')
        L.check(lib.xtrl_ff_pack_layers(C.cast(a1, C.c_void_p), C.cast(a2, C.c_void_p), C.cast(ai, C.c_void_p), len(w1s[sl]), dim, hidden,
                                        _stream()), 'xtrl_ff_pack_layers')


def attn_pack_layers(wqs, wks, wvs, wos, wgs, wms, imgs):
    """attn_pack for several layers of ONE variant (wgs / wms: lists or None) at once (<= 8 per launch)"""
    lib = L.load()
    inner, dim = wqs[0].shape
    f32 = torch.float32
    for i in range(0, len(wqs), 8):
        sl = slice(i, i + 8)
        arrs = [_ptr_array(t[sl], f32, nm) for t, nm in ((wqs, 'wq'), (wks, 'wk'), (wvs, 'wv'), (wos, 'wo'))]
        ag = _ptr_array(wgs[sl], f32, 'wg') if wgs is not None else None
        am = _ptr_array(wms[sl], f32, 'wm') if wms is not None else None
        ai = _ptr_array(imgs[sl], torch.uint8, 'img')
        cast = lambda a: C.cast(a, C.c_void_p) if a is not None else None
        L.check(lib.xtrl_attn_pack_layers(*[cast(a) for a in arrs], cast(ag), cast(am), cast(ai), len(wqs[sl]), dim, inner // 16, 16,
                                          _stream()), 'xtrl_attn_pack_layers')


def attn_pre_fwd(x, ln_gamma, img, cos_tab, sin_tab, heads, gate_bias=None, mix_bias=None):
    """x (b, n, dim) fp32 -> q, k, v bf16 (b, H, n, 16) [, gate pre-activation (b, H, n, 16)] [, mix logits (b, n, 16)]:
    LayerNorm, the projections (biases of the optional blocks added), partial rotary on q and k."""
    lib = L.load()
    b, n, _ = x.shape
    bf, f32 = torch.bfloat16, torch.float32
    nb = 4 if gate_bias is not None else 3
    qkv = torch.empty(nb, b, heads, n, 16, dtype=bf, device=x.device)
    mix = torch.empty(b, n, 16, dtype=bf, device=x.device) if mix_bias is not None else None
    io = L.AttnIO(x=_ptr(x, f32, 'x'), ln_gamma=_ptr(ln_gamma, f32, 'ln_gamma'), cos_tab=_ptr(cos_tab, f32, 'cos_tab'),
                  sin_tab=_ptr(sin_tab, f32, 'sin_tab'), w_img=_ptr(img, torch.uint8, 'img'),
                  q=_ptr(qkv[0]), k=_ptr(qkv[1]), v=_ptr(qkv[2]), gate=_ptr(qkv[3]) if nb == 4 else None, mix=_ptr(mix),
                  gate_bias=_ptr(gate_bias, f32, 'gate_bias'), mix_bias=_ptr(mix_bias, f32, 'mix_bias'))
    L.check(lib.xtrl_attn_pre_fwd(C.byref(io), b, n, heads, _stream()), 'xtrl_attn_pre_fwd')
    return tuple(qkv[i] for i in range(nb)) + ((mix,) if mix is not None else ())


def attn_pre_bwd(x, dq, dk, dv, ln_gamma, img, cos_tab, sin_tab, d_ln_gamma, d_wq, d_wk, d_wv, dg=None, dmix=None,
                 d_wg=None, d_bg=None, d_wm=None, d_bm=None, d_res=None):
    """dx of the LayerNorm -> projection path (+ d_res, the gradient arriving over the block's residual path, when given);
    parameter gradients accumulated (+=)."""
    lib = L.load()
    b, heads, n, _ = dq.shape
    dx = torch.empty_like(x)
    f32, bf = torch.float32, torch.bfloat16
    io = L.AttnIO(x=_ptr(x, f32, 'x'), ln_gamma=_ptr(ln_gamma, f32, 'ln_gamma'), cos_tab=_ptr(cos_tab, f32, 'cos_tab'),
                  sin_tab=_ptr(sin_tab, f32, 'sin_tab'), w_img=_ptr(img, torch.uint8, 'img'),
                  dq=_ptr(dq, bf, 'dq'), dk=_ptr(dk, bf, 'dk'), dv=_ptr(dv, bf, 'dv'), dx=_ptr(dx, f32, 'dx'),
                  d_ln_gamma=_ptr(d_ln_gamma, f32), d_wq=_ptr(d_wq, f32), d_wk=_ptr(d_wk, f32), d_wv=_ptr(d_wv, f32),
                  d_gate=_ptr(dg, bf, 'dg'), d_mix=_ptr(dmix, bf, 'dmix'), d_wg=_ptr(d_wg, f32), d_bg=_ptr(d_bg, f32),
                  d_wm=_ptr(d_wm, f32), d_bm=_ptr(d_bm, f32), dy=_ptr(d_res, f32, 'd_res'))
    L.check(lib.xtrl_attn_pre_bwd(C.byref(io), b, n, heads, _stream()), 'xtrl_attn_pre_bwd')
    return dx


def attn_post_fwd(x, o, lens, img):
    """y = x + mask * (o . Wo^T); x (b, n, dim) fp32, o (b, H, n, 16) bf16, lens (b,) int32."""
    lib = L.load()
    b, heads, n, _ = o.shape
    y = torch.empty_like(x)
    io = L.AttnIO(x=_ptr(x, torch.float32, 'x'), o=_ptr(o, torch.bfloat16, 'o'), lens=_ptr(lens, torch.int32, 'lens'),
                  w_img=_ptr(img, torch.uint8, 'img'), y=_ptr(y, torch.float32, 'y'))
    L.check(lib.xtrl_attn_post_fwd(C.byref(io), b, n, heads, _stream()), 'xtrl_attn_post_fwd')
    return y


def attn_post_bwd(dy, o, lens, img, d_wo):
    """d_o (b, H, n, 16) bf16 of the output projection; d_wo accumulated (+=).  (dx of the residual path is dy.)"""
    lib = L.load()
    b, heads, n, _ = o.shape
    d_o = torch.empty_like(o)
    io = L.AttnIO(dy=_ptr(dy, torch.float32, 'dy'), o=_ptr(o, torch.bfloat16, 'o'), lens=_ptr(lens, torch.int32, 'lens'),
                  w_img=_ptr(img, torch.uint8, 'img'), d_o=_ptr(d_o, torch.bfloat16, 'd_o'), d_wo=_ptr(d_wo, torch.float32, 'd_wo'))
    L.check(lib.xtrl_attn_post_bwd(C.byref(io), b, n, heads, _stream()), 'xtrl_attn_post_bwd')
    return d_o


def mlp_supported(in_dim, hidden, out) -> bool:
    return L.load().xtrl_mlp_img_bytes(in_dim, hidden, out) > 0


class SegTable:
    """a fixed list of rectangular blocks (src tensor view -> dst tensor view, both 1-D or 2-D fp32 with unit inner stride) for
    xtrl_seg2d_copy; built once, replayed every step (the views' addresses are captured here)."""

    def __init__(self, pairs):
        segs = []
        for src, dst in pairs:
            assert src.shape == dst.shape and src.dtype == dst.dtype == torch.float32 and src.dim() in (1, 2)
            if src.dim() == 1:
                src, dst = src[None], dst[None]
            assert src.stride(1) == 1 and dst.stride(1) == 1
            rows, cols = src.shape
            segs.append(L.Seg2d(src.data_ptr(), dst.data_ptr(), rows, cols, src.stride(0) if rows > 1 else cols,
                                dst.stride(0) if rows > 1 else cols))
        self.keep = pairs
        self.chunks = [(L.Seg2d * len(segs[i:i + 8]))(*segs[i:i + 8]) for i in range(0, len(segs), 8)]

    def run(self, accumulate=False):
        lib = L.load()
        for ch in self.chunks:
            L.check(lib.xtrl_seg2d_copy(C.cast(ch, C.c_void_p), len(ch), int(accumulate), _stream()), 'xtrl_seg2d_copy')


def mlp_pack(w1p, w2p, out=None):
    """operand images of the zero-padded dense head weights w1p (hidden, in_dim), w2p (out, hidden)."""
    lib = L.load()
    hidden, in_dim = w1p.shape
    outp = w2p.shape[0]
    nbytes = lib.xtrl_mlp_img_bytes(in_dim, hidden, outp)
    if nbytes < 0:
        raise L.XtrlError(f'mlp_pack: ({in_dim}, {hidden}, {outp}) is not built')
    out = torch.empty(nbytes, dtype=torch.uint8, device=w1p.device) if out is None else out
    L.check(lib.xtrl_mlp_pack(_ptr(w1p, torch.float32, 'w1p'), _ptr(w2p, torch.float32, 'w2p'), _ptr(out, torch.uint8, 'img'), in_dim, hidden,
                              outp, _stream()), 'xtrl_mlp_pack')
    return out


def mlp_fwd(srcs, n, b1p, b2p, img, out_cols, n_silu, per_episode2=False):
    """y (N, out_cols) bf16 = W2 act(W1 [srcs] + b1) + b2; srcs: 2 or 3 fp32 tensors of 48-wide rows (N tokens, n per episode)."""
    lib = L.load()
    f32 = torch.float32
    N = srcs[0].numel() // 48
    y = torch.empty(N, out_cols, dtype=torch.bfloat16, device=srcs[0].device)
    io = L.MlpIO(src0=_ptr(srcs[0], f32, 'src0'), src1=_ptr(srcs[1], f32, 'src1'), src2=_ptr(srcs[2], f32, 'src2') if len(srcs) > 2 else None,
                 src2_per_episode=int(per_episode2), b1=_ptr(b1p, f32, 'b1'), b2=_ptr(b2p, f32, 'b2'), w_img=_ptr(img, torch.uint8, 'img'),
                 y=_ptr(y), n_silu=n_silu, out_cols=out_cols)
    L.check(lib.xtrl_mlp_fwd(C.byref(io), N, n, 48 * len(srcs), b1p.numel(), b2p.numel(), _stream()), 'xtrl_mlp_fwd')
    return y


def mlp_bwd(srcs, n, gy, b1p, img, d_w1p, d_b1p, d_w2p, d_b2p, n_silu, per_episode2=False, out=None, scales=None, accumulate=0):
    """gradients w.r.t. every source ((N, 48) fp32 each); padded parameter gradients accumulated (+=).  out: the tensors that
    receive them (default: new ones); source s gets (bit s of `accumulate` ? out[s] : 0) + scales[s] * dX_s."""
    lib = L.load()
    f32 = torch.float32
    N = srcs[0].numel() // 48
    dsrc = [torch.empty(N, 48, dtype=f32, device=gy.device) for _ in srcs] if out is None else list(out)
    if accumulate and out is None:
        raise L.XtrlError('mlp_bwd: accumulate needs the tensors to add to (out=)')
    sc = list(scales) + [1.] * (3 - len(scales)) if scales is not None else [1., 1., 1.]
    io = L.MlpIO(src0=_ptr(srcs[0], f32, 'src0'), src1=_ptr(srcs[1], f32, 'src1'), src2=_ptr(srcs[2], f32, 'src2') if len(srcs) > 2 else None,
                 src2_per_episode=int(per_episode2), b1=_ptr(b1p, f32, 'b1'), w_img=_ptr(img, torch.uint8, 'img'), gy=_ptr(gy, torch.bfloat16, 'gy'),
                 d_src0=_ptr(dsrc[0]), d_src1=_ptr(dsrc[1]), d_src2=_ptr(dsrc[2]) if len(srcs) > 2 else None,
                 d_w1=_ptr(d_w1p, f32), d_b1=_ptr(d_b1p, f32), d_w2=_ptr(d_w2p, f32), d_b2=_ptr(d_b2p, f32), n_silu=n_silu, out_cols=gy.shape[-1],
                 d_src_scale=(C.c_float * 3)(*[float(v) for v in sc]), d_src_accumulate=int(accumulate))
    L.check(lib.xtrl_mlp_bwd(C.byref(io), N, n, 48 * len(srcs), b1p.numel(), d_b2p.numel(), _stream()), 'xtrl_mlp_bwd')
    return dsrc


def small_dw(dy, d_w, *, feats=None, ids=None, d_b=None, out_kj=False):
    """d_w (+=) sum_t f[t] (x) dy[t] for a skinny input layer (K <= 16 features, or the one-hot of ids) into 48 outputs."""
    K = d_w.shape[0] if out_kj else d_w.shape[1]
    L.check(L.load().xtrl_small_dw(_ptr(feats, torch.float32, 'feats'), _ptr(ids, torch.int64, 'ids'), _ptr(dy, torch.float32, 'dy'),
                                   _ptr(d_w, torch.float32, 'd_w'), _ptr(d_b, torch.float32, 'd_b'), dy.numel() // 48, K, 48, int(out_kj),
                                   _stream()), 'xtrl_small_dw')
    return d_w


def ln_fwd(x, gamma):
    """F.layer_norm(x) * gamma over the last dim (48), fp32."""
    y = torch.empty_like(x)
    L.check(L.load().xtrl_ln_fwd(_ptr(x, torch.float32, 'x'), _ptr(gamma, torch.float32, 'gamma'), _ptr(y), x.numel() // x.shape[-1],
                                 x.shape[-1], _stream()), 'xtrl_ln_fwd')
    return y


def ln_bwd(x, dy, gamma, d_gamma):
    dx = torch.empty_like(x)
    L.check(L.load().xtrl_ln_bwd(_ptr(x, torch.float32, 'x'), _ptr(dy, torch.float32, 'dy'), _ptr(gamma, torch.float32, 'gamma'), _ptr(dx),
                                 _ptr(d_gamma, torch.float32, 'd_gamma'), x.numel() // x.shape[-1], x.shape[-1], _stream()), 'xtrl_ln_bwd')
    return dx


def learn_prep_fwd(src, idx, rs_mean, rs_var, w_in, w_se, b_se, w_act, b_act, reward_embed, reward_flag, *, continuous=False):
    """minibatch prologue (csrc/learn_prep.cu): gather the episodes `idx` (b,) int64 out of the episode-major trajectory tensors
    `src` (dict: states (E,T,S), actions (E,T) i32 | (E,T,A) f32, rewards, old_log_probs, returns, old_values (E,T) f32, dones (E,T)
    bool/u8, lens (E,) i32[, gene_ids (E,) i32]), normalise (state || prev reward), and form the model's input rows.
    Returns dict(x0, se, ae_next (b,T,48) f32, packed (b,T,S+1), actions, old_log_probs, returns, old_values, dones, lens[, gene_ids],
    rs_sum (S+2,): masked column sums of packed and the masked row count)."""
    lib = L.load()
    f32, dev = torch.float32, idx.device
    E, T, S = src['states'].shape
    b = idx.numel()
    A = w_act.shape[1] if continuous else w_act.shape[0]
    z = lambda *shape, dt=f32: torch.empty(*shape, dtype=dt, device=dev)
    out = dict(x0=z(b, T, 48), se=z(b, T, 48), ae_next=z(b, T, 48), packed=z(b, T, S + 1),
               actions=z(b, T, A) if continuous else z(b, T, dt=torch.int32), old_log_probs=z(b, T, A) if continuous else z(b, T),
               returns=z(b, T), old_values=z(b, T), dones=z(b, T, dt=torch.uint8), lens=z(b, dt=torch.int32),
               rs_sum=torch.zeros(S + 2, dtype=f32, device=dev))
    has_genes = src.get('gene_ids') is not None
    if has_genes:
        out['gene_ids'] = z(b, dt=torch.int32)
    dn = src['dones'].view(torch.uint8) if src['dones'].dtype == torch.bool else src['dones']
    io = L.PrepIO(states=_ptr(src['states'], f32, 'states'),
                  actions=_ptr(src['actions'], f32 if continuous else torch.int32, 'actions'), rewards=_ptr(src['rewards'], f32, 'rewards'),
                  old_log_probs=_ptr(src['old_log_probs'], f32, 'old_log_probs'), returns=_ptr(src['returns'], f32, 'returns'),
                  old_values=_ptr(src['old_values'], f32, 'old_values'), dones=_ptr(dn, torch.uint8, 'dones'), lens=_ptr(src['lens'], torch.int32, 'lens'),
                  gene_ids=_ptr(src['gene_ids'], torch.int32, 'gene_ids') if has_genes else None, idx=_ptr(idx, torch.int64, 'idx'),
                  rs_mean=_ptr(rs_mean, f32, 'rs_mean'), rs_var=_ptr(rs_var, f32, 'rs_var'), w_in=_ptr(w_in, f32, 'w_in'), w_se=_ptr(w_se, f32, 'w_se'),
                  b_se=_ptr(b_se, f32, 'b_se'), w_act=_ptr(w_act, f32, 'w_act'), b_act=_ptr(b_act, f32, 'b_act') if continuous else None,
                  reward_embed=_ptr(reward_embed, f32, 'reward_embed'), reward_flag=_ptr(reward_flag, f32, 'reward_flag'),
                  x0=_ptr(out['x0']), se=_ptr(out['se']), ae_next=_ptr(out['ae_next']), packed=_ptr(out['packed']), mb_actions=_ptr(out['actions']),
                  mb_old_log_probs=_ptr(out['old_log_probs']), mb_returns=_ptr(out['returns']), mb_old_values=_ptr(out['old_values']),
                  mb_dones=_ptr(out['dones']), mb_lens=_ptr(out['lens']), mb_gene_ids=_ptr(out['gene_ids']) if has_genes else None,
                  rs_sum=_ptr(out['rs_sum']))
    L.check(lib.xtrl_learn_prep_fwd(C.byref(io), b, T, S, A, int(continuous), _stream()), 'xtrl_learn_prep_fwd')
    return out


def learn_prep_bwd(dx0, d_se, d_ae_next, packed, mb_actions, reward_flag, d_w_in, d_w_se, d_b_se, d_w_act, d_b_act, d_reward_embed, *,
                   continuous=False):
    """every skinny parameter gradient of the prologue in one pass; accumulated (+=) into the d_* tensors."""
    lib = L.load()
    f32 = torch.float32
    b, T, _ = dx0.shape
    S = packed.shape[-1] - 1
    A = d_w_act.shape[1] if continuous else d_w_act.shape[0]
    io = L.PrepBwdIO(dx0=_ptr(dx0, f32, 'dx0'), d_se=_ptr(d_se, f32, 'd_se'), d_ae_next=_ptr(d_ae_next, f32, 'd_ae_next'), packed=_ptr(packed, f32, 'packed'),
                     mb_actions=_ptr(mb_actions, f32 if continuous else torch.int32, 'actions'), reward_flag=_ptr(reward_flag, f32, 'reward_flag'), T=T,
                     d_w_in=_ptr(d_w_in, f32), d_w_se=_ptr(d_w_se, f32), d_b_se=_ptr(d_b_se, f32), d_w_act=_ptr(d_w_act, f32),
                     d_b_act=_ptr(d_b_act, f32) if continuous else None, d_reward_embed=_ptr(d_reward_embed, f32))
    L.check(lib.xtrl_learn_prep_bwd(C.byref(io), b, T, S, A, int(continuous), _stream()), 'xtrl_learn_prep_bwd')


def attn_core_supported(n, dim_head) -> bool:
    return bool(L.load().xtrl_attn_core_supported(n, dim_head))


def attn_core_fwd(q, k, v, *, dropout=0., rng_step=None, rng_key=0, scale=None):
    """causal attention core, q/k/v (b, H, n, 16) bf16 -> o (b, H, n, 16) bf16, lse (b, H, n) fp32."""
    lib = L.load()
    b, heads, n, dh = q.shape
    bf = torch.bfloat16
    o = torch.empty_like(q)
    lse = torch.empty(b, heads, n, dtype=torch.float32, device=q.device)
    if dropout > 0. and rng_step is None:
        raise L.XtrlError('attn_core_fwd: dropout needs the rng_step counter tensor')
    io = L.AttnCoreIO(q=_ptr(q, bf, 'q'), k=_ptr(k, bf, 'k'), v=_ptr(v, bf, 'v'), o_out=_ptr(o), lse=_ptr(lse),
                      rng_step=_ptr(rng_step, torch.int64, 'rng_step'), rng_key=rng_key, dropout=dropout,
                      scale=dh ** -0.5 if scale is None else scale)
    L.check(lib.xtrl_attn_core_fwd(C.byref(io), b * heads, n, dh, _stream()), 'xtrl_attn_core_fwd')
    return o, lse


def attn_core_bwd(q, k, v, o, lse, d_o, *, dropout=0., rng_step=None, rng_key=0, scale=None):
    lib = L.load()
    b, heads, n, dh = q.shape
    bf = torch.bfloat16
    dqkv = torch.empty(3, *q.shape, dtype=bf, device=q.device)
    io = L.AttnCoreIO(q=_ptr(q, bf, 'q'), k=_ptr(k, bf, 'k'), v=_ptr(v, bf, 'v'), o=_ptr(o, bf, 'o'), d_o=_ptr(d_o, bf, 'd_o'),
                      dq=_ptr(dqkv[0]), dk=_ptr(dqkv[1]), dv=_ptr(dqkv[2]), lse=_ptr(lse, torch.float32, 'lse'),
                      rng_step=_ptr(rng_step, torch.int64, 'rng_step'), rng_key=rng_key, dropout=dropout,
                      scale=dh ** -0.5 if scale is None else scale)
    L.check(lib.xtrl_attn_core_bwd(C.byref(io), b * heads, n, dh, _stream()), 'xtrl_attn_core_bwd')
    return dqkv[0], dqkv[1], dqkv[2]


# --------------------------------------------------------------------------------------------
# (1) rollout step
# --------------------------------------------------------------------------------------------

@dataclass
class ModelConfig:
    """shape of the WorldModelActorCritic the kernels are asked to run (x_transformers_rl.py:281-369,
    :709-735).  world_model keys accepted: attn_dim_head, heads, depth, attn_gate_values,
    add_value_residual, learned_value_residual_mix -- anything else is an error."""
    state_dim: int
    num_actions: int
    reward_range: tuple
    max_timesteps: int
    dim: int = 48
    heads: int = 4
    dim_head: int = 16
    depth: int = 1
    num_bins: int = 100
    dim_gene: int = 0
    gate_values: bool = False
    value_residual: bool = False
    learned_mix: bool = False
    rotary_true_pos: bool = False
    kv_bf16: bool = False
    dense_bf16: bool = False      # tcgen05 bf16 GEMMs for QKV/out/FFN/heads (1e-2 tolerance class)
    continuous: bool = False      # continuous_actions: Gaussian (mean, log-variance) actor head, float action vectors
    squash: bool = True           # squash_continuous

    WORLD_MODEL_KEYS = ('attn_dim_head', 'heads', 'depth', 'attn_gate_values', 'add_value_residual',
                        'learned_value_residual_mix')

    @classmethod
    def from_learner_kwargs(cls, state_dim, num_actions, reward_range, max_timesteps, world_model, hidden_dim=48,
                            num_bins=100, dim_gene=0, **extra):
        unknown = set(world_model) - set(cls.WORLD_MODEL_KEYS)
        if unknown:
            raise L.XtrlError(f'world_model keys {sorted(unknown)} are not supported by the B200 kernels '
                              f'(supported: {cls.WORLD_MODEL_KEYS})')
        vres = bool(world_model.get('add_value_residual', False))
        return cls(state_dim=state_dim, num_actions=num_actions, reward_range=tuple(reward_range),
                   max_timesteps=max_timesteps, dim=hidden_dim, heads=world_model.get('heads', 8),
                   dim_head=world_model.get('attn_dim_head', 64), depth=world_model['depth'], num_bins=num_bins,
                   dim_gene=dim_gene, gate_values=bool(world_model.get('attn_gate_values', False)),
                   value_residual=vres, learned_mix=vres and bool(world_model.get('learned_value_residual_mix', False)),
                   **extra)

    @property
    def inner(self):
        return self.heads * self.dim_head

    def c_struct(self):
        return L.ModelDesc(self.state_dim, self.num_actions, self.dim, self.heads, self.dim_head, self.depth,
                           self.num_bins, self.dim_gene, int(self.gate_values), int(self.value_residual),
                           int(self.learned_mix), self.max_timesteps, int(self.rotary_true_pos), int(self.kv_bf16),
                           float(self.reward_range[0]), float(self.reward_range[1]), int(self.dense_bf16),
                           int(self.continuous), int(self.squash))


def weight_layout(cfg: ModelConfig) -> L.WeightLayout:
    lib = L.load()
    out = L.WeightLayout()
    d = cfg.c_struct()
    L.check(lib.xtrl_wm_weight_layout(C.byref(d), C.byref(out)), 'xtrl_wm_weight_layout')
    return out


def pack_weights(sd: dict, cfg: ModelConfig, out=None, device=None):
    """reference-style state_dict -> the packed fp32 blob xtrl_wm_step reads (K-major weights)."""
    lay = weight_layout(cfg)
    device = device if device is not None else next(iter(sd.values())).device
    blob = torch.zeros(lay.total, device=device, dtype=torch.float32) if out is None else out.zero_()
    D, I, H, S, A, K, G = cfg.dim, cfg.inner, cfg.heads, cfg.state_dim, cfg.num_actions, cfg.num_bins, cfg.dim_gene

    def put(off, t, rows=None, ld=None):
        t = t.detach().to(device=device, dtype=torch.float32)
        if t.ndim == 1:
            blob[off:off + t.numel()] = t
        else:  # (rows, cols) written with leading dimension ld
            r, c = t.shape
            ld = c if ld is None else ld
            blob[off:off + r * ld].view(r, ld)[:, :c] = t

    tf = 'transformer.attn_layers'
    in_proj = torch.cat((sd['transformer.project_in.weight'].T, sd['to_state_embed.weight'].T), dim=1)   # (S, 2D)
    put(lay.in_proj, in_proj)
    put(lay.state_embed_bias, sd['to_state_embed.bias'])
    if cfg.continuous:
        put(lay.action_embed, sd['action_embeds.weight'].T)           # nn.Linear(A, D) -> [A][D]
        put(lay.action_embed_bias, sd['action_embeds.bias'])
    else:
        put(lay.action_embed, sd['action_embeds.embed.weight'])
    put(lay.reward_embed, sd['reward_embed'])
    if G > 0:
        put(lay.latent_w, sd['latent_to_embed.weight'].T)
        put(lay.latent_b, sd['latent_to_embed.bias'])
    put(lay.final_gamma, sd[f'{tf}.final_norm.gamma'])
    put(lay.head1_w, torch.cat((sd['action_head.0.weight'].T, sd['critic_head.0.weight'].T), dim=1))
    put(lay.head1_b, torch.cat((sd['action_head.0.bias'], sd['critic_head.0.bias'])))
    a_pad, k_pad = ((2 * A if cfg.continuous else A) + 3) & ~3, (K + 3) & ~3
    put(lay.actor2_w, sd['action_head.2.weight'].T, ld=a_pad)
    put(lay.actor2_b, sd['action_head.2.bias'])
    put(lay.critic2_w, sd['critic_head.2.weight'].T, ld=k_pad)
    put(lay.critic2_b, sd['critic_head.2.bias'])
    for l in range(cfg.depth):
        pa, pf = f'{tf}.layers.{2 * l}', f'{tf}.layers.{2 * l + 1}'
        ll = lay.layer[l]
        put(ll.ln1_gamma, sd[f'{pa}.0.gamma'])
        cols = [sd[f'{pa}.1.to_q.weight'].T, sd[f'{pa}.1.to_k.weight'].T, sd[f'{pa}.1.to_v.weight'].T]
        bias = [torch.zeros(3 * I, device=device)]
        if cfg.gate_values:
            cols.append(sd[f'{pa}.1.to_v_gate.weight'].T)
            bias.append(sd[f'{pa}.1.to_v_gate.bias'].to(device))
        if cfg.learned_mix:
            hp = (H + 3) & ~3
            w, bb = torch.zeros(D, hp, device=device), torch.zeros(hp, device=device)
            if l > 0:
                w[:, :H] = sd[f'{pa}.1.to_value_residual_mix.0.weight'].T
                bb[:H] = sd[f'{pa}.1.to_value_residual_mix.0.bias']
            cols.append(w)
            bias.append(bb)
        qkv = torch.cat([c.to(device=device, dtype=torch.float32) for c in cols], dim=1)
        assert qkv.shape == (D, lay.n_qkv), (qkv.shape, lay.n_qkv)
        put(ll.qkv_w, qkv)
        put(ll.qkv_b, torch.cat([b_.float() for b_ in bias]))
        put(ll.out_w, sd[f'{pa}.1.to_out.weight'].T)
        put(ll.ln2_gamma, sd[f'{pf}.0.gamma'])
        put(ll.ff1_w, sd[f'{pf}.1.ff.0.0.weight'].T)
        put(ll.ff1_b, sd[f'{pf}.1.ff.0.0.bias'])
        put(ll.ff2_w, sd[f'{pf}.1.ff.2.weight'].T)
        put(ll.ff2_b, sd[f'{pf}.1.ff.2.bias'])
    return blob


class WorldModelStepper:
    """Owns the packed weights and the KV cache of `num_envs` envs and advances all of them one
    timestep per call (x_transformers_rl.py:1250-1289, vectorised)."""

    def __init__(self, cfg: ModelConfig, num_envs: int, device='cuda'):
        self.lib = L.load()
        self.cfg, self.num_envs, self.device = cfg, num_envs, torch.device(device)
        self.desc = cfg.c_struct()
        nbytes = self.lib.xtrl_wm_kv_cache_bytes(C.byref(self.desc), num_envs)
        if nbytes < 0:
            L.check(self.lib.xtrl_wm_weight_layout(C.byref(self.desc), C.byref(L.WeightLayout())), 'model config')
        self.kv_cache = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
        self.weights = None
        self.tc_weights = None
        self.pos = torch.zeros(num_envs, device=self.device, dtype=torch.int32)

    def load_state_dict(self, sd):
        self.weights = pack_weights(sd, self.cfg, out=self.weights, device=self.device)
        if self.cfg.dense_bf16:                    # bf16 UMMA operand images, converted on the device
            if self.tc_weights is None:
                n = self.lib.xtrl_wm_tc_weights_bytes(C.byref(self.desc))
                self.tc_weights = torch.empty(n, device=self.device, dtype=torch.uint8)
            L.check(self.lib.xtrl_wm_pack_tc(C.byref(self.desc), _ptr(self.weights), _ptr(self.tc_weights), _stream()),
                    'xtrl_wm_pack_tc')

    def reset(self):
        self.pos.zero_()

    def step(self, state, prev_action, prev_reward, rs_mean, rs_var, *, uniform=None, forced_action=None,
             genes=None, gene_ids=None, reward_flag=1., out=None, want_logits=False, advance=True, pos_uniform=None,
             normal=None):
        """state (B,S) f32; prev_action (B,) i32 (-1 none); prev_reward (B,) f32; uniform (B,) f32.
        pos_uniform: int -> every env is at that timestep (lock-step rollout; self.pos is not read or advanced).
        Returns dict(action i32, log_prob, value[, raw_actions, value_logits])."""
        if self.weights is None:
            raise L.XtrlError('WorldModelStepper: load_state_dict() first')
        B, cfg, dev = self.num_envs, self.cfg, self.device
        if state.shape != (B, cfg.state_dim):
            raise L.XtrlError(f'step: state shape {tuple(state.shape)} != {(B, cfg.state_dim)}')
        out = {} if out is None else out
        cont, A = cfg.continuous, cfg.num_actions
        # continuous: prev_action / forced_action / action are float (B, A), log_prob is (B, A), `normal` replaces `uniform`
        for name, dt, shape in (('action', torch.float32 if cont else torch.int32, (B, A) if cont else (B,)),
                                ('log_prob', torch.float32, (B, A) if cont else (B,)), ('value', torch.float32, (B,))):
            if name not in out:
                out[name] = torch.empty(shape, device=dev, dtype=dt)
        if want_logits:
            out.setdefault('raw_actions', torch.empty(B, 2 * A if cont else A, device=dev))
            out.setdefault('value_logits', torch.empty(B, cfg.num_bins, device=dev))
        io = L.StepIO(_ptr(state, torch.float32, 'state'), _ptr(None if cont else prev_action, torch.int32, 'prev_action'),
                      _ptr(prev_reward, torch.float32, 'prev_reward'), _ptr(self.pos if pos_uniform is None else None, torch.int32),
                      _ptr(rs_mean, torch.float32, 'rs_mean'), _ptr(rs_var, torch.float32, 'rs_var'),
                      _ptr(genes, torch.float32, 'genes'), _ptr(gene_ids, torch.int32, 'gene_ids'),
                      _ptr(uniform, torch.float32, 'uniform'), _ptr(None if cont else forced_action, torch.int32, 'forced_action'),
                      float(reward_flag), int(pos_uniform or 0), _ptr(self.tc_weights), _ptr(None if cont else out['action']),
                      _ptr(out['log_prob']), _ptr(out['value']), _ptr(out.get('raw_actions')), _ptr(out.get('value_logits')),
                      _ptr(prev_action if cont else None, torch.float32, 'prev_action (continuous)'),
                      _ptr(normal if cont else None, torch.float32, 'normal'),
                      _ptr(forced_action if cont else None, torch.float32, 'forced_action (continuous)'),
                      _ptr(out['action'] if cont else None))
        L.check(self.lib.xtrl_wm_step(C.byref(self.desc), _ptr(self.weights), _ptr(self.kv_cache), C.byref(io), B,
                                      _stream()), 'xtrl_wm_step')
        if advance and pos_uniform is None:
            self.pos.add_(1)
        return out
