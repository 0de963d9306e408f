"""Learn-side (teacher-forced, full-sequence) world-model / actor-critic.

Host-side plumbing for `Agent.learn`.  Two expressions of the same model live here: the plain torch-op one (the fp32
class, pinned to the reference's goldens on CPU and to the oracle's autograd on the GPU), and -- `forward(fused=True)`,
the bf16 class -- the same layers as libxtrl tcgen05 kernels strung together by small `torch.autograd.Function`s
(attention projections, attention core, output projection, feed-forward, final norm, head MLPs,
forward and backward; csrc/learn_*.cu).  The losses with their backward, GAE, the optimiser and the rollout step run
in libxtrl.so in both classes.  Parameter names and shapes equal the reference's
`WorldModelActorCritic.state_dict()` (x_transformers_rl.py:281-369 with x-transformers'
ContinuousTransformerWrapper/Decoder underneath, :721-734) so `Agent.save/load`
checkpoints interchange and `ops.pack_weights` can feed the rollout kernel from the same dict.

All parameters are views into ONE flat fp32 buffer (and one flat gradient buffer): the optimiser
kernel, the EMA copy and the NCCL gradient all-reduce each touch a single contiguous tensor.
"""
from __future__ import annotations


import torch
import torch.nn.functional as F
from torch import nn

from . import ops


class _Norm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return F.layer_norm(x, x.shape[-1:]) * self.gamma


class _Rotary(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.register_buffer('inv_freq', 1. / (10000 ** (torch.arange(0, dim, 2).float() / dim)))


class _SafeEmbedding(nn.Module):
    """embedding that yields zeros for id -1 (x_transformers_rl.py:181-195)."""

    def __init__(self, n, dim):
        super().__init__()
        self.embed = nn.Embedding(n, dim)

    def forward(self, ids):
        w = self.embed.weight
        if ids.is_cuda and torch.is_grad_enabled() and w.requires_grad and w.shape[0] <= 16 and w.shape[1] == 48:
            return _EmbedRows.apply(ids, w)          # learn pass: same rows, weight gradient by one pass over dy
        has = ids >= 0
        return self.embed(ids.clamp(min=0)) * has[..., None].to(w.dtype)


class _Attention(nn.Module):
    def __init__(self, dim, heads, dim_head, gate, mix):
        super().__init__()
        inner = heads * dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)
        if gate:
            self.to_v_gate = nn.Linear(dim, inner)
            nn.init.constant_(self.to_v_gate.weight, 0.)
            nn.init.constant_(self.to_v_gate.bias, 10.)
        if mix:
            self.to_value_residual_mix = nn.Sequential(nn.Linear(dim, heads), nn.Sigmoid())


class _FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.ff = nn.Sequential(nn.Sequential(nn.Linear(dim, dim * mult), nn.GELU()), nn.Identity(),
                                nn.Linear(dim * mult, dim))


class _Decoder(nn.Module):
    def __init__(self, cfg: ops.ModelConfig):
        super().__init__()
        self.dim = cfg.dim
        self.rotary_pos_emb = _Rotary(cfg.dim_head // 2)
        self.layers = nn.ModuleList()
        for l in range(cfg.depth):
            self.layers.append(nn.ModuleList([_Norm(cfg.dim), _Attention(cfg.dim, cfg.heads, cfg.dim_head, cfg.gate_values,
                                                                          cfg.learned_mix and l > 0)]))
            self.layers.append(nn.ModuleList([_Norm(cfg.dim), _FeedForward(cfg.dim)]))
        self.final_norm = _Norm(cfg.dim)


class _Wrapper(nn.Module):
    def __init__(self, cfg: ops.ModelConfig):
        super().__init__()
        self.project_in = nn.Linear(cfg.state_dim, cfg.dim, bias=False)
        self.attn_layers = _Decoder(cfg)


def _rotate(t, cos, sin, rot):
    x, rest = t[..., :rot], t[..., rot:]
    pairs = x.unflatten(-1, (-1, 2))
    half = torch.stack((-pairs[..., 1], pairs[..., 0]), dim=-1).flatten(-2)
    return torch.cat((x * cos + half * sin, rest), dim=-1)


class _FusedFeedForward(torch.autograd.Function):
    """autograd seam around xtrl_ff_fwd / xtrl_ff_bwd (csrc/learn_ff.cu): x -> x + FF(LN(x)).  Only x is saved; the
    backward recomputes on the tensor cores and ADDS the parameter gradients straight into the parameters' .grad
    views of the flat gradient buffer (so the parameters are not autograd inputs here)."""

    @staticmethod
    def forward(ctx, x, norm, ff, img, rng_step, key, p_drop):
        lin1, lin2 = ff.ff[0][0], ff.ff[2]
        x = x.float().contiguous()
        ctx.save_for_backward(x)
        ctx.mods, ctx.img, ctx.rng = (norm, lin1, lin2), img, (rng_step, key, p_drop)
        return ops.ff_fwd(x, norm.gamma.data, lin1.bias.data, lin2.bias.data, img, dropout=p_drop, rng_step=rng_step, rng_key=key)

    @staticmethod
    def backward(ctx, dy):
        (x,), (norm, lin1, lin2), (rng_step, key, p_drop) = ctx.saved_tensors, ctx.mods, ctx.rng
        dx = ops.ff_bwd(x, dy.float().contiguous(), norm.gamma.data, lin1.bias.data, ctx.img, norm.gamma.grad, lin1.weight.grad,
                        lin1.bias.grad, lin2.weight.grad, lin2.bias.grad, dropout=p_drop, rng_step=rng_step, rng_key=key)
        return dx, None, None, None, None, None, None


class _EmbedRows(torch.autograd.Function):
    """rows of an embedding table by id (-1 -> zeros); the weight gradient is one xtrl_small_dw pass over dy instead of
    torch's sort-based embedding backward."""

    @staticmethod
    def forward(ctx, ids, weight):
        ctx.save_for_backward(ids)
        ctx.shape = weight.shape
        return F.embedding(ids.clamp(min=0), weight) * (ids >= 0)[..., None].to(weight.dtype)

    @staticmethod
    def backward(ctx, dy):
        (ids,) = ctx.saved_tensors
        dw = torch.zeros(ctx.shape, dtype=torch.float32, device=dy.device)
        return None, ops.small_dw(dy.float().contiguous(), dw, ids=ids.long().contiguous(), out_kj=True)


class _SkinnyLinear(torch.autograd.Function):
    """fp32 Linear from a few input features (no gradient to the input) with the weight / bias gradient as one
    xtrl_small_dw pass."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.float().contiguous()
        ctx.save_for_backward(x)
        ctx.shape = weight.shape
        return F.linear(x, weight.float(), bias.float())

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dw = torch.zeros(ctx.shape, dtype=torch.float32, device=dy.device)
        db = torch.zeros(ctx.shape[0], dtype=torch.float32, device=dy.device)
        ops.small_dw(dy.float().contiguous(), dw, feats=x, d_b=db)
        return None, dw, db


class _FusedPrep(torch.autograd.Function):
    """autograd seam around xtrl_learn_prep_fwd / _bwd (csrc/learn_prep.cu): the minibatch gather + RSNorm + input embeddings.
    Returns (x0, se, ae_next); everything else the prologue produced is in `out` (not differentiable).  The parameters are not
    autograd inputs: their gradients are accumulated straight into the .grad views of the flat gradient buffer."""

    @staticmethod
    def forward(ctx, anchor, model, src, idx, rs_mean, rs_var, flag, out_box):
        cont = model.cfg.continuous
        ae = model.action_embeds
        out = ops.learn_prep_fwd(src, idx, rs_mean, rs_var, model.transformer.project_in.weight.data, model.to_state_embed.weight.data,
                                 model.to_state_embed.bias.data, ae.weight.data if cont else ae.embed.weight.data,
                                 ae.bias.data if cont else None, model.reward_embed.data, flag, continuous=cont)
        out_box.update(out)
        ctx.model, ctx.flag = model, flag
        ctx.save_for_backward(out['packed'], out['actions'])
        return out['x0'], out['se'], out['ae_next']

    @staticmethod
    def backward(ctx, dx0, d_se, d_ae):
        packed, acts = ctx.saved_tensors
        model = ctx.model
        cont = model.cfg.continuous
        ae = model.action_embeds
        c = lambda t: None if t is None else t.float().contiguous()
        ops.learn_prep_bwd(c(dx0), c(d_se), c(d_ae), packed, acts, ctx.flag, model.transformer.project_in.weight.grad,
                           model.to_state_embed.weight.grad, model.to_state_embed.bias.grad, ae.weight.grad if cont else ae.embed.weight.grad,
                           ae.bias.grad if cont else None, model.reward_embed.grad, continuous=cont)
        return None, None, None, None, None, None, None, None


class _FracGrad(torch.autograd.Function):
    """forward identity, backward x frac (x_transformers_rl.py:114-116) without the detach / scale / add round trip."""

    @staticmethod
    def forward(ctx, x, frac):
        ctx.frac = frac
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.frac, None


class _Head:
    """device state of one fused head MLP (csrc/learn_mlp.cu): zero-padded dense copies of the head's weights, their
    operand image, and padded gradient scratch; `refresh` follows the live parameters, `scatter` adds the scratch back
    into the parameters' .grad views.  `rows1` / `rows2`: (parameter-weight, bias, first padded row, rows) of each layer."""

    def __init__(self, in_dim, hid_p, out_p, out_cols, n_silu, dev):
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.in_dim, self.hid_p, self.out_p, self.out_cols, self.n_silu = in_dim, hid_p, out_p, out_cols, n_silu
        self.w1, self.b1, self.w2, self.b2 = z(hid_p, in_dim), z(hid_p), z(out_p, hid_p), z(out_p)
        # the padded gradient scratch is ONE buffer (one fill per backward) viewed as the four tensors
        sizes = (hid_p * in_dim, hid_p, out_p * hid_p, out_p)
        self.gflat = z(sum(sizes))
        parts = self.gflat.split(sizes)
        self.gw1, self.gb1, self.gw2, self.gb2 = parts[0].view(hid_p, in_dim), parts[1], parts[2].view(out_p, hid_p), parts[3]
        self.img = torch.empty(ops.L.load().xtrl_mlp_img_bytes(in_dim, hid_p, out_p), dtype=torch.uint8, device=dev)
        self.l1, self.l2 = [], []          # (linear, first row) ; (linear, first row, hidden columns)
        self._tables = None

    def _seg_tables(self):
        """(refresh, scatter) block tables over the CURRENT addresses of the parameters' data / grad views (they live in the
        model's flat buffers; rebuilt if those move)"""
        key = tuple((lin.weight.data_ptr(), lin.weight.grad.data_ptr() if lin.weight.grad is not None else 0) for lin, *_ in self.l1 + self.l2)
        if self._tables is None or self._tables[0] != key:
            fwd, bwd = [], []
            for lin, r0 in self.l1:
                n = lin.weight.shape[0]
                fwd += [(lin.weight.data, self.w1[r0:r0 + n]), (lin.bias.data, self.b1[r0:r0 + n])]
                bwd += [(self.gw1[r0:r0 + n], lin.weight.grad), (self.gb1[r0:r0 + n], lin.bias.grad)]
            for lin, r0, cols in self.l2:
                n = lin.weight.shape[0]
                fwd += [(lin.weight.data, self.w2[r0:r0 + n, :cols]), (lin.bias.data, self.b2[r0:r0 + n])]
                bwd += [(self.gw2[r0:r0 + n, :cols], lin.weight.grad), (self.gb2[r0:r0 + n], lin.bias.grad)]
            self._tables = (key, ops.SegTable(fwd), ops.SegTable(bwd))
        return self._tables[1], self._tables[2]

    def refresh(self):
        self._seg_tables()[0].run()                              # one launch: every parameter block into the padded copies
        ops.mlp_pack(self.w1, self.w2, out=self.img)

    def scatter(self):
        self._seg_tables()[1].run(accumulate=True)               # one launch: padded gradient blocks += into the .grad views


class _FusedHead(torch.autograd.Function):
    """(src0, src1[, src2 per episode]) -> head output (b, n, out_cols) bf16 through xtrl_mlp_fwd / _bwd."""

    @staticmethod
    def forward(ctx, head, n, src0, src1, src2):
        srcs = [t.float().contiguous() for t in (src0, src1, src2) if t is not None]
        ctx.save_for_backward(*srcs)
        ctx.head, ctx.n = head, n
        y = ops.mlp_fwd(srcs, n, head.b1, head.b2, head.img, head.out_cols, head.n_silu, per_episode2=len(srcs) == 3)
        return y.view(-1, n, head.out_cols)

    @staticmethod
    def backward(ctx, gy):
        srcs, head, n = list(ctx.saved_tensors), ctx.head, ctx.n
        head.gflat.zero_()
        gy = gy.reshape(-1, head.out_cols).to(torch.bfloat16).contiguous()
        d = ops.mlp_bwd(srcs, n, gy, head.b1, head.img, head.gw1, head.gb1, head.gw2, head.gb2, head.n_silu, per_episode2=len(srcs) == 3)
        head.scatter()
        b = srcs[0].numel() // (48 * n)
        d2 = d[2].view(b, n, 48).sum(1) if len(srcs) == 3 else None
        return None, None, d[0].view(b, n, 48), d[1].view(b, n, 48), d2


class _FusedHeads3(torch.autograd.Function):
    """the three heads of a pass without a latent input as ONE autograd node: world-model head on (embed, ae_next), actor and
    critic on (frac_gradient(embed), se).  The backward runs the three xtrl_mlp_bwd launches into shared gradient buffers:
    the world-model head writes d_embed, the other two add frac * dX (frac_gradient, x_transformers_rl.py:114-116, :533) into
    it and share d_se -- no per-head gradient tensors, no adds, no separate scaling pass."""

    @staticmethod
    def forward(ctx, heads, n, frac, embed, se, ae_next):
        embed, se, ae_next = (t.float().contiguous() for t in (embed, se, ae_next))
        ctx.save_for_backward(embed, se, ae_next)
        ctx.heads, ctx.n, ctx.frac = heads, n, float(frac)
        run = lambda h, srcs: ops.mlp_fwd(srcs, n, h.b1, h.b2, h.img, h.out_cols, h.n_silu).view(-1, n, h.out_cols)
        sp_done = run(heads['pred'], [embed, ae_next])
        n_pred = heads['pred'].out_cols - 1                          # the world-model head's last column is the done logit
        # the two dense tensors the loss kernel reads (it would copy the slices anyway); their gradients come back as two
        # dense tensors and are joined by one cat instead of autograd's zero-fill + copy + zero-fill + copy + add
        return run(heads['actor'], [embed, se]), run(heads['critic'], [embed, se]), sp_done[..., :n_pred].contiguous(), sp_done[..., n_pred].contiguous()

    @staticmethod
    def backward(ctx, g_actor, g_critic, g_sp, g_dl):
        g_pred = torch.cat([g_sp.reshape(-1, g_sp.shape[-1]), g_dl.reshape(-1, 1)], dim=1)
        embed, se, ae_next = ctx.saved_tensors
        heads, n, frac = ctx.heads, ctx.n, ctx.frac
        d_embed, d_se, d_ae = torch.empty_like(embed), torch.empty_like(se), torch.empty_like(ae_next)
        flat = lambda t: t.view(-1, 48)

        def run(h, gy, srcs, outs, scales, acc):
            h.gflat.zero_()
            gy = gy.reshape(-1, h.out_cols).to(torch.bfloat16).contiguous()
            ops.mlp_bwd(srcs, n, gy, h.b1, h.img, h.gw1, h.gb1, h.gw2, h.gb2, h.n_silu, out=[flat(o) for o in outs], scales=scales, accumulate=acc)
            h.scatter()

        run(heads['pred'], g_pred, [embed, ae_next], (d_embed, d_ae), (1., 1.), 0)
        run(heads['actor'], g_actor, [embed, se], (d_embed, d_se), (frac, 1.), 1)
        run(heads['critic'], g_critic, [embed, se], (d_embed, d_se), (frac, 1.), 3)
        return None, None, None, d_embed, d_se, d_ae


class _FusedNorm(torch.autograd.Function):
    """the Decoder's final LayerNorm through xtrl_ln_fwd / _bwd (csrc/learn_ln.cu)."""

    @staticmethod
    def forward(ctx, x, norm):
        x = x.float().contiguous()
        ctx.save_for_backward(x)
        ctx.norm = norm
        return ops.ln_fwd(x, norm.gamma.data)

    @staticmethod
    def backward(ctx, dy):
        (x,), norm = ctx.saved_tensors, ctx.norm
        return ops.ln_bwd(x, dy.float().contiguous(), norm.gamma.data, norm.gamma.grad), None


class _FusedAttnPre(torch.autograd.Function):
    """x -> q, k, v (b, H, n, 16) bf16 [, value-gate pre-activation (b, H, n, 16)] [, value-residual mix logits (b, n, 16)]
    through xtrl_attn_pre_fwd / _bwd (csrc/learn_attn.cu): LayerNorm, the projections, partial rotary, head-major layout.
    Returns only the LayerNorm path's dx; the residual path of x is autograd's.  Parameter gradients go straight into the
    flat gradient buffer, as in _FusedFeedForward."""

    @staticmethod
    def forward(ctx, x, norm, attn, img, cos, sin, heads, gate, mix):
        x = x.float().contiguous()
        ctx.save_for_backward(x)
        ctx.mods, ctx.aux, ctx.flags = (norm, attn), (img, cos, sin), (gate, mix)
        return ops.attn_pre_fwd(x, norm.gamma.data, img, cos, sin, heads,
                                gate_bias=attn.to_v_gate.bias.data if gate else None,
                                mix_bias=attn.to_value_residual_mix[0].bias.data if mix else None)

    @staticmethod
    def backward(ctx, dq, dk, dv, *extra):
        (x,), (norm, attn), (img, cos, sin), (gate, mix) = ctx.saved_tensors, ctx.mods, ctx.aux, ctx.flags
        extra = list(extra)
        dg = extra.pop(0).contiguous() if gate else None
        dmix = extra.pop(0).contiguous() if mix else None
        lin_m = attn.to_value_residual_mix[0] if mix else None
        dx = ops.attn_pre_bwd(x, dq.contiguous(), dk.contiguous(), dv.contiguous(), norm.gamma.data, img, cos, sin,
                              norm.gamma.grad, attn.to_q.weight.grad, attn.to_k.weight.grad, attn.to_v.weight.grad,
                              dg=dg, dmix=dmix, d_wg=attn.to_v_gate.weight.grad if gate else None,
                              d_bg=attn.to_v_gate.bias.grad if gate else None,
                              d_wm=lin_m.weight.grad if mix else None, d_bm=lin_m.bias.grad if mix else None)
        return dx, None, None, None, None, None, None, None, None


class _FusedAttnCore(torch.autograd.Function):
    """causal softmax(q k^T / 4) v with attention dropout for dim_head 16, n <= 512 (csrc/learn_attn_core.cu)."""

    @staticmethod
    def forward(ctx, q, k, v, rng_step, key, p_drop):
        o, lse = ops.attn_core_fwd(q, k, v, dropout=p_drop, rng_step=rng_step, rng_key=key)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.rng = (rng_step, key, p_drop)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse = ctx.saved_tensors
        rng_step, key, p_drop = ctx.rng
        dq, dk, dv = ops.attn_core_bwd(q, k, v, o, lse, d_o.contiguous(), dropout=p_drop, rng_step=rng_step, rng_key=key)
        return dq, dk, dv, None, None, None


class _FusedAttnPost(torch.autograd.Function):
    """(x, o) -> x + mask * to_out(o) through xtrl_attn_post_fwd / _bwd; o (b, H, n, 16) bf16 from the attention core."""

    @staticmethod
    def forward(ctx, x, o, attn, img, lens):
        x, o = x.float().contiguous(), o.contiguous()
        ctx.save_for_backward(o, lens)
        ctx.attn, ctx.img = attn, img
        return ops.attn_post_fwd(x, o, lens, img)

    @staticmethod
    def backward(ctx, dy):
        o, lens = ctx.saved_tensors
        dy = dy.float().contiguous()
        return dy, ops.attn_post_bwd(dy, o, lens, ctx.img, ctx.attn.to_out.weight.grad), None, None, None


class _FusedAttnBlock(torch.autograd.Function):
    """the whole plain attention block y = x + mask * to_out(attn(rotary(to_qkv(LN(x))))) (no value gate / value residual) as ONE
    autograd node over the same kernels as _FusedAttnPre / _FusedAttnCore / _FusedAttnPost.  Its point is the backward: the
    gradient arriving over the residual path (dy) is added inside xtrl_attn_pre_bwd, so autograd has no (b, n, 48) add to do
    per layer."""

    @staticmethod
    def forward(ctx, x, norm, attn, img, cos, sin, heads, lens, rng_step, key, p_drop):
        x = x.float().contiguous()
        q, k, v = ops.attn_pre_fwd(x, norm.gamma.data, img, cos, sin, heads)[:3]
        o, lse = ops.attn_core_fwd(q, k, v, dropout=p_drop, rng_step=rng_step, rng_key=key)
        ctx.save_for_backward(x, q, k, v, o, lse, lens)
        ctx.mods, ctx.aux, ctx.rng = (norm, attn), (img, cos, sin), (rng_step, key, p_drop)
        return ops.attn_post_fwd(x, o, lens, img)

    @staticmethod
    def backward(ctx, dy):
        x, q, k, v, o, lse, lens = ctx.saved_tensors
        (norm, attn), (img, cos, sin), (rng_step, key, p_drop) = ctx.mods, ctx.aux, ctx.rng
        dy = dy.float().contiguous()
        d_o = ops.attn_post_bwd(dy, o, lens, img, attn.to_out.weight.grad)
        dq, dk, dv = ops.attn_core_bwd(q, k, v, o, lse, d_o, dropout=p_drop, rng_step=rng_step, rng_key=key)
        dx = ops.attn_pre_bwd(x, dq, dk, dv, norm.gamma.data, img, cos, sin, norm.gamma.grad, attn.to_q.weight.grad,
                              attn.to_k.weight.grad, attn.to_v.weight.grad, d_res=dy)
        return (dx,) + (None,) * 10


class WorldModelActorCritic(nn.Module):
    def __init__(self, cfg: ops.ModelConfig, dropout=0.25, frac_actor_critic_head_gradient=0.5, reward_dropout=0.5):
        super().__init__()
        self.cfg, self.dropout, self.frac, self.reward_dropout_p = cfg, dropout, frac_actor_critic_head_gradient, reward_dropout
        D, S, A, K = cfg.dim, cfg.state_dim, cfg.num_actions, cfg.num_bins
        self.transformer = _Wrapper(cfg)
        self.reward_embed = nn.Parameter(torch.ones(D) * 1e-2)
        self.action_embeds = nn.Linear(A, D) if cfg.continuous else _SafeEmbedding(A, D)      # :306-309
        self.to_state_embed = nn.Linear(S, D)
        self.to_pred_done = nn.Sequential(nn.Linear(2 * D, 1))
        self.to_pred = nn.Sequential(nn.Linear(2 * D, D), nn.SiLU(), nn.Linear(D, 2 * (S + 1)))
        self.evolutionary = cfg.dim_gene > 0
        if self.evolutionary:
            self.latent_to_embed = nn.Linear(cfg.dim_gene, D)
        hin = (3 if self.evolutionary else 2) * D
        self.critic_head = nn.Sequential(nn.Linear(hin, 2 * D), nn.SiLU(), nn.Linear(2 * D, K))
        self.action_head = nn.Sequential(nn.Linear(hin, 2 * D), nn.SiLU(), nn.Linear(2 * D, 2 * A if cfg.continuous else A))
        self._flatten()

    # -- one flat fp32 buffer behind every parameter ------------------------------------------
    def _flatten(self):
        params = list(self.parameters())
        sizes = [p.numel() for p in params]
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        flat = torch.empty(offs[-1], dtype=torch.float32, device=params[0].device)
        for p, o in zip(params, offs):
            flat[o:o + p.numel()].copy_(p.data.reshape(-1))
        self._offsets = offs
        self._bind(flat, torch.zeros_like(flat))
        self._fused = None                       # lazily built device state of the fused learn-side kernels

    def fused_state(self):
        """per-layer operand images, head state and the dropout counter of the tcgen05 learn-side kernels.  A shape
        those kernels are not built for raises XtrlError -- the bf16 class never falls back to another back end."""
        if self._fused is None:
            cfg, dev = self.cfg, self.flat_params.device
            if dev.type != 'cuda':
                raise ops.L.XtrlError('the tensor-core (bf16 class) learn pass needs CUDA parameters: there is no CPU path')
            lib = ops.L.load()
            mix = lambda l: bool(cfg.value_residual and cfg.learned_mix and l > 0)
            A_out = self.action_head[2].weight.shape[0]
            kin = (3 if self.evolutionary else 2) * cfg.dim
            n_pred = 2 * (cfg.state_dim + 1)
            problems = []
            if cfg.dim != 48 or not ops.ff_supported(cfg.dim, 4 * cfg.dim):
                problems.append(f'hidden_dim={cfg.dim} (built: 48)')
            elif not all(ops.attn_supported(cfg.dim, cfg.heads, cfg.dim_head, cfg.gate_values, mix(l)) for l in range(cfg.depth)):
                problems.append(f'heads={cfg.heads}, dim_head={cfg.dim_head} (built: dim_head 16, heads in (4, 8))')
            if cfg.num_bins > 112:
                problems.append(f'critic bins={cfg.num_bins} (built: <= 112)')
            if A_out > 16:
                problems.append(f'actor head width={A_out} (built: <= 16)')
            if n_pred + 1 > 32 or cfg.state_dim > 16:
                problems.append(f'state_dim={cfg.state_dim} (built: <= 15)')
            if not problems and not (ops.mlp_supported(kin, 96, 112) and ops.mlp_supported(96, 64, 32)):
                problems.append(f'head input width {kin}')
            if problems:
                raise ops.L.XtrlError('the tensor-core (bf16 class) learn kernels are not built for ' + '; '.join(problems) +
                                      ". Pass accelerate_kwargs=dict(mixed_precision='no') for the fp32 parity class.")
            u8 = lambda nbytes: [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(cfg.depth)]
            attn_img = [torch.empty(lib.xtrl_attn_img_bytes(cfg.dim, cfg.heads, cfg.dim_head, int(cfg.gate_values), int(mix(l))),
                                    dtype=torch.uint8, device=dev) for l in range(cfg.depth)]
            critic = _Head(kin, 96, 112, cfg.num_bins, 96, dev)
            critic.l1, critic.l2 = [(self.critic_head[0], 0)], [(self.critic_head[2], 0, 96)]
            actor = _Head(kin, 96, 16, A_out, 96, dev)
            actor.l1, actor.l2 = [(self.action_head[0], 0)], [(self.action_head[2], 0, 96)]
            pred = _Head(96, 64, 32, n_pred + 1, 48, dev)      # hidden unit 48 = the (linear) done logit, routed to output n_pred
            pred.l1, pred.l2 = [(self.to_pred[0], 0), (self.to_pred_done[0], 48)], [(self.to_pred[2], 0, 48)]
            pred.w2[n_pred, 48] = 1.
            self._fused = dict(ff_img=u8(lib.xtrl_ff_img_bytes(cfg.dim, 4 * cfg.dim)), attn_img=attn_img,
                               heads=dict(critic=critic, actor=actor, pred=pred),
                               rng_step=torch.zeros((), dtype=torch.int64, device=dev))
        return self._fused

    def _bind(self, flat, flat_grad):
        self.flat_params, self.flat_grads = flat, flat_grad
        for p, o in zip(self.parameters(), self._offsets):
            p.data = flat[o:o + p.numel()].view(p.shape)
            p.grad = flat_grad[o:o + p.numel()].view(p.shape)
        self.seg_offsets = torch.tensor(self._offsets, dtype=torch.int64, device=flat.device)

    def _apply(self, fn, *a, **k):            # .to(device) must keep the flat-buffer aliasing
        flat, grad = fn(self.flat_params), fn(self.flat_grads)
        super()._apply(fn, *a, **k)
        self._bind(flat, grad)
        return self

    def state_dict_from_flat(self, flat):
        """a reference-style state_dict whose tensors view `flat` (e.g. the EMA buffer)."""
        out = {}
        for (name, p), o in zip(self.named_parameters(), self._offsets):
            out[name] = flat[o:o + p.numel()].view(p.shape)
        return out

    # -- full-sequence forward (x_transformers_rl.py:479-559 with mask=, :928-935) ---------------
    def forward(self, state, rewards, actions, next_actions, mask, latent_gene=None, reward_flag=1., fused=False):
        """state (b,n,S) normalised; rewards (b,n) normalised prev rewards; actions (b,n) prev actions
        (-1 none; continuous: float (b,n,A), zeros for none); next_actions likewise; mask (b,n) bool.
        Returns raw head outputs: raw_actions (b,n,A | 2A), value_logits (b,n,K), state_pred_raw (b,n,2(S+1)),
        done_logit (b,n)."""
        cfg, tfm = self.cfg, self.transformer
        dec = tfm.attn_layers
        b, n, _ = state.shape
        h, dh, rot = cfg.heads, cfg.dim_head, cfg.dim_head // 2
        p_drop = self.dropout if self.training else 0.
        x = tfm.project_in(state) + self.action_embeds(actions) + rewards[..., None] * self.reward_embed * reward_flag
        ang = torch.arange(n, device=state.device, dtype=torch.float32)[:, None] * dec.rotary_pos_emb.inv_freq[None]
        ang = ang.repeat_interleave(2, dim=-1)                       # (n, rot) interleaved (t0,t0,t1,t1,..)
        cos, sin = ang.cos().to(x.dtype), ang.sin().to(x.dtype)
        # Episodes are right-padded, so every padded key lies in the causal future of every valid query:
        # the causal flag alone reproduces the reference's key-padding + causal mask on valid rows, and
        # padded rows are zeroed after to_out exactly as the reference does (App. A.1) -- no mask tensor.
        if fused:
            return self._forward_fused(x, state, next_actions, mask, latent_gene, ang, p_drop)
        first_v = None
        for l in range(cfg.depth):
            norm, attn = dec.layers[2 * l]
            xn = norm(x)
            split = lambda t: t.view(b, n, h, dh).transpose(1, 2)
            q, k, v = split(attn.to_q(xn)), split(attn.to_k(xn)), split(attn.to_v(xn))
            if cfg.value_residual:
                if l == 0:
                    first_v = v
                elif cfg.learned_mix:
                    mix = attn.to_value_residual_mix(xn).transpose(1, 2)[..., None]
                    v = first_v.lerp(v, mix)
                else:
                    v = first_v * 0.5 + v * 0.5
            q, k = _rotate(q, cos, sin, rot), _rotate(k, cos, sin, rot)
            out = F.scaled_dot_product_attention(q, k, v, is_causal=True, dropout_p=p_drop)
            out = out.transpose(1, 2).reshape(b, n, h * dh)
            if cfg.gate_values:
                out = out * attn.to_v_gate(xn).sigmoid()
            out = attn.to_out(out).masked_fill(~mask[..., None], 0.)
            x = x + out
            norm, ff = dec.layers[2 * l + 1]
            hid = ff.ff[0](norm(x))
            hid = F.dropout(hid, p_drop, self.training)
            x = x + ff.ff[2](hid)
        embed = dec.final_norm(x)

        ewa = torch.cat((embed, self.action_embeds(next_actions)), dim=-1)          # :516-529
        state_pred_raw = self.to_pred(ewa)
        done_logit = self.to_pred_done(ewa)[..., 0]

        embed = embed.detach() * (1. - self.frac) + embed * self.frac             # frac_gradient :114-116,:533
        ac_in = torch.cat((embed, self.to_state_embed(state)), dim=-1)
        if self.evolutionary:
            lat = self.latent_to_embed(latent_gene)
            ac_in = torch.cat((ac_in, lat[:, None].expand(-1, n, -1).to(ac_in.dtype)), dim=-1)
        return self.action_head(ac_in), self.critic_head(ac_in), state_pred_raw, done_logit


    def _forward_fused(self, x, state, next_actions, mask, latent_gene, ang, p_drop):
        """the bf16 class from torch-made inputs (model-level tests): input embeddings by torch ops, then the fused trunk / heads."""
        n = state.shape[1]
        lens32 = mask.sum(dim=1, dtype=torch.int32)          # episodes are right-padded: mask == position < length
        with torch.autocast('cuda', enabled=False):
            se32 = _SkinnyLinear.apply(state, self.to_state_embed.weight, self.to_state_embed.bias)
            lat = self.latent_to_embed(latent_gene.float()) if self.evolutionary else None
        embed = self._fused_trunk(x, lens32, n, ang, p_drop)
        return self._fused_heads(embed, se32, self.action_embeds(next_actions), lat, n)

    def forward_prepared(self, src, idx, rs_mean, rs_var, flag, latent_gene_fn=None):
        """the bf16 class as Agent.learn runs it: prologue kernel (gather + RSNorm + input embeddings, csrc/learn_prep.cu) ->
        fused trunk -> fused heads.  Returns (raw_actions, value_logits, state_pred_raw, done_logit, prologue outputs dict)."""
        cfg = self.cfg
        n = src['states'].shape[1]
        box = {}
        x0, se, ae_next = _FusedPrep.apply(self.reward_embed, self, src, idx, rs_mean, rs_var, flag, box)
        lat = None
        if self.evolutionary:
            lat = self.latent_to_embed(latent_gene_fn(box['gene_ids'].long()).float())
        p_drop = self.dropout if self.training else 0.
        embed = self._fused_trunk(x0, box['lens'], n, None, p_drop)
        return (*self._fused_heads(embed, se, ae_next, lat, n), box)

    def _rotary_tables(self, n, ang=None):
        """cos / sin of position * inv_freq for the kernels' partial rotary, [n][dim_head / 4] each.  inv_freq is a constant
        buffer, so the tables are computed once per sequence length (outside graph capture) instead of six small launches
        per pass; `ang` (n, dim_head / 2, every frequency twice) is what callers that already formed the angles pass in."""
        cache = self.fused_state().setdefault('rotary', {})
        if n in cache:
            return cache[n]
        if ang is None:
            ang = torch.arange(n, device=self.flat_params.device, dtype=torch.float32)[:, None] * self.transformer.attn_layers.rotary_pos_emb.inv_freq[None]
            ang = ang.repeat_interleave(2, dim=-1)
        ang4 = ang[:, ::2].contiguous()
        tabs = (ang4.cos(), ang4.sin())
        if not torch.cuda.is_current_stream_capturing():
            cache[n] = tabs
        return tabs

    def _fused_trunk(self, x, lens32, n, ang, p_drop):
        """every sub-layer of the Decoder and the final norm as libxtrl tcgen05 kernels behind autograd seams (csrc/learn_*.cu).
        Unsupported shapes raise (fused_state / attention core)."""
        cfg, dec = self.cfg, self.transformer.attn_layers
        h, dh = cfg.heads, cfg.dim_head
        fs = self.fused_state()
        if not ops.attn_core_supported(n, dh):
            raise ops.L.XtrlError(f'the tensor-core attention core is built for dim_head=16 and sequences up to 512 (got dim_head={dh}, '
                                  f"n={n}). Pass accelerate_kwargs=dict(mixed_precision='no') for the fp32 parity class.")
        fs['rng_step'].add_(1)                    # one counter bump per pass: fresh dropout masks on every graph replay
        # bf16 operand images follow the current weights: one launch for the feed-forward layers, one per attention variant
        # (a learned value-residual mix exists from layer 1 on)
        ffs, ats = [dec.layers[2 * l + 1][1] for l in range(cfg.depth)], [dec.layers[2 * l][1] for l in range(cfg.depth)]
        ops.ff_pack_layers([f.ff[0][0].weight.data for f in ffs], [f.ff[2].weight.data for f in ffs], fs['ff_img'])
        mixed = lambda l: bool(cfg.value_residual and cfg.learned_mix and l > 0)
        for group in ([l for l in range(cfg.depth) if not mixed(l)], [l for l in range(cfg.depth) if mixed(l)]):
            if group:
                g = [ats[l] for l in group]
                ops.attn_pack_layers([t.to_q.weight.data for t in g], [t.to_k.weight.data for t in g], [t.to_v.weight.data for t in g],
                                     [t.to_out.weight.data for t in g], [t.to_v_gate.weight.data for t in g] if cfg.gate_values else None,
                                     [t.to_value_residual_mix[0].weight.data for t in g] if mixed(group[0]) else None,
                                     [fs['attn_img'][l] for l in group])
        cos4, sin4 = self._rotary_tables(n, ang)
        first_v = None
        plain = not (cfg.gate_values or cfg.value_residual)
        for l in range(cfg.depth):
            norm, attn = dec.layers[2 * l]
            if plain:                                                # one autograd node per block (residual add inside pre_bwd)
                x = _FusedAttnBlock.apply(x, norm, attn, fs['attn_img'][l], cos4, sin4, h, lens32, fs['rng_step'], 1000 + l, p_drop)
                norm, ff = dec.layers[2 * l + 1]
                x = _FusedFeedForward.apply(x, norm, ff, fs['ff_img'][l], fs['rng_step'], l, p_drop)
                continue
            has_mix = bool(cfg.value_residual and cfg.learned_mix and l > 0)
            outs = _FusedAttnPre.apply(x, norm, attn, fs['attn_img'][l], cos4, sin4, h, cfg.gate_values, has_mix)
            q, k, v = outs[:3]
            if cfg.value_residual:                                   # the small element-wise tail stays with autograd
                if l == 0:
                    first_v = v
                elif cfg.learned_mix:
                    v = first_v.lerp(v, outs[-1][..., :h].sigmoid().transpose(1, 2)[..., None])
                else:
                    v = first_v * 0.5 + v * 0.5
            out = _FusedAttnCore.apply(q, k, v, fs['rng_step'], 1000 + l, p_drop)
            if cfg.gate_values:
                out = out * outs[3].sigmoid()
            x = _FusedAttnPost.apply(x, out, attn, fs['attn_img'][l], lens32)
            norm, ff = dec.layers[2 * l + 1]
            x = _FusedFeedForward.apply(x, norm, ff, fs['ff_img'][l], fs['rng_step'], l, p_drop)
        return _FusedNorm.apply(x, dec.final_norm)

    def _fused_heads(self, embed, se32, ae_next, lat, n):
        """the three head MLPs (critic, actor, world model + done) as tcgen05 kernels (csrc/learn_mlp.cu)"""
        cfg = self.cfg
        hd = self.fused_state()['heads']
        for hs in hd.values():
            hs.refresh()
        n_pred = 2 * (cfg.state_dim + 1)
        if lat is None:                                              # one node, shared gradient buffers (see _FusedHeads3)
            return _FusedHeads3.apply(hd, n, self.frac, embed, se32, ae_next)
        sp_done = _FusedHead.apply(hd['pred'], n, embed, ae_next, None)
        emb_ac = _FracGrad.apply(embed, self.frac)
        return (_FusedHead.apply(hd['actor'], n, emb_ac, se32, lat), _FusedHead.apply(hd['critic'], n, emb_ac, se32, lat),
                sp_done[..., :n_pred], sp_done[..., n_pred])


class FusedPPOLoss(torch.autograd.Function):
    """autograd seam around xtrl_ppo_loss_fwd_bwd: forward launches the fused kernel, which already
    produced d loss / d head-outputs; backward hands those (scaled by the incoming scalar) to autograd
    so the sequence model's own backward continues from them."""

    @staticmethod
    def forward(ctx, cfg, raw_actions, value_logits, state_pred_raw, done_logit, actions, old_log_probs, returns,
                old_values, target, dones, episode_lens, unit_grad=False):
        # unit_grad: the caller promises to call .backward() on the returned total itself (incoming gradient == 1), so the
        # stored gradients are handed over as they are instead of through four multiplications by one (47 us at c3)
        ctx.unit_grad = bool(unit_grad)
        scalars, grads = ops.ppo_loss_fwd_bwd(cfg, raw_actions.contiguous(), value_logits.contiguous(),
                                              state_pred_raw.contiguous(), done_logit.contiguous(), actions,
                                              old_log_probs, returns, old_values, target, dones, episode_lens)
        ctx.save_for_backward(grads['raw_actions'], grads['value_logits'], grads['state_pred_raw'], grads['done_logit'])
        ctx.mark_non_differentiable(scalars)
        return scalars[0].clone(), scalars

    @staticmethod
    def backward(ctx, g_total, _g_scalars):
        g_ra, g_vl, g_sp, g_dl = ctx.saved_tensors
        if ctx.unit_grad:
            return (None, g_ra, g_vl, g_sp, g_dl) + (None,) * 8
        s = g_total.to(g_ra.dtype)
        return (None, g_ra * s, g_vl * s, g_sp * s, g_dl * s) + (None,) * 8
