"""Clean-room stand-in for the slice of `x-transformers` (>=2.3.12, pyproject.toml:36) that
the reference touches (x_transformers_rl.py:44-47, 721-734, 505-512).  TEST INFRASTRUCTURE.

PARITY UNPINNED: the package is not in /root/reference nor in this image; behaviour below
is restated from the package's published behaviour.

Surface:  Decoder(dim, heads, depth, attn_dim_head, rotary_pos_emb, attn_dropout,
          ff_dropout, verbose, attn_gate_values, add_value_residual,
          learned_value_residual_mix)
          ContinuousTransformerWrapper(dim_in, dim_out, max_seq_len, probabilistic,
          attn_layers)(x, sum_embeds=, return_embeddings=, return_intermediates=,
          mask= | cache=)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import sqrt

import torch
import torch.nn.functional as F
from torch import nn, cat


@dataclass
class Intermediates:
    cached_kv: tuple | None = None      # un-rotated (k, v), each (b, h, n, dh)
    values: torch.Tensor | None = None  # this layer's pre-mix values for the new tokens


@dataclass
class LayerIntermediates:
    attn_intermediates: list = field(default_factory=list)
    last_hidden: torch.Tensor | None = None
    cache_length: int = 0


class LayerNorm(nn.Module):
    """bias-less LayerNorm: F.layer_norm(eps 1e-5, no affine) * gamma."""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.dim = dim

    def forward(self, x):
        return F.layer_norm(x, (self.dim,)) * self.gamma


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, base=10000):
        super().__init__()
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer('inv_freq', inv_freq)

    def forward(self, pos):
        freqs = torch.einsum('i,j->ij', pos.float(), self.inv_freq)
        # interleaved duplication (t0,t0,t1,t1,...)
        return torch.stack((freqs, freqs), dim=-1).flatten(-2)[None]  # (1, n, dim)


def rotate_half(x):
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def apply_rotary_pos_emb(t, freqs):
    rot_dim, seq_len = freqs.shape[-1], t.shape[-2]
    freqs = freqs[:, -seq_len:, :]
    freqs = freqs[:, None]  # b 1 n d
    t_rot, t_pass = t[..., :rot_dim], t[..., rot_dim:]
    t_rot = t_rot * freqs.cos() + rotate_half(t_rot) * freqs.sin()
    return cat((t_rot, t_pass), dim=-1)


class Attention(nn.Module):
    def __init__(self, dim, heads, dim_head, dropout, gate_values, learned_value_residual_mix):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)
        self.dropout = nn.Dropout(dropout)
        self.to_v_gate = None
        if gate_values:
            self.to_v_gate = nn.Linear(dim, inner)
            nn.init.constant_(self.to_v_gate.weight, 0.)
            nn.init.constant_(self.to_v_gate.bias, 10.)
        self.to_value_residual_mix = None
        if learned_value_residual_mix:
            self.to_value_residual_mix = nn.Sequential(nn.Linear(dim, heads), nn.Sigmoid())

    def forward(self, x, mask=None, rotary_pos_emb=None, cache: Intermediates | None = None,
                value_residual=None):
        b, n, h, dh = *x.shape[:2], self.heads, self.dim_head
        split = lambda t: t.view(b, n, h, dh).transpose(1, 2)
        q, k, v = split(self.to_q(x)), split(self.to_k(x)), split(self.to_v(x))

        orig_values = v
        if value_residual is not None:
            if self.to_value_residual_mix is not None:
                mix = self.to_value_residual_mix(x).transpose(1, 2)[..., None]  # b h n 1
            else:
                mix = 0.5
            v = value_residual.lerp(v, mix) if torch.is_tensor(mix) else value_residual * (1 - mix) + v * mix

        if cache is not None:
            ck, cv = cache.cached_kv
            k, v = cat((ck, k), dim=-2), cat((cv, v), dim=-2)
        inter = Intermediates(cached_kv=(k, v), values=orig_values)

        if rotary_pos_emb is not None:
            q = apply_rotary_pos_emb(q, rotary_pos_emb)
            k = apply_rotary_pos_emb(k, rotary_pos_emb)

        sim = torch.einsum('bhid,bhjd->bhij', q, k) * self.scale
        i, j = sim.shape[-2:]
        mask_value = -torch.finfo(sim.dtype).max
        if mask is not None:
            sim = sim.masked_fill(~mask[:, None, None, :], mask_value)
        causal = torch.ones((i, j), dtype=torch.bool, device=x.device).triu(j - i + 1)
        sim = sim.masked_fill(causal, mask_value)
        attn = sim.softmax(dim=-1, dtype=torch.float32).type(sim.dtype)
        attn = self.dropout(attn)
        out = torch.einsum('bhij,bhjd->bhid', attn, v)
        out = out.transpose(1, 2).reshape(b, n, h * dh)
        if self.to_v_gate is not None:
            out = out * self.to_v_gate(x).sigmoid()
        out = self.to_out(out)
        if mask is not None:
            out = out.masked_fill(~mask[..., None], 0.)
        return out, inter


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, dropout=0.):
        super().__init__()
        self.ff = nn.Sequential(
            nn.Sequential(nn.Linear(dim, dim * mult), nn.GELU()),
            nn.Dropout(dropout),
            nn.Linear(dim * mult, dim),
        )

    def forward(self, x):
        return self.ff(x)


class Decoder(nn.Module):
    """causal pre-norm stack ('a','f') x depth + final norm."""

    def __init__(self, dim, depth, heads=8, attn_dim_head=64, rotary_pos_emb=False,
                 attn_dropout=0., ff_dropout=0., verbose=True, attn_gate_values=False,
                 add_value_residual=False, learned_value_residual_mix=False, **unknown):
        super().__init__()
        if unknown:
            raise TypeError(f'shim Decoder: unsupported kwargs {sorted(unknown)}')
        self.dim, self.depth, self.heads, self.dim_head = dim, depth, heads, attn_dim_head
        self.add_value_residual = add_value_residual
        # rotary_emb_dim defaults to dim_head // 2 (the "< 32" case only warns; verbose=False)
        self.rotary_pos_emb = RotaryEmbedding(attn_dim_head // 2) if rotary_pos_emb else None
        self.layers = nn.ModuleList()
        for ind in range(depth):
            attn = Attention(dim, heads, attn_dim_head, attn_dropout, attn_gate_values,
                             learned_value_residual_mix and add_value_residual and ind > 0)
            ff = FeedForward(dim, dropout=ff_dropout)
            self.layers.append(nn.ModuleList([LayerNorm(dim), attn]))
            self.layers.append(nn.ModuleList([LayerNorm(dim), ff]))
        self.final_norm = LayerNorm(dim)

    def forward(self, x, mask=None, cache: LayerIntermediates | None = None,
                return_hiddens=False, seq_pos_offset=0):
        rotary = None
        if self.rotary_pos_emb is not None:
            # NOTE: positions are arange(x.shape[1] + seq_pos_offset) -- with a
            # one-token input and a cache this is position 0 for every key (identity rotation)
            pos = torch.arange(x.shape[1] + seq_pos_offset, device=x.device)
            rotary = self.rotary_pos_emb(pos)

        prev_len = 0
        attn_cache = []
        if cache is not None:
            assert mask is None
            prev_len = cache.cache_length
            x = x[:, -1:]
            attn_cache = list(cache.attn_intermediates)
        inters = []
        first_values = None
        for ind in range(self.depth):
            norm, attn = self.layers[2 * ind]
            layer_cache = attn_cache[ind] if attn_cache else None
            out, inter = attn(norm(x), mask=mask, rotary_pos_emb=rotary, cache=layer_cache,
                              value_residual=first_values if self.add_value_residual else None)
            if self.add_value_residual and first_values is None:
                first_values = inter.values
            inters.append(inter)
            x = out + x
            norm, ff = self.layers[2 * ind + 1]
            x = ff(norm(x)) + x
        x = self.final_norm(x)
        if not return_hiddens:
            return x
        return x, LayerIntermediates(attn_intermediates=inters, last_hidden=x,
                                     cache_length=prev_len + x.shape[1])


class ContinuousTransformerWrapper(nn.Module):
    def __init__(self, *, max_seq_len, attn_layers, dim_in=None, dim_out=None,
                 probabilistic=False, **unknown):
        super().__init__()
        if unknown:
            raise TypeError(f'shim ContinuousTransformerWrapper: unsupported kwargs {sorted(unknown)}')
        dim = attn_layers.dim
        self.max_seq_len = max_seq_len
        self.attn_layers = attn_layers
        self.project_in = nn.Linear(dim_in, dim, bias=False) if dim_in is not None else nn.Identity()
        assert dim_out is None, 'shim covers dim_out=None only (x_transformers_rl.py:723)'
        self.project_out = nn.Identity()

    def forward(self, x, return_embeddings=False, return_intermediates=False, mask=None,
                sum_embeds=None, cache=None, **unknown):
        if unknown:
            raise TypeError(f'shim wrapper.forward: unsupported kwargs {sorted(unknown)}')
        x = self.project_in(x)
        if sum_embeds is not None:
            x = x + sum_embeds
        x, inter = self.attn_layers(x, mask=mask, cache=cache, return_hiddens=True)
        assert return_embeddings
        if not return_intermediates:
            return x
        return x, inter
