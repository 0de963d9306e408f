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
