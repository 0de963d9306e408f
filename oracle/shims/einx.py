"""Stand-in for the four einx patterns the reference uses (x_transformers_rl.py:24-25,194,
437-438,499,909; distributed.py:93).  TEST INFRASTRUCTURE.  SURVEY.md Appendix A.6."""
import torch


def _key(pattern):
    return ''.join(pattern.split())


def multiply(pattern, a, b):
    k = _key(pattern)
    if k == 'bn...,bn->bn...':
        return a * b.reshape(*b.shape, *((1,) * (a.ndim - b.ndim)))
    if k == '...,d->...d':
        a = torch.as_tensor(a)
        return a[..., None] * b
    raise NotImplementedError(f'einx shim: multiply({pattern!r})')


def less(pattern, a, b):
    k = _key(pattern)
    if k == 'n,b->bn':
        return a[None, :] < b[:, None]
    if k == 'ji->(ij)':
        return (a[None, :] < b[:, None]).reshape(-1)
    raise NotImplementedError(f'einx shim: less({pattern!r})')


def where(pattern, cond, x, y):
    k = _key(pattern)
    if k == 'bn,bnd,':
        return torch.where(cond[..., None], x, torch.as_tensor(y, dtype=x.dtype, device=x.device))
    raise NotImplementedError(f'einx shim: where({pattern!r})')
