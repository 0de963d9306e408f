"""CUDA-event timings of the learn-side fused sub-layer kernels at the c3 minibatch shape (256 episodes x 500 steps,
8 heads): one JSON line per kernel with the average launch time and its algorithmic HBM bytes.
    python scripts/bench_learn_kernels.py [--b 256] [--n 500] [--heads 8] [--dropout 0.25] [--only core]"""
import argparse, importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module('x-transformers-rl_b200.ops')

ap = argparse.ArgumentParser()
ap.add_argument('--b', type=int, default=256); ap.add_argument('--n', type=int, default=500); ap.add_argument('--heads', type=int, default=8)
ap.add_argument('--dropout', type=float, default=0.25); ap.add_argument('--iters', type=int, default=20); ap.add_argument('--only', default='')
args = ap.parse_args()
dev, b, n, H, D, p = 'cuda', args.b, args.n, args.heads, 48, args.dropout
I, N = 16 * H, args.b * args.n
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s, sc=1.: torch.randn(*s, generator=g, device=dev) * sc
bf = lambda *s: r(*s).bfloat16()
x, dy = r(b, n, D), r(b, n, D, sc=1e-3)
gamma, w1, b1, w2, b2 = 1 + r(D, sc=.1), r(4 * D, D, sc=D ** -.5), r(4 * D, sc=.1), r(D, 4 * D, sc=(4 * D) ** -.5), r(D, sc=.1)
wq, wk, wv, wo = r(I, D, sc=D ** -.5), r(I, D, sc=D ** -.5), r(I, D, sc=D ** -.5), r(D, I, sc=I ** -.5)
ff_img, at_img = ops.ff_pack(w1, w2), ops.attn_pack(wq, wk, wv, wo)
inv_freq = 1. / (10000 ** (torch.arange(0, 8, 2, device=dev).float() / 8))
ang = torch.arange(n, device=dev, dtype=torch.float32)[:, None] * inv_freq[None]
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
lens = torch.full((b,), n, dtype=torch.int32, device=dev)
step = torch.zeros((), dtype=torch.int64, device=dev)
q, k, v, d_o = bf(b, H, n, 16), bf(b, H, n, 16), bf(b, H, n, 16), bf(b, H, n, 16)
o, lse = ops.attn_core_fwd(q, k, v, dropout=p, rng_step=step, rng_key=1)
z = lambda t: torch.zeros_like(t)
gr = dict(gamma=z(gamma), w1=z(w1), b1=z(b1), w2=z(w2), b2=z(b2), wq=z(wq), wk=z(wk), wv=z(wv), wo=z(wo))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# critic head: [embed | state_embed] (96) -> 96 -> 100 bins
w1p, b1p, w2p, b2p = r(96, 96, sc=96 ** -.5), r(96, sc=.1), r(112, 96, sc=96 ** -.5), r(112, sc=.1)
w2p[100:] = 0; b2p[100:] = 0
mlp_img = ops.mlp_pack(w1p, w2p)
src = [r(N, 48), r(N, 48)]
gy = (r(N, 100) * 1e-3).bfloat16()
gm = [z(w1p), z(b1p), z(w2p), z(b2p)]
xf, dyf, dgam = x.view(N, D), dy.view(N, D), z(gamma)

cases = {
    'ff_fwd': (lambda: ops.ff_fwd(x, gamma, b1, b2, ff_img, dropout=p, rng_step=step, rng_key=0), N * D * 4 * 2),
    'ff_bwd': (lambda: ops.ff_bwd(x, dy, gamma, b1, ff_img, gr['gamma'], gr['w1'], gr['b1'], gr['w2'], gr['b2'], dropout=p, rng_step=step, rng_key=0), N * D * 4 * 3),
    'attn_pre_fwd': (lambda: ops.attn_pre_fwd(x, gamma, at_img, cos, sin, H), N * D * 4 + N * 3 * I * 2),
    'attn_pre_bwd': (lambda: ops.attn_pre_bwd(x, q, k, v, gamma, at_img, cos, sin, gr['gamma'], gr['wq'], gr['wk'], gr['wv']), N * D * 4 * 2 + N * 3 * I * 2),
    'attn_post_fwd': (lambda: ops.attn_post_fwd(x, o, lens, at_img), N * D * 4 * 2 + N * I * 2),
    'attn_post_bwd': (lambda: ops.attn_post_bwd(dy, o, lens, at_img, gr['wo']), N * D * 4 + N * I * 2 * 2),
    'mlp_critic_fwd': (lambda: ops.mlp_fwd(src, n, b1p, b2p, mlp_img, 100, 96), N * 96 * 4 + N * 100 * 2),
    'mlp_critic_bwd': (lambda: ops.mlp_bwd(src, n, gy, b1p, mlp_img, *gm, 96), N * 96 * 4 * 2 + N * 100 * 2),
    'ln_fwd': (lambda: ops.ln_fwd(xf, gamma), N * D * 4 * 2),
    'ln_bwd': (lambda: ops.ln_bwd(xf, dyf, gamma, dgam), N * D * 4 * 3),
    'attn_core_fwd': (lambda: ops.attn_core_fwd(q, k, v, dropout=p, rng_step=step, rng_key=1), N * I * 2 * 4 + N * H * 4),
    'attn_core_bwd': (lambda: ops.attn_core_bwd(q, k, v, o, lse, d_o, dropout=p, rng_step=step, rng_key=1), N * I * 2 * 8 + N * H * 4),
}
for name, (fn, nbytes) in cases.items():
    if args.only and args.only not in name:
        continue
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.iters):
        flush.zero_()                                   # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sum(ts) / len(ts)
    scores = b * H * n * (n + 1) // 2 if 'core' in name else None
    print(json.dumps(dict(kernel=name, us=round(us, 1), tokens=N, algorithmic_bytes=nbytes, gbps=round(nbytes / us / 1e3, 1),
                          scores_per_s=None if scores is None else round(scores / us * 1e6 / 1e9, 2), dropout=p)))
