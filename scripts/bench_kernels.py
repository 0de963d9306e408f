"""Roofline microbench of the non-step kernels (GAE, fused PPO loss, evolve, optimiser) through the C ABI.
Each kernel: L2 flushed between iterations (256 MB write), CUDA-event timing per launch, algorithmic
bytes from SURVEY 8d / DESIGN.md.  One JSON line per kernel.   python scripts/bench_kernels.py [--iters 10]"""
import argparse, importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module('x-transformers-rl_b200').ops
ap = argparse.ArgumentParser(); ap.add_argument('--iters', type=int, default=10); ap.add_argument('--only', default='')
a = ap.parse_args()
dev = 'cuda'
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.
flush_buf = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
_flush_sink = torch.zeros(1, dtype=torch.int64, device=dev)


_FLUSH_WRITE = os.environ.get('XTRL_FLUSH', '') == 'write'


def flush_l2():
    """replace the contents of L2 by reading 256 MB.  A write (fill_) would leave L2 full of dirty lines, and the timed kernel
    would then pay for their write-back on top of its own traffic (up to 126 MB, ~20 us at full HBM speed -- more than the
    small kernels here take).  XTRL_FLUSH=write restores the old behaviour for comparison."""
    if _FLUSH_WRITE:
        flush_l2()
    else:
        torch.sum(flush_buf.view(torch.int64), dim=0, keepdim=True, out=_flush_sink)



def timeit(fn, iters=a.iters):
    """median device time of fn's launches, L2 flushed before each.  The launches are replayed from a CUDA graph, so the
    Python / ctypes time of the wrapper (tens of microseconds, more than some of these kernels) hides behind the flush
    instead of being counted between the two events."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    ms = []
    for _ in range(iters):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def timeit_rot(fns, iters=a.iters):
    """per-launch device time of small kernels: len(fns) launches on DIFFERENT buffers captured back to back in one graph, so
    the ~10 us a graph launch costs between the two events is shared by all of them; the buffer sets together exceed L2 (and
    L2 is flushed before every replay), so every launch still reads its inputs from DRAM."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for f in fns:
            f()
    ms = []
    for _ in range(iters):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / len(fns))
    ms.sort()
    return ms[len(ms) // 2]


def report(name, alg_bytes, ms, **extra):
    gbs = alg_bytes / (ms / 1e3) / 1e9
    print(json.dumps(dict(kernel=name, ms=ms, algorithmic_bytes=alg_bytes, achieved_gbs=gbs, peak_gbs=peak, frac=gbs / peak, **extra)))


E, T, A, K, S = 4096, 500, 4, 100, 8
g = torch.Generator(device=dev).manual_seed(0)
if not a.only or 'gae' in a.only:
    r, v = torch.randn(E, T, device=dev, generator=g), torch.randn(E, T, device=dev, generator=g)
    bd = torch.rand(E, T, device=dev, generator=g) < 0.02
    out = torch.empty_like(r)
    report('gae_episode_major (E=4096,T=500)', 13 * E * T, timeit(lambda: ops.gae(r, v, bd, out=out)), unit_bytes=13)
    rt, vt, bt = r.T.contiguous(), v.T.contiguous(), bd.T.contiguous()
    outt = torch.empty_like(rt)
    report('gae_time_major (T=500,B=4096)', 13 * E * T, timeit(lambda: ops.gae(rt, vt, bt, time_major=True, out=outt)), unit_bytes=13)
    sets = [(torch.randn(T, E, device=dev, generator=g), torch.randn(T, E, device=dev, generator=g), torch.rand(T, E, device=dev, generator=g) < 0.02,
             torch.empty(T, E, device=dev)) for _ in range(8)]
    report('gae_time_major (T=500,B=4096), 8 launches on 8 buffer sets per graph', 13 * E * T,
           timeit_rot([(lambda s_=s_: ops.gae(s_[0], s_[1], s_[2], time_major=True, out=s_[3])) for s_ in sets]), unit_bytes=13,
           note='the single-launch line above includes the fixed cost of launching a one-node graph')
    del sets
    E8 = 8 * E                                              # c5 total: 32768 episodes (213 MB)
    r8, v8 = torch.randn(T, E8, device=dev, generator=g), torch.randn(T, E8, device=dev, generator=g)
    b8 = torch.rand(T, E8, device=dev, generator=g) < 0.02
    o8 = torch.empty_like(r8)
    report('gae_time_major (T=500,B=32768)', 13 * E8 * T, timeit(lambda: ops.gae(r8, v8, b8, time_major=True, out=o8)), unit_bytes=13)
    r8e, v8e, b8e = r8.T.contiguous(), v8.T.contiguous(), b8.T.contiguous()
    report('gae_episode_major (E=32768,T=500)', 13 * E8 * T, timeit(lambda: ops.gae(r8e, v8e, b8e, out=o8.view(E8, T))), unit_bytes=13)
    del r8, v8, b8, o8, r8e, v8e, b8e
if not a.only or 'loss' in a.only:
    for dt, nm in ((torch.float32, 'fp32'), (torch.bfloat16, 'bf16')):
        b = 1024                                            # 512k timesteps: 0.5 GB of logits+grads in fp32
        ra, vl = torch.randn(b, T, A, device=dev, generator=g).to(dt), torch.randn(b, T, K, device=dev, generator=g).to(dt)
        sp, dl = torch.randn(b, T, 2 * (S + 1), device=dev, generator=g).to(dt), torch.randn(b, T, device=dev, generator=g).to(dt)
        actions = torch.randint(0, A, (b, T), device=dev, generator=g).int()
        olp, ret, ov = -torch.rand(b, T, device=dev, generator=g), torch.randn(b, T, device=dev, generator=g), torch.randn(b, T, device=dev, generator=g)
        tgt, dn = torch.randn(b, T, S + 1, device=dev, generator=g), torch.rand(b, T, device=dev, generator=g) < 0.05
        lens = torch.full((b,), T, dtype=torch.int32, device=dev)
        cfg = ops.LossConfig(num_actions=A, num_bins=K, state_dim=S, reward_range=(-5., 5.))
        sz = 4 if dt == torch.float32 else 2
        per = (2 * A + 2 * K) * sz + 21 + (4 * (S + 1) + 2) * sz + 4 * (S + 1) + 1
        report(f'ppo_loss_fwd_bwd {nm} (b=1024,T=500,A=4,K=100,S=8)', per * b * T,
               timeit(lambda: ops.ppo_loss_fwd_bwd(cfg, ra, vl, sp, dl, actions, olp, ret, ov, tgt, dn, lens)), unit_bytes=per,
               note='includes memset + stats + finalize launches and the output allocations of the wrapper')
if not a.only or 'evolve' in a.only:
    for (I, gg, s, k, G, nm) in ((4096, 16, 4, 2, 128, 'scaled 4096x16x128'), (1, 64, 16, 4, 128, 'c4 1x64x128'), (8, 8, 4, 2, 128, 'c4 8x8x128')):
        c = gg - s
        genes = torch.nn.functional.normalize(torch.randn(I * gg, G, device=dev, generator=g), dim=-1)
        fit = torch.randn(I * gg, device=dev, generator=g)
        nt, nc, nm_ = torch.randn(I, c, s, device=dev, generator=g), torch.randn(I, c, G, device=dev, generator=g), torch.randn(I, gg - 1, G, device=dev, generator=g)
        alg = 2 * I * gg * G * 4 + 4 * I * gg + 4 * I * (c * s + c * G + (gg - 1) * G)
        report(f'evolve ({nm})', alg, timeit(lambda: ops.evolve_(genes, fit, nt, nc, nm_, num_islands=I, num_genes_per_island=gg,
                                                                  num_selected=s, tournament_size=k)))
if not a.only or 'opt' in a.only:
    for n, nm in ((208827, 'c3 model, 208,827 params'), (64 << 20, 'synthetic 64Mi params')):
        nt = 50
        offs = torch.linspace(0, n, nt + 1, device=dev).long(); offs[-1] = n
        P, G_, M, V, PI, EM = (torch.randn(n, device=dev, generator=g) * 0.1 for _ in range(6))
        V.abs_()
        ws = None
        report(f'adopt_atan2_step+clip+ema ({nm})', 40 * n,
               timeit(lambda: ops.adopt_atan2_step(P, G_, M, V, PI, EM, offs, lr=8e-4, init_lr=8e-4, betas=(0.9, 0.99), regen_rate=1e-4,
                                                   cautious_factor=0.1, max_grad_norm=0.5, first_step=False, ema_mode=2, ema_decay=0.9)),
               unit_bytes=40, note='one cooperative launch, two grid-wide barriers; the gradient is re-read by each phase (56 B/elem as built beyond L2)')
