#!/usr/bin/env python
"""bench.py -- rollout-and-learn hot path on B200 vs the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2|c5|c1]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch: the rollout of
every env of the workload for T timesteps (one xtrl_wm_step launch per timestep, inputs resident in
HBM) followed by the GAE kernel over the trajectory buffer.  `value` = env-steps/sec, whole job.
Extra keys: `roofline` (dominant kernel = wm_step_kernel, CUDA-event timed per launch), `cpu_baseline`
(the oracle port timed on this box's host cores), `e2e` (the same rollout through the public
Learner API with a HOST-memory env: per-step pinned H2D of states/rewards, D2H of actions),
`learn` (the second half of BASELINE.json's metric: PPO learn-timesteps/sec of Agent.learn on the trajectories, timed over
>= 3 calls with CUDA events, max over ranks, with its own `roofline` (tensor pipe + HBM as built), `clocks`,
`kernel_share`, `cpu_baseline`; `learn_value` repeats it at the top level), `replicas_identical` (N > 1), `clocks`,
`gpu_launches`.  Both arms print the same `config` keys.

Workloads (BASELINE.json configs):
  c3  4096 vectorised synthetic envs, S8 A4 T500, depth 4 heads 8 (the default; largest 1-GPU config)
  c2  LunarLander-shaped: 192 episodes (64 x 3 genes), depth 4 heads 4 + gates + value residual, genes 32
  c5  4096 envs per GPU, depth 6 heads 8 (32768 envs on 8 GPUs)
  c1  README Sim: S5 A2 T10 depth 1 heads 4, 64 envs
Multi-GPU: episodes shard across ranks with no data-path collective during rollouts (weak scaling:
per-GPU envs fixed); the learn leg sums one flat gradient buffer over the ranks per optimiser step -- inside the optimiser
kernel, over CUDA-IPC peer memory (NVLink / NVSwitch), or by NCCL all_reduce where that cannot be set up (`learn.collective`).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's version banner (printed to stdout at NCCL_DEBUG=VERSION) out of it
if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):       # both levels print "NCCL version ..." on stdout
    del os.environ['NCCL_DEBUG']

METRIC = 'env-steps/sec (rollout) and PPO learn-timesteps/sec'      # BASELINE.json's metric: `value` = the rollout half, `learn_value` = the learn half

WORKLOADS = {
    'c1': dict(S=5, A=2, T=10, envs=64, reward_range=(-1., 1.), world_model=dict(attn_dim_head=16, heads=4, depth=1),
               genes=0, learn_batch=8),
    'c2': dict(S=8, A=4, T=500, envs=192, reward_range=(-5., 5.),
               world_model=dict(attn_dim_head=16, heads=4, depth=4, attn_gate_values=True, add_value_residual=True,
                                learned_value_residual_mix=True), genes=3, gene_dim=32, learn_batch=8),
    'c3': dict(S=8, A=4, T=500, envs=4096, reward_range=(-5., 5.), world_model=dict(attn_dim_head=16, heads=8, depth=4),
               genes=0, learn_batch=256),
    # c4: evolutionary policy optimisation, 64 genes (1 island x 64, 16 selected, tournaments of 4), gene dim 128;
    # 64 episodes x 64 genes = 4096 (episode, gene) pairs per GPU; fitness partials are all-gathered across ranks
    'c4': dict(S=8, A=4, T=500, envs=4096, reward_range=(-5., 5.),
               world_model=dict(attn_dim_head=16, heads=4, depth=4, attn_gate_values=True, add_value_residual=True,
                                learned_value_residual_mix=True), genes=64, gene_dim=128, learn_batch=64,
               pool=dict(num_selected=16, tournament_size=4)),
    'c5': dict(S=8, A=4, T=500, envs=4096, reward_range=(-5., 5.), world_model=dict(attn_dim_head=16, heads=8, depth=6),
               genes=0, learn_batch=256),
}


def algorithmic_bytes_per_env_step(w, t, kv_bytes=4):
    """KV read L*2*I*(t+1) + append L*2*I elements, + 4(S+3) in + 12 out bytes."""
    wm = w['world_model']
    L, I = wm['depth'], wm['heads'] * wm['attn_dim_head']
    return L * 2 * I * (t + 2) * kv_bytes + 4 * (w['S'] + 3) + 12


def learn_flops_per_timestep(w):
    """forward MACs x 2, x 3 for forward + backward (c3: 2 779 776)."""
    wm = w['world_model']
    L, H, dh, D, S, A, K, T = wm['depth'], wm['heads'], wm['attn_dim_head'], 48, w['S'], w['A'], 100, w['T']
    I = H * dh
    gate, vres = int(bool(wm.get('attn_gate_values'))), int(bool(wm.get('learned_value_residual_mix')))
    G = w.get('gene_dim', 0) if w['genes'] > 0 else 0
    hin = (3 if G else 2) * D
    macs = (L * (3 * D * I + I * D + 8 * D * D + gate * D * I + vres * D * H) + L * 2 * I * (T + 1) // 2 + 2 * S * D
            + hin * 2 * D * 2 + 2 * D * A + 2 * D * K + G * D + 2 * D * D + D * 2 * (S + 1) + 2 * D)
    return 3 * 2 * macs


def learn_hbm_bytes_per_timestep(w):
    """HBM bytes per token the learn pass moves AS BUILT: per layer the fp32 residual stream crosses 4
    sub-layer kernels forward and backward, q/k/v/o and their gradients cross as bf16 [b][H][n][16]; then final norm,
    three heads (48-wide fp32 sources in, bf16 logits out and back) and the loss kernel's rows."""
    wm = w['world_model']
    L, I, D, S, A, K = wm['depth'], wm['heads'] * wm['attn_dim_head'], 48, w['S'], w['A'], 100
    res = 4 * D                                                     # one fp32 residual row
    fwd_layer = (res + 3 * I * 2) + (3 * I * 2 + I * 2 + 4 * I // 16) + (res + I * 2 + res) + (res + res)       # pre, core, post, ff
    bwd_layer = (2 * res + res) + (res + I * 2 + I * 2) + (5 * I * 2 + 4 * I // 16 + 3 * I * 2) + (res + 3 * I * 2 + res)   # ff, post, core, pre
    W = A + K + 2 * (S + 1) + 1                                     # head output columns (actor, critic, world model + done)
    heads = (3 * 2 * res + 2 * W) + (3 * 2 * res + 2 * W + 3 * 2 * res)    # fwd: sources in, bf16 logits out; bwd: + source gradients out
    loss = 2 * 2 * W + 4 * (S + 1) + 21
    return L * (fwd_layer + bwd_layer) + 5 * res + heads + loss            # 5 res: final LayerNorm fwd (2) + bwd (3)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit())
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unsampled'])
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in self.rows if len(r) >= 7)]
        top = sm[len(sm) // 2:]                      # the loaded half of the samples
        return dict(sm_mhz=top[len(top) // 2], sm_max_mhz=float(self.rows[0][1]), reasons=reasons, samples=len(sm))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port, batch 1, one timestep per forward
# ------------------------------------------------------------------------------------------------

def cpu_rollout_port(w, episodes, max_seconds=None, threads=None):
    """the reference's rollout loop (x_transformers_rl.py:1220-1355) restated on CPU torch: batch-1
    episodes, growing KV cache, RSNorm -> model -> Discrete -> sample, per timestep.
    Returns (env_steps, seconds, threads)."""
    import torch
    from oracle import restated as R
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import random_state_dict, oracle_config
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = oracle_config(dict(state_dim=w['S'], num_actions=w['A'], reward_range=w['reward_range'], world_model=w['world_model'],
                             evolutionary=w['genes'] > 0, latent_gene_pool=dict(dim=w.get('gene_dim', 0))))
    sd = random_state_dict(cfg, seed=0)
    gene = torch.nn.functional.normalize(torch.randn(1, cfg.dim_gene), dim=-1) if cfg.evolutionary else None
    mean, var = torch.zeros(w['S'] + 1), torch.ones(w['S'] + 1)
    g = torch.Generator().manual_seed(0)
    steps, t0 = 0, time.perf_counter()
    with torch.no_grad():
        for _ in range(episodes):
            cache, prev_a, prev_r = None, torch.full((1, 1), -1), torch.zeros(1, 1)
            for t in range(w['T']):
                state = torch.randn(1, w['S'], generator=g)
                packed = R.rsnorm_eval(torch.cat((state, prev_r), -1), mean, var)
                ra, va, _, _, cache = R.wm_forward(sd, cfg, packed[:, None, :-1], prev_a, packed[:, -1:], 1., gene=gene, cache=cache)
                probs, logp, _ = R.discrete_stats(ra[:, 0])
                a = R.sample_inverse_cdf(probs, torch.rand(1, generator=g))
                _ = logp.gather(-1, a[:, None]); _ = R.hl_decode(va[:, 0], cfg)
                prev_a, prev_r = a[:, None], torch.randn(1, 1, generator=g)
                steps += 1
                if max_seconds is not None and time.perf_counter() - t0 > max_seconds and steps >= w['T']:
                    return steps, time.perf_counter() - t0, threads
    return steps, time.perf_counter() - t0, threads


def cpu_learn_port(w, episodes=8, threads=None, max_seconds=20.):
    """the reference's learn minibatch (x_transformers_rl.py:903-992) restated on CPU torch: teacher-forced forward,
    the four losses, autograd backward -- on `episodes` synthetic episodes of T steps.  Returns timesteps/s."""
    import torch
    from oracle import restated as R
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import random_state_dict, oracle_config
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = oracle_config(dict(state_dim=w['S'], num_actions=w['A'], reward_range=w['reward_range'], world_model=w['world_model'],
                             evolutionary=w['genes'] > 0, latent_gene_pool=dict(dim=w.get('gene_dim', 0))))
    sd = {k: v.requires_grad_() for k, v in random_state_dict(cfg, seed=0).items()}
    g = torch.Generator().manual_seed(0)
    b, T, S, A = episodes, w['T'], w['S'], w['A']
    batch = dict(states=torch.randn(b, T, S, generator=g), actions=torch.randint(0, A, (b, T), generator=g),
                 rewards=torch.randn(b, T, generator=g), old_log_probs=-torch.rand(b, T, generator=g), returns=torch.randn(b, T, generator=g),
                 old_values=torch.randn(b, T, generator=g), dones=torch.zeros(b, T, dtype=torch.bool), gene_ids=torch.zeros(b, dtype=torch.long),
                 episode_lens=torch.full((b,), T))
    genes = torch.randn(max(w['genes'], 1), cfg.dim_gene) if cfg.evolutionary else None
    mean, var = torch.zeros(S + 1), torch.ones(S + 1)
    n, t0 = 0, time.perf_counter()
    while True:
        loss, _, _ = R.learn_minibatch(sd, cfg, batch, mean, var, genes=genes)
        loss.backward()
        n += 1
        if time.perf_counter() - t0 > max_seconds or n >= 8:
            break
    secs = time.perf_counter() - t0
    return n * b * T / secs, n, secs, threads


def run_reference(args, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_rollout_port(dict(w, T=min(w['T'], 50)), 1)
    steps, secs, threads = cpu_rollout_port(w, args.steps)
    v = steps / secs
    sample = f'{args.steps} batch-1 episode(s) x {w["T"]} timesteps, growing KV cache, oracle port (reference deps absent)'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': secs / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.workload, w, args.gpus),
        'cpu_baseline': dict(value=v, unit='env-steps/s', cores=threads, kind='port', sample=sample),
        'e2e': dict(value=v, unit='env-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
    }
    if not args.no_learn:
        lv, n, lsecs, lthreads = cpu_learn_port(w)
        line['learn'] = dict(metric='PPO learn-timesteps/sec', value=lv, unit='timesteps/s', cores=lthreads, kind='port',
                             sample=f'{n} minibatch(es) of 8 episodes x {w["T"]} steps, fwd + losses + backward in {lsecs:.1f}s, oracle/restated.py')
        line['learn_value'] = lv
    print(json.dumps(line))


def dense_name(dense):
    return 'tcgen05 bf16 GEMMs, fp32 accumulate (TMEM)' if dense == 'bf16' else 'fp32 FFMA'


def dtype_name(kv, dense):
    return ('bf16' if dense == 'bf16' else 'f32') + f' dense / {"f32" if kv == "fp32" else "bf16"} KV cache, f32 accumulate'


def workload_config(name, w, n_gpus, **extra):
    wm = w['world_model']
    return dict(workload=f'{name}: {w["envs"]} vectorised synthetic envs per GPU, state_dim={w["S"]} num_actions={w["A"]} '
                         f'T={w["T"]} depth={wm["depth"]} heads={wm["heads"]} dim_head={wm["attn_dim_head"]} hidden=48',
                envs_per_gpu=w['envs'], total_envs=w['envs'] * n_gpus, timesteps=w['T'],
                parallelism=f'episode-sharded dp{n_gpus}, no data-path collective in rollout',
                launch='the T step launches of a pass are replayed as one CUDA graph (device-resident synthetic env)',
                l2='inputs larger than L2 (KV cache GBs per pass); tiny workloads (c1/c2) are L2-resident by nature',
                **extra)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------

def run_b200(args, w):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module('x-transformers-rl_b200')
    envs_mod = importlib.import_module('x-transformers-rl_b200.envs')
    learner_mod = importlib.import_module('x-transformers-rl_b200.learner')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl b200 needs a CUDA device; there is no CPU path (use --impl reference)')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    B, T, S = w['envs'], w['T'], w['S']
    evolutionary = w['genes'] > 0
    def make_learner(kv, dense):
        kw = dict(state_dim=S, num_actions=w['A'], reward_range=w['reward_range'], max_timesteps=T, world_model=w['world_model'],
                  batch_size=w['learn_batch'], num_episodes_per_update=(B // max(w['genes'], 1)) * world, epochs=args.learn_epochs,
                  evolutionary=evolutionary,
                  agent_kwargs=dict(save_path=os.path.join(tempfile.gettempdir(), f'xtrl_bench_{os.getpid()}.pt'),
                                    kv_cache_dtype=kv, rollout_dense_dtype=dense))
        if evolutionary:
            kw['latent_gene_pool'] = dict(dim=w['gene_dim'], num_genes_per_island=w['genes'],
                                          **w.get('pool', dict(num_selected=2, tournament_size=2)))
        kw['accelerate_kwargs'] = dict(mixed_precision='bf16' if args.learn_dtype == 'bf16' else 'no')
        return learner_mod.Learner(**kw)

    learner = make_learner(args.kv, args.dense)
    acc, dev = learner.accelerator, learner.device
    rank = acc.process_index
    assert acc.num_processes == world == args.gpus or world == 1, (acc.num_processes, world, args.gpus)
    torch.manual_seed(1234 + rank)
    env = envs_mod.SyntheticVecEnv(B, S, T, device=dev, seed=rank)
    pairs = learner.episode_genes_for_process
    gids = torch.tensor([p[1] for p in pairs[:B]], dtype=torch.int32, device=dev)
    agent = learner.agent

    def one_pass(e, lrn=None):
        lrn = lrn or learner
        buf, cum = lrn.rollout_vectorized(e, gids, T)
        returns = pkg.ops.gae(buf.rewards, buf.values, buf.is_boundary, agent.gamma, agent.lam, time_major=True)
        return buf, returns

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), out

    for _ in range(args.warmup):
        one_pass(env)
    with ClockSampler(dev.index or 0) as clocks:
        ms, (buf, _) = timed(lambda: one_pass(env), args.steps)
    value = world * B * T * args.steps / (ms / 1e3)
    gpu_launches = args.steps * (T + 1)

    result = {
        'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': dtype_name(args.kv, args.dense), 'data': 'synthetic',
        'config': workload_config(args.workload, w, world),                 # the same keys as the reference arm prints
        'precision': dict(kv_cache=args.kv, rollout_dense=dense_name(args.dense), learn=args.learn_dtype),
        'gpu_launches': gpu_launches, 'clocks': clocks.summary(),
    }

    # ---- roofline of the dominant kernel: CUDA events around every wm_step launch of one more pass ----
    def roofline_pass(lrn, kv):
        st = lrn._stepper(B, T)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
        orig = st.step

        def wrapped(*a, **k):
            t = k['pos_uniform']
            evs[t][0].record(); r = orig(*a, **k); evs[t][1].record()
            return r
        st.step = wrapped
        graphed, lrn.use_cuda_graph = lrn.use_cuda_graph, False        # per-launch events need the eager launch loop
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); one_pass(env, lrn); e1.record()
        torch.cuda.synchronize(dev)
        st.step, lrn.use_cuda_graph = orig, graphed
        kern_ms = [a.elapsed_time(b) for a, b in evs]
        kvb = 4 if kv == 'fp32' else 2
        alg = sum(algorithmic_bytes_per_env_step(w, t, kvb) for t in range(T)) * B / T        # bytes per launch (avg)
        avg_s = sum(kern_ms) / T / 1e3
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
        peak = peaks.get('hbm_gbs', 6650.0)
        achieved = alg / avg_s / 1e9
        traffic = None
        prof = os.path.join(ROOT, 'profiles', 'r02_wm_step_traffic.json')
        if os.path.exists(prof):
            traffic = json.load(open(prof)).get(f'{args.workload}_{kv}')
        return dict(kernel='wm_step_kernel', bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                    traffic=traffic, traffic_source='profiles/r02_wm_step_traffic.json (one ncu --set full capture of the launch at t=250; '
                                                             'a recorded number, not measured in this run)' if traffic else None,
                    peak_source='MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650 GB/s',
                    bytes_per_launch=alg, launch_ms_avg=avg_s * 1e3, launch_ms_first=kern_ms[0], launch_ms_last=kern_ms[-1],
                    kernel_share_of_step=sum(kern_ms) / e0.elapsed_time(e1))

    result['roofline'] = roofline_pass(learner, args.kv)

    # ---- the other precision classes of the same rollout (same workload, same measurement) ----------
    if not args.no_variants and world == 1:
        result['variants'] = {}
        for kv, dense in (('fp32', 'fp32'), ('bf16', 'bf16')):
            if (kv, dense) == (args.kv, args.dense):
                continue
            learner._steppers.clear()                       # one KV cache alive at a time
            torch.cuda.empty_cache()
            lv = make_learner(kv, dense)
            for _ in range(2):
                one_pass(env, lv)
            vms, _ = timed(lambda: one_pass(env, lv), 2)
            rf = roofline_pass(lv, kv)
            result['variants'][f'kv_{kv}+dense_{dense}'] = dict(
                value=B * T * 2 / (vms / 1e3), unit='env-steps/s', dtype=dtype_name(kv, dense), roofline_frac=rf['frac'],
                roofline_achieved_gbs=rf['achieved'], parity_class='1e-5 (fp32, equals the reference)' if dense == 'fp32' and kv == 'fp32' else '1e-2 (bf16)')
            lv._steppers.clear(); del lv
            torch.cuda.empty_cache()

    # ---- learn leg: PPO learn-timesteps/sec of Agent.learn on the trajectories just collected --------
    if not args.no_learn:
        buf.gene_ids.copy_(gids)
        for _ in range(2):
            agent.learn(buf)                               # warm-up (allocator, graph capture)
        agent.event_log = []
        with ClockSampler(dev.index or 0) as lclocks:
            lms, processed = timed(lambda: agent.learn(buf), args.learn_steps)
        torch.cuda.synchronize(dev)
        fb_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in agent.event_log)
        opt_ms = sum(e1.elapsed_time(e2) for _, e1, e2 in agent.event_log)
        n_opt = len(agent.event_log)
        agent.event_log = None
        lval = world * processed * args.learn_steps / (lms / 1e3)
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
        tf_peak, hbm_peak = peaks.get('bf16_tflops_sustained', 1400.0), peaks.get('hbm_gbs', 6650.0)
        fl, by = learn_flops_per_timestep(w), learn_hbm_bytes_per_timestep(w)
        per_gpu = lval / world
        fused = bool(agent.fused_learn)
        result['learn'] = dict(
            metric='PPO learn-timesteps/sec', value=lval, unit='timesteps/s', learn_calls=args.learn_steps, epochs=args.learn_epochs,
            minibatch_episodes=w['learn_batch'], optimizer_steps=n_opt, ms_per_learn_call=lms / args.learn_steps,
            ms_per_minibatch=lms / max(n_opt, 1), dtype=args.learn_dtype, fused_sequence_model=fused,
            clocks=lclocks.summary(),
            kernel_share=dict(fwd_bwd_graph=fb_ms / lms, optimizer=opt_ms / lms,
                              note='CUDA events around every graph replay (prep + sequence model + loss + backward) and every optimiser '
                                   'launch (with the gradient exchange, see collective) inside the timed calls; the rest is GAE, the '
                                   'minibatch gathers and host gaps'),
            collective=('none (one GPU)' if world == 1 else
                        'fused: the optimiser kernel sums the ranks\' gradient buffers over CUDA-IPC peer memory (xtrl_adopt_atan2_step_peers)'
                        if agent._peers is not None else 'torch.distributed.all_reduce (NCCL) before the optimiser launch'),
            roofline=dict(bound='tensor', achieved=per_gpu * fl / 1e12, peak=tf_peak, unit='TFLOP/s', frac=per_gpu * fl / 1e12 / tf_peak,
                          flops_per_timestep=fl, traffic=None,
                          peak_source='MEASURED_PEAKS.json bf16_tflops_sustained (of measured; a kernel timed inside a long step)' if peaks else 'fallback',
                          hbm_as_built=dict(bytes_per_timestep=by, achieved=per_gpu * by / 1e9, peak=hbm_peak, unit='GB/s',
                                            frac=per_gpu * by / 1e9 / hbm_peak,
                                            note='activation bytes the sub-layer kernels move per token as built (DESIGN.md 3.6); '
                                                 'the attention core is instruction-bound, neither pipe')),
            note='bf16 class (default): sequence-model layers, final norm and the head MLPs = libxtrl tcgen05 kernels (LN+QKV+rotary, '
                 'causal attention core with dropout, out-proj+mask+residual, LN+FF+GELU+dropout+residual, heads; fwd and bwd, '
                 'SURVEY 8f-1); GAE, fused loss fwd+bwd, one-launch (gradient exchange +) clip+AdoptAtan2+EMA = libxtrl kernels')
        result['learn_value'] = lval
        if world > 1:                                      # replicas must stay bit-identical (the driver's test box has one GPU)
            st = torch.cat([agent.model.flat_params, agent.ema_params, agent.opt_m, agent.opt_v, agent.rsnorm.running_mean,
                            agent.rsnorm.running_variance])
            h = st.view(torch.int32).to(torch.int64).sum().reshape(1)
            hs = [torch.empty_like(h) for _ in range(world)]
            dist.all_gather(hs, h)
            result['replicas_identical'] = bool(all(int(x) == int(hs[0]) for x in hs)) and bool(torch.isfinite(st).all())

    # ---- gene pool (evolutionary workloads): evolve_ launch + the cross-rank fitness gather ----------------------
    if evolutionary:
        dist_mod = importlib.import_module('x-transformers-rl_b200.distributed')
        pool = agent.gene_pool
        fit = torch.randn(pool.num_genes, device=dev)
        for _ in range(3):
            pool.evolve_(fit)
        ems_, _ = timed(lambda: [pool.evolve_(fit) for _ in range(20)], 1)
        gms_, _ = timed(lambda: [dist_mod.reduce_fitness(fit) for _ in range(20)], 1)
        result['evolve'] = dict(genes=pool.num_genes, dim=pool.dim_gene, ms_per_evolve=ems_ / 20,
                                ms_per_fitness_gather=gms_ / 20, ranks=world,
                                note='evolve_ = 3 torch.randn draws (reference order) + ONE xtrl_evolve launch; fitness = all-gather of '
                                     'per-rank partial sums, added in rank order (bit-identical on every rank)')

    # ---- e2e: the same rollout through the public API with a HOST-memory env -------------------------
    if not args.no_e2e:
        henv = envs_mod.HostVecEnv(B, S, T, seed=rank)
        learner.rollout_vectorized(henv, gids, T)
        ems, _ = timed(lambda: learner.rollout_vectorized(henv, gids, T), 1)
        result['e2e'] = dict(value=world * B * T / (ems / 1e3), unit='env-steps/s',
                             h2d_bytes_per_step=T * B * (S + 1) * 4, d2h_bytes_per_step=T * B * 4,
                             path='Learner.rollout_vectorized(HostVecEnv): pinned H2D of states+rewards and D2H of actions every timestep')

    # ---- cpu baseline beside it (rank 0, N=1 only) ---------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        if 'learn' in result:
            lv, n, lsecs, lthreads = cpu_learn_port(w)
            result['learn']['cpu_baseline'] = dict(value=lv, unit='timesteps/s', cores=lthreads, kind='port',
                                                   sample=f'{n} minibatch(es) of 8 episodes x {T} steps, fwd + losses + backward in {lsecs:.1f}s, oracle/restated.py')
        steps, secs, threads = cpu_rollout_port(w, episodes=10 ** 6, max_seconds=args.cpu_seconds)
        result['cpu_baseline'] = dict(value=steps / secs, unit='env-steps/s', cores=threads, kind='port',
                                      sample=f'{steps} env-steps of batch-1 episodes (T={T}) in {secs:.1f}s, oracle/restated.py')
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=('b200', 'reference'))
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--kv', default='fp32', choices=('fp32', 'bf16'))
    ap.add_argument('--dense', default='bf16', choices=('fp32', 'bf16'),
                    help='rollout QKV/out/FFN/head GEMMs: bf16 = tcgen05 tensor cores (1e-2 class), fp32 = FFMA (1e-5 class)')
    ap.add_argument('--no-variants', action='store_true')
    ap.add_argument('--learn-dtype', default='bf16', choices=('fp32', 'bf16'))
    ap.add_argument('--learn-epochs', type=int, default=4)
    ap.add_argument('--learn-steps', type=int, default=3, help='timed Agent.learn calls (after 2 warm-up calls)')
    ap.add_argument('--cpu-seconds', type=float, default=15.)
    ap.add_argument('--no-learn', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == '__main__':
    main()
