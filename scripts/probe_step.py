"""quick perf probe of xtrl_wm_step at a BASELINE shape (not the bench; used while tuning)."""
import argparse, importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import restated as R
from helpers import random_state_dict, model_config
ops = importlib.import_module('x-transformers-rl_b200').ops

ap = argparse.ArgumentParser()
ap.add_argument('--envs', type=int, default=4096); ap.add_argument('--T', type=int, default=500)
ap.add_argument('--depth', type=int, default=4); ap.add_argument('--heads', type=int, default=8)
ap.add_argument('--bf16', action='store_true'); ap.add_argument('--reps', type=int, default=2); ap.add_argument('--tc', action='store_true')
a = ap.parse_args()
dev = 'cuda'
cfg = R.Config(state_dim=8, num_actions=4, heads=a.heads, depth=a.depth, reward_range=(-5., 5.))
sd = random_state_dict(cfg, seed=0)
B, T = a.envs, a.T
st = ops.WorldModelStepper(model_config(cfg, T, kv_bf16=a.bf16, dense_bf16=a.tc), B, dev)
st.load_state_dict({k: v.to(dev) for k, v in sd.items()})
states = torch.randn(T, B, 8, device=dev); rewards = torch.randn(T, B, device=dev); u = torch.rand(T, B, device=dev)
mean, var = torch.zeros(9, device=dev), torch.ones(9, device=dev)
acts = torch.full((T + 1, B), -1, dtype=torch.int32, device=dev)
prev_r = torch.zeros(T + 1, B, device=dev); prev_r[1:] = rewards
logp = torch.empty(T, B, device=dev); val = torch.empty(T, B, device=dev)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
for rep in range(a.reps):
    st.reset(); torch.cuda.synchronize(); t0 = time.time()
    for t in range(T):
        ev[t][0].record()
        st.step(states[t], acts[t], prev_r[t], mean, var, uniform=u[t], out=dict(action=acts[t + 1], log_prob=logp[t], value=val[t]))
        ev[t][1].record()
    torch.cuda.synchronize(); wall = time.time() - t0
    ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    kern = sum(ms) / 1e3
    sz = 2 if a.bf16 else 4
    I, L = a.heads * 16, a.depth
    bytes_total = B * sum(L * 2 * I * (t + 2) * sz + 4 * 11 + 12 for t in range(T))
    print(f'rep {rep}: wall {wall:.3f}s kernels {kern:.3f}s  env-steps/s wall {B*T/wall/1e6:.2f}M kern {B*T/kern/1e6:.2f}M  '
          f'achieved {bytes_total/kern/1e9:.0f} GB/s  step ms t=0 {ms[0]:.3f} t=100 {ms[100%T]:.3f} t=250 {ms[250%T]:.3f} t={T-1} {ms[T-1]:.3f}')
