"""probe: G independent env groups on G CUDA streams (phase-desynchronised step kernels)."""
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
ap.add_argument('--bf16', action='store_true'); ap.add_argument('--groups', type=int, default=2)
ap.add_argument('--skew', type=float, default=0.5)
a = ap.parse_args()
dev = 'cuda'
cfg = R.Config(state_dim=8, num_actions=4, heads=a.heads, depth=a.depth, reward_range=(-5., 5.))
sd = {k: v.to(dev) for k, v in random_state_dict(cfg, seed=0).items()}
G, T = a.groups, a.T
B = a.envs // G
sts, data, streams = [], [], []
for g in range(G):
    st = ops.WorldModelStepper(model_config(cfg, T, kv_bf16=a.bf16), B, dev); st.load_state_dict(sd)
    sts.append(st); streams.append(torch.cuda.Stream())
    acts = torch.full((T + 1, B), -1, dtype=torch.int32, device=dev)
    prev_r = torch.zeros(T + 1, B, device=dev); prev_r[1:] = torch.randn(T, B, device=dev)
    data.append(dict(states=torch.randn(T, B, 8, device=dev), acts=acts, prev_r=prev_r, u=torch.rand(T, B, device=dev),
                     logp=torch.empty(T, B, device=dev), val=torch.empty(T, B, device=dev)))
mean, var = torch.zeros(9, device=dev), torch.ones(9, device=dev)
spin = torch.empty(1 << 22, device=dev)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    for g in range(1, G):            # initial skew: delay group g by a few fill kernels
        with torch.cuda.stream(streams[g]):
            for _ in range(int(a.skew * g * 8)):
                spin.fill_(1.0)
    for t in range(T):
        for g in range(G):
            d = data[g]
            with torch.cuda.stream(streams[g]):
                sts[g].step(d['states'][t], d['acts'][t], d['prev_r'][t], mean, var, uniform=d['u'][t],
                            out=dict(action=d['acts'][t + 1], log_prob=d['logp'][t], value=d['val'][t]), pos_uniform=t)
    torch.cuda.synchronize(); wall = time.time() - t0
    sz = 2 if a.bf16 else 4
    bytes_total = a.envs * sum(a.depth * 2 * a.heads * 16 * (t + 2) * sz + 56 for t in range(T))
    print(f'G={G} rep {rep}: wall {wall:.3f}s env-steps/s {a.envs*T/wall/1e6:.2f}M achieved {bytes_total/wall/1e9:.0f} GB/s')
