import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import restated as R
from helpers import random_state_dict, rel_err
import test_gpu_kernels as T
for kw in (dict(heads=8, depth=4, state_dim=8, num_actions=4), dict(heads=8, depth=6, state_dim=8, num_actions=4)):
  for seed in (1,2,3):
    cfg = R.Config(reward_range=(-5., 5.), **kw)
    sd = random_state_dict(cfg, seed=seed)
    B, Tn = 37, 40
    g = torch.Generator().manual_seed(2)
    states, prev_r = torch.randn(B, Tn, 8, generator=g), torch.randn(B, Tn, generator=g)
    prev_a = torch.randint(-1, 4, (B, Tn), generator=g)
    want_ra, want_va = T._rollout_oracle(sd, cfg, states, prev_a, prev_r, None)
    for name, extra in (('tc', dict(dense_bf16=True)), ('tc+kv', dict(dense_bf16=True, kv_bf16=True)), ('kv only', dict(kv_bf16=True))):
        ra, va, val = T._rollout_kernel(sd, cfg, states, prev_a, prev_r, **extra)
        wv = R.hl_decode(want_va, cfg)
        # also: median relative error per element for value
        print(kw['depth'], seed, name, 'ra %.4f va %.4f val(max-norm) %.4f  val abs-max-diff %.4f of range 10' % (rel_err(ra, want_ra), rel_err(va, want_va), rel_err(val, wv), (val-wv).abs().max()))
