"""where does the host-env rollout spend its time?  (wait on the GPU vs host work per timestep)"""
import importlib, os, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module('x-transformers-rl_b200.learner'); E = importlib.import_module('x-transformers-rl_b200.envs')
B, T = 4096, 500
lrn = L.Learner(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=T, world_model=dict(attn_dim_head=16, heads=8, depth=4),
                batch_size=256, num_episodes_per_update=B, agent_kwargs=dict(save_path=os.path.join(tempfile.gettempdir(), 'p.pt'), rollout_dense_dtype='bf16'))
gids = torch.zeros(B, dtype=torch.int32, device=lrn.device)
henv = E.HostVecEnv(B, 8, T, seed=0)
import types
orig_sync = torch.cuda.Event.synchronize
wait = [0.]
def timed_sync(self):
    t0 = time.perf_counter(); orig_sync(self); wait[0] += time.perf_counter() - t0
torch.cuda.Event.synchronize = timed_sync
for g in (1, 2):
    lrn.rollout_groups = g
    lrn.rollout_vectorized(henv, gids, T)
    torch.cuda.synchronize(); wait[0] = 0.; t0 = time.perf_counter()
    lrn.rollout_vectorized(henv, gids, T)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print(f'groups {g}: total {tot*1e3:.0f} ms, waiting on events {wait[0]*1e3:.0f} ms, host work {(tot-wait[0])*1e3:.0f} ms  ({(tot-wait[0])/T*1e6:.0f} us per timestep)')
denv = E.SyntheticVecEnv(B, 8, T, device=lrn.device)
lrn.rollout_vectorized(denv, gids, T); lrn.rollout_vectorized(denv, gids, T)
torch.cuda.synchronize(); t0 = time.perf_counter(); lrn.rollout_vectorized(denv, gids, T); torch.cuda.synchronize()
print(f'device env (graph): {1e3*(time.perf_counter()-t0):.0f} ms')
import cProfile, pstats
lrn.rollout_groups = 1
torch.cuda.Event.synchronize = orig_sync
pr = cProfile.Profile(); pr.enable(); lrn.rollout_vectorized(henv, gids, T); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
