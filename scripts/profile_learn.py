"""torch-profiler view of one Agent.learn minibatch loop at the c3 shape (which ops dominate)."""
import importlib, os, sys, tempfile, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module('x-transformers-rl_b200.learner'); E = importlib.import_module('x-transformers-rl_b200.envs')
B, T = 1024, 500
dt = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
learner = L.Learner(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=T, world_model=dict(attn_dim_head=16, heads=8, depth=4),
                    batch_size=256, num_episodes_per_update=B, epochs=1, agent_kwargs=dict(save_path=os.path.join(tempfile.gettempdir(), 'p.pt')),
                    accelerate_kwargs=dict(mixed_precision='bf16' if dt == 'bf16' else 'no'))
env = E.SyntheticVecEnv(B, 8, T, device=learner.device)
gids = torch.zeros(B, dtype=torch.int32, device=learner.device)
buf, _ = learner.rollout_vectorized(env, gids, T)
learner.agent.learn(buf)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    learner.agent.learn(buf); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=28, max_name_column_width=60))
