"""torchrun worker for tests/test_gpu_multi.py: 2+ ranks, NCCL.  Each rank rolls out ITS shard of the
(episode, gene) pairs with a differently seeded env, learns with the flat-gradient all-reduce, and the ranks
then prove their replicas stayed bit-identical (params, EMA, optimiser state, RSNorm, gene pool)."""
import importlib
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
L = importlib.import_module('x-transformers-rl_b200.learner')
E = importlib.import_module('x-transformers-rl_b200.envs')

torch.manual_seed(1000 + int(os.environ.get('RANK', '0')))   # DIFFERENT local init per rank: the Agent broadcasts rank 0's weights / genes
learner = L.Learner(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=16, batch_size=8,
                    num_episodes_per_update=16, epochs=2, evolutionary=True, evolve_every=1, evolve_after_step=0,
                    latent_gene_pool=dict(dim=32, num_genes_per_island=4, num_selected=2, tournament_size=2),
                    world_model=dict(attn_dim_head=16, heads=4, depth=2, attn_gate_values=True),
                    agent_kwargs=dict(save_path=os.path.join(tempfile.gettempdir(), f'dp_{os.getpid()}.pt')),
                    # bf16: the learn pass runs the fused tcgen05 sub-layers (weight gradients by atomicAdd -- replicas
                    # must still agree bit for bit, because every rank applies the same all-reduced gradient)
                    accelerate_kwargs=dict(mixed_precision='bf16' if os.environ.get('XTRL_DP_MIXED') == 'bf16' else 'no'))
learner.agent.capture_allreduce = os.environ.get('XTRL_DP_CAPTURE') == '1'
acc = learner.accelerator
rank, world = acc.process_index, acc.num_processes
assert world >= 2 and dist.get_backend() == 'nccl'
pairs = learner.episode_genes_for_process
assert len(pairs) == 16 * 4 // world
env = E.SyntheticVecEnv(len(pairs), 8, 16, device=learner.device, seed=100 + rank)      # different data per rank
init = learner.agent.model.flat_params.clone()
learner(env, 3, seed=0)
ag = learner.agent
state = torch.cat([ag.model.flat_params, ag.ema_params, ag.opt_m, ag.opt_v, ag.rsnorm.running_mean,
                   ag.rsnorm.running_variance, ag.gene_pool.genes.flatten()])
gathered = [torch.empty_like(state) for _ in range(world)]
dist.all_gather(gathered, state)
ok = all(torch.equal(gathered[0], g) for g in gathered[1:])
changed = not torch.equal(init, ag.model.flat_params)
finite = bool(torch.isfinite(state).all())
if rank == 0:
    print(f'DP_RESULT identical={ok} changed={changed} finite={finite} world={world} steps={learner.timings["env_steps"]}')
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok and changed and finite) else 1)
