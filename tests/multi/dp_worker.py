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
                    agent_kwargs=dict(save_path=os.path.join(tempfile.gettempdir(), f'dp_{os.getpid()}.pt'),
                                      peer_allreduce=os.environ.get('XTRL_DP_PEERS', '1') == '1'),
                    # bf16: the learn pass runs the fused tcgen05 sub-layers (weight gradients by atomicAdd -- replicas
                    # must still agree bit for bit, because every rank applies the same all-reduced gradient)
                    accelerate_kwargs=dict(mixed_precision='bf16' if os.environ.get('XTRL_DP_MIXED') == 'bf16' else 'no'))
learner.agent.capture_allreduce = os.environ.get('XTRL_DP_CAPTURE') == '1'
acc = learner.accelerator
rank, world = acc.process_index, acc.num_processes
assert world >= 2 and dist.get_backend() == 'nccl'


def fused_exchange_check():
    """xtrl_adopt_atan2_step_peers (gradient sum over the ranks' mapped buffers inside the optimiser kernel) against
    NCCL all_reduce + the plain optimiser launch on the same inputs; three steps (first step, then two with moments)."""
    ops = importlib.import_module('x-transformers-rl_b200.ops')
    D = importlib.import_module('x-transformers-rl_b200.distributed')
    dev, n, tail = learner.device, 10007, 5                  # odd size: the last 16-byte unit is partial
    pg = D.PeerGroup.create(n + tail, dev)
    assert pg is not None, 'peer-memory exchange could not be set up on this box'
    offs = torch.tensor([0, 1000, 1003, 5000, n], device=dev)
    common = torch.Generator(device=dev).manual_seed(5)      # the same state on every rank
    st = [[(torch.randn(n, device=dev, generator=common) * 0.1) for _ in range(5)] for _ in range(1)][0]
    A, B = [t.clone() for t in st], [t.clone() for t in st]
    mine = torch.Generator(device=dev).manual_seed(50 + rank)
    worst = 0.
    for step in range(3):
        g = torch.randn(n + tail, device=dev, generator=mine)
        pg.local.copy_(g)
        summed = g.clone()
        dist.all_reduce(summed)
        kw = dict(lr=8e-4, init_lr=8e-4, betas=(0.9, 0.99), regen_rate=1e-4, cautious_factor=0.1, max_grad_norm=0.5,
                  first_step=(step == 0), ema_mode=2, ema_decay=0.9, grad_scale=1. / world)
        ops.adopt_atan2_step(A[0], pg.local[:n], A[1], A[2], A[3], A[4], offs, peers=pg, **kw)
        ops.adopt_atan2_step(B[0], summed[:n], B[1], B[2], B[3], B[4], offs, **kw)
        torch.cuda.synchronize()
        pg.check()
        if world == 2:                                       # a two-term sum does not depend on the order
            assert torch.equal(pg.reduced, summed), 'sum over ranks differs from NCCL'
            assert all(torch.equal(x, y) for x, y in zip(A, B)), f'optimiser state differs at step {step}'
        else:
            worst = max(worst, float((pg.reduced - summed).abs().max()), max(float((x - y).abs().max()) for x, y in zip(A, B)))
            assert worst < 1e-5, worst
    allp = [torch.empty_like(A[0]) for _ in range(world)]
    dist.all_gather(allp, A[0])
    assert all(torch.equal(allp[0], q) for q in allp[1:]), 'ranks disagree after the fused exchange'
    return worst


rank, world = learner.accelerator.process_index, learner.accelerator.num_processes
fused_worst = fused_exchange_check() if os.environ.get('XTRL_DP_PEERS', '1') == '1' else None
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
    print(f'DP_RESULT identical={ok} changed={changed} finite={finite} world={world} steps={learner.timings["env_steps"]} '
          f'peers={ag._peers is not None} fused_exchange_worst={fused_worst}')
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok and changed and finite) else 1)
