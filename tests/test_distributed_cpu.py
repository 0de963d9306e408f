"""CPU, world_size 2, gloo: the host-side multi-process logic of the data-parallel path -- episode/gene
sharding, seed sync, deterministic fitness reduction, the flat gradient (+ RSNorm mean) all-reduce and the
Accelerator facade.  No kernel runs here (those need a GPU); this is the plumbing around them."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    D = importlib.import_module('x-transformers-rl_b200.distributed')
    acc = D.Accelerator(cpu=True)                      # initialises gloo from the env, like torchrun does
    assert acc.num_processes == world and acc.process_index == rank and acc.is_main_process == (rank == 0)
    assert D.is_distributed()
    res = {}
    # (1) sharding: cartesian_prod(episodes, genes).chunk(world)[rank]  (reference :1146-1154)
    pairs = D.shard_episode_genes(4, 3, world, rank)
    res['pairs'] = pairs
    # (2) seed sync: every rank ends with the same integer
    torch.manual_seed(100 + rank)
    res['seed'] = D.sync_seed('cpu')
    # (3) fitness: per-rank partial sums -> identical totals on every rank, rank-order summation
    partial = torch.zeros(3)
    for ep, gene in pairs:
        partial[gene] += float(ep * 10 + gene + 0.1 * rank)
    res['fitness'] = D.reduce_fitness(partial).tolist()
    res['acc_reduce'] = acc.reduce(partial.clone()).tolist()
    # (4) flat gradient buffer with the RSNorm batch mean in its tail: one mean all-reduce
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    D.all_reduce_mean_(flat)
    res['flat'] = flat.tolist()
    acc.wait_for_everyone()
    out[rank] = res
    dist.destroy_process_group()


def test_world_size_2_gloo_plumbing():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    # sharding covers every (episode, gene) pair exactly once, in the reference's order
    all_pairs = [tuple(p) for p in r0['pairs'] + r1['pairs']]
    assert all_pairs == [(e, g) for e in range(4) for g in range(3)]
    assert len(r0['pairs']) == len(r1['pairs']) == 6
    assert r0['seed'] == r1['seed']
    assert r0['fitness'] == r1['fitness'] == r0['acc_reduce']
    want = [sum(e * 10 + g + (0.0 if i < 6 else 0.1) for i, (e, gg) in enumerate(all_pairs) if gg == g) for g in range(3)]
    assert r0['fitness'] == pytest.approx(want, rel=1e-6)
    assert r0['flat'] == r1['flat'] == [1.5 * i for i in range(10)]


def test_single_process_helpers_are_identity():
    import sys
    sys.path.insert(0, ROOT)
    D = importlib.import_module('x-transformers-rl_b200.distributed')
    assert not D.is_distributed()
    t = torch.tensor([1., 2.])
    assert D.reduce_fitness(t) is t
    assert D.all_reduce_mean_(t.clone()).tolist() == [1., 2.]
    assert D.shard_episode_genes(2, 1, 1, 0) == [[0, 0], [1, 0]]
    # the peer-memory gradient exchange is a multi-GPU construct: one process, or a CPU device, keeps the plain path
    assert D.PeerGroup.create(16, torch.device('cpu')) is None
    with pytest.raises(TypeError):
        D.PeerGroup()


def test_ema_schedule_matches_oracle():
    """host-side EMA schedule (schedule.py) == the oracle's EmaState, step for step."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import restated as R
    S = importlib.import_module('x-transformers-rl_b200.schedule')
    sched = S.EmaSchedule(beta=0.9, update_after_step=7, update_every=3, update_model_with_ema_every=11)
    ora = R.EmaState(beta=0.9, update_after_step=7, update_every=3, update_model_with_ema_every=11)
    ema, online = {'w': torch.zeros(3)}, {'w': torch.ones(3)}
    mine_ema, mine_online = torch.zeros(3), torch.ones(3)
    for step in range(40):
        online['w'] += 0.5; mine_online += 0.5
        mode, decay, back = sched.next()
        ora.update(ema, online)
        if mode == S.EMA_COPY:
            mine_ema = mine_online.clone()
        elif mode == S.EMA_LERP:
            mine_ema = torch.lerp(mine_ema, mine_online, 1. - decay)
        if back:
            mine_online = mine_ema.clone()
        assert torch.allclose(mine_ema, ema['w']) and torch.allclose(mine_online, online['w']), step
