"""GPU, >= 2 devices: data-parallel replicas stay bit-identical through rollout + learn (NCCL flat-gradient
all-reduce, fitness gather, synced evolve seed).  Skipped on a single-GPU box; the host-side logic of the
same path is covered on CPU by tests/test_distributed_cpu.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
@pytest.mark.parametrize('mixed,peers', (('fp32', '1'), ('bf16', '1'), ('bf16', '0')))
def test_data_parallel_replicas_stay_identical(mixed, peers):
    """peers='1': the gradient sum over ranks runs inside the optimiser kernel over CUDA-IPC peer memory (and is checked
    against NCCL + the plain launch on the same inputs); '0': torch.distributed.all_reduce."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tests', 'multi', 'dp_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, XTRL_DP_MIXED=mixed, XTRL_DP_PEERS=peers))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'DP_RESULT identical=True changed=True finite=True' in res.stdout
    assert f'peers={peers == "1"}' in res.stdout, res.stdout[-500:]
