"""One process per GPU; torch.distributed (NCCL over NVLink 5 / NVSwitch on the box, gloo in the CPU
tests) is the plumbing.  The data path shards by episode and needs exactly four small collectives
 -- replacing accelerate + x_transformers_rl/distributed.py:1-99 of the reference:

  sync_seed            all-reduce(sum) of one random int64          distributed.py:42-48, :1211
  reduce_fitness       all-gather of per-rank (num_genes,) partials, summed in rank order (bit-identical
                       on every rank, unlike a ring all-reduce)     x_transformers_rl.py:1362
  all_reduce_grads     ONE all-reduce of the flat fp32 gradient buffer (+ the RSNorm batch mean packed in
                       its tail), averaged                          DDP in accelerator.prepare :885, :601
  barrier                                                          :1359

Nothing here touches CUDA directly: it works on whatever device the tensors live on, so the same code
runs under gloo on CPU tensors in tests/test_distributed_cpu.py.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class Accelerator:
    """the slice of accelerate.Accelerator the reference touches (x_transformers_rl.py:1106,
    train_lander.py:56): .device, .is_main_process, .num_processes, .process_index,
    .wait_for_everyone(), .reduce(), .log().  Launch with torchrun; rank / world come from the env."""

    def __init__(self, device=None, mixed_precision=None, backend=None, cpu=False, **unknown):
        if unknown:
            raise TypeError(f'accelerate_kwargs {sorted(unknown)} are not supported')
        world = int(os.environ.get('WORLD_SIZE', '1'))
        if world > 1 and not (dist.is_available() and dist.is_initialized()):
            use_cuda = torch.cuda.is_available() and not cpu
            kw = {}
            if use_cuda:
                local = int(os.environ.get('LOCAL_RANK', '0'))
                torch.cuda.set_device(local)
                kw['device_id'] = torch.device('cuda', local)       # rank <-> GPU binding is explicit, not guessed
            dist.init_process_group(backend or ('nccl' if use_cuda else 'gloo'), **kw)
        self.num_processes = dist.get_world_size() if is_distributed() else 1
        self.process_index = dist.get_rank() if is_distributed() else 0
        self.is_main_process = self.process_index == 0
        if device is not None:
            self.device = torch.device(device)
        elif torch.cuda.is_available() and not cpu:
            self.device = torch.device('cuda', torch.cuda.current_device())
        else:
            self.device = torch.device('cpu')
        if mixed_precision not in (None, 'no', 'fp32', 'bf16'):
            raise ValueError(f"mixed_precision={mixed_precision!r}: only None/'bf16' (tensor-core class) or 'no'/'fp32'")
        # Precision classes of the learn pass (north star: fp32 1e-5, bf16 1e-2).  The DEFAULT is the bf16 class: the
        # sequence model runs as libxtrl tcgen05 kernels (bf16 operands, fp32 accumulate, fp32 master weights).
        # mixed_precision='no' selects the fp32 parity class explicitly (the exact torch expression, oracle-pinned).
        self.mixed_precision = None if mixed_precision in ('no', 'fp32') else 'bf16'
        self.last_logs = None

    def wait_for_everyone(self):
        if is_distributed():
            dist.barrier()

    def reduce(self, t, reduction='sum'):
        out = reduce_fitness(t)
        return out / self.num_processes if reduction == 'mean' else out

    def log(self, logs):              # no tracker configured in the reference either
        self.last_logs = logs


def sync_seed(device, max_size=int(1e6)) -> int:
    """every rank draws an int, the sum is the shared seed (reference distributed.py:42-48)."""
    rand = torch.randint(0, max_size, (), device=device if torch.device(device).type == 'cuda' else 'cpu')
    if is_distributed():
        dist.all_reduce(rand)
    return int(rand.item())


def reduce_fitness(partial: torch.Tensor) -> torch.Tensor:
    """sum of per-rank partial fitness, deterministic: all-gather then add in rank order."""
    if not is_distributed():
        return partial
    parts = [torch.empty_like(partial) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, partial.contiguous())
    out = torch.zeros_like(partial)
    for p in parts:
        out += p
    return out


def all_reduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """in-place mean all-reduce of one flat buffer (gradients [+ RSNorm batch mean in the tail])."""
    if is_distributed():
        dist.all_reduce(flat)
        flat.div_(dist.get_world_size())
    return flat


def all_reduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    """in-place SUM all-reduce of the flat gradient buffer; the 1/world factor is applied by the consumers (the
    optimiser kernel scales gradients by xtrl_opt_cfg.grad_scale, the RSNorm tail is divided on its 9 floats)."""
    if is_distributed():
        dist.all_reduce(flat)
    return flat


class _DeviceBlock:
    """a raw device allocation seen by torch through __cuda_array_interface__ (torch.as_tensor keeps this object alive)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr='|u1', data=(ptr, False), version=3, strides=None)


class PeerGroup:
    """The flat gradient buffers of all ranks of one node, mapped into each other's address space (CUDA IPC over NVLink /
    NVSwitch), so that the optimiser kernel itself sums them: csrc/optim.cu, xtrl_adopt_atan2_step_peers.  Replaces DDP's
    gradient all-reduce (x_transformers_rl.py:885, :981) without an NCCL call on the critical path.

    local   this rank's [n_reduce] fp32 buffer (lives in the shared block; the model's .grad views point into it)
    reduced [n_reduce] fp32, the sum over ranks after a step
    Construction is collective (all ranks); it returns None from `create` when any rank could not set it up, in which
    case every rank keeps the NCCL all-reduce."""

    def __init__(self):
        raise TypeError('use PeerGroup.create')

    @classmethod
    def create(cls, n_reduce: int, device: torch.device):
        import ctypes as C
        from . import _lib as L
        if not is_distributed() or device.type != 'cuda' or dist.get_world_size() > 8:
            return None
        lib = L.load()
        world, rank = dist.get_world_size(), dist.get_rank()
        self = object.__new__(cls)
        self.lib, self.world, self.rank, self.n_reduce, self.epoch = lib, world, rank, n_reduce, 0
        self.flag_off = (n_reduce * 4 + 255) // 256 * 256
        nbytes = self.flag_off + 2 * world * 4
        self.base, self.opened, err = None, [], None
        handle = C.create_string_buffer(64)
        try:
            with torch.cuda.device(device):
                p = C.c_void_p()
                L.check(lib.xtrl_peer_alloc(nbytes, C.byref(p), handle), 'xtrl_peer_alloc')
                self.base = p.value
        except Exception as e:                                   # noqa: BLE001 -- reported to every rank below
            err = repr(e)
        gathered = [None] * world
        dist.all_gather_object(gathered, (handle.raw if err is None else None, err, _hostname()))
        ok = all(h is not None for h, _, _ in gathered) and len({hn for _, _, hn in gathered}) == 1
        ptrs = [None] * world
        if ok:
            try:
                with torch.cuda.device(device):
                    for r in range(world):
                        if r == rank:
                            ptrs[r] = self.base
                        else:
                            q = C.c_void_p()
                            L.check(lib.xtrl_peer_open(gathered[r][0], C.byref(q)), 'xtrl_peer_open')
                            ptrs[r] = q.value
                            self.opened.append(q.value)
            except Exception as e:                               # noqa: BLE001
                err = repr(e)
        flags = [None] * world
        dist.all_gather_object(flags, ok and err is None)
        if not all(flags):
            self.close()
            if rank == 0:
                import warnings
                warnings.warn('peer-memory gradient exchange unavailable (' + '; '.join(str(e) for _, e, _ in gathered if e) + f' {err or ""}); '
                              'using the NCCL all-reduce')
            return None
        self.grads_arr = (C.c_void_p * world)(*ptrs)
        self.flags_arr = (C.c_void_p * world)(*[q + self.flag_off for q in ptrs])
        block = torch.as_tensor(_DeviceBlock(self.base, nbytes), device=device)
        self.local = block[:n_reduce * 4].view(torch.float32)
        self.reduced = torch.zeros(n_reduce, device=device, dtype=torch.float32)
        self.status = torch.zeros(1, device=device, dtype=torch.int32)
        torch.cuda.synchronize(device)
        dist.barrier()
        return self

    def ctx(self, timeout_s: float = 60.):
        """the xtrl_peer_ctx of the next step (bumps the epoch; every rank must call in lock step)"""
        from . import _lib as L
        import ctypes as C
        self.epoch += 1
        return L.PeerCtx(C.cast(self.grads_arr, C.c_void_p), C.cast(self.flags_arr, C.c_void_p), self.reduced.data_ptr(),
                         self.status.data_ptr(), self.n_reduce, self.rank, self.world, self.epoch & 0xffffffff, timeout_s)

    def check(self):
        """raises if a step gave up waiting for a peer (one host sync: call once per learn(), not per step)"""
        if int(self.status.item()) != 0:
            from . import _lib as L
            raise L.XtrlError('peer-memory gradient exchange timed out waiting for another rank')

    def close(self):
        for q in self.opened:
            self.lib.xtrl_peer_close(q)
        self.opened = []

    def __del__(self):
        # the block itself is left to process exit: other ranks may still have it mapped
        try:
            self.close()
        except Exception:                                        # noqa: BLE001 -- interpreter shutdown
            pass


def _hostname() -> str:
    import socket
    return socket.gethostname()


def world_size() -> int:
    return dist.get_world_size() if is_distributed() else 1


def broadcast_from_rank0_(*tensors: torch.Tensor) -> None:
    """what accelerator.prepare (DDP) does at wrap time in the reference (x_transformers_rl.py:885): every replica starts
    from rank 0's values, whatever each rank's RNG produced locally."""
    if is_distributed():
        for t in tensors:
            dist.broadcast(t, src=0)


def shard_episode_genes(num_episodes_per_update: int, num_genes: int, num_processes: int, process_index: int):
    """the (episode, gene) pairs this rank rolls out: cartesian_prod(...).chunk(world)[rank]
    (x_transformers_rl.py:1146-1154)."""
    pairs = torch.cartesian_prod(torch.arange(num_episodes_per_update), torch.arange(num_genes))
    if pairs.ndim == 1:
        pairs = pairs[:, None]
    if pairs.size(0) % num_processes != 0:
        # unequal shards would make ranks issue different numbers of gradient all-reduces (a hang) and bias the
        # mean-of-means gradient; the reference hides this behind an all-gather + sharded DataLoader (:868-885)
        raise ValueError(f'num_episodes_per_update * num_genes = {pairs.size(0)} must be a multiple of the number of '
                         f'processes ({num_processes}): every rank learns from an equal shard of episodes')
    return pairs.chunk(num_processes, dim=0)[process_index].tolist()
