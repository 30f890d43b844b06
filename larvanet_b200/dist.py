"""Data-parallel plumbing for the training path (one process per GPU, torch.distributed over NCCL/NVLink).

The reference is single-process; batch-sharded training is a new capability whose contract is:
the result of G ranks, each holding B/G patches, must equal the reference's single-process step on the global batch B.
With a mean-reduced L1 loss that means every rank scales its local gradient SUM by 1/(numel_global * exits) -- done
inside the weight-gradient kernels via `grad_scale()` -- and the exchange is a plain SUM all-reduce of the flat gradient
arena (3.3 MB fp32 for LarvaNet M=4: latency-bound, one bucket) plus the 8-byte loss sum.  Inference shards frames
round-robin and needs no collective.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(total, rank, world):
    """[begin, end) of the contiguous shard of `total` items owned by `rank` (first `total % world` ranks get one more)."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def grad_scale(local_hr_numel, world, num_exits):
    """Scale applied to a rank's local sum of sign(out - truth) so that the all-reduced SUM equals the gradient of the
    reference's mean-reduced multi-exit loss over the global batch (reference models/LarvaNet.py:104-109)."""
    return 1.0 / (float(local_hr_numel) * world * num_exits)


def allreduce_gradients(flat_grad, loss_sum=None, group=None):
    """SUM all-reduce of the flat gradient arena (and the loss accumulator).  Returns the async work handles."""
    works = [dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=True)]
    if loss_sum is not None:
        works.append(dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group, async_op=True))
    return works


class SymmetricGradExchange:
    """In-place SUM all-reduce of the gradient arena through symmetric (peer-mapped) memory over NVLink: the two-shot
    kernel of torch.distributed._symmetric_memory (every rank reduces 1/G of the buffer straight out of its peers' HBM,
    then every rank gathers), stream-ordered, ~3x lower latency than NCCL's ring / tree for this 3.3 MB latency-bound
    message.  The arena has to live in a symmetric allocation: `buffer` replaces the engine's gradient arena.

    Construction is collective.  Every rank reports success or failure and the group only switches over if ALL ranks
    succeeded; otherwise everybody keeps the NCCL all-reduce (`allreduce_gradients`)."""

    TAIL = 96   # floats behind the gradient arena: [0:2] this rank's loss (one double), [16:64] device-barrier flags

    def __init__(self, numel, device, group=None):
        """Local half of the construction (no collective): allocate this rank's symmetric buffer."""
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.group_name = self.group.group_name
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        pad = 4 * 32 * self.world                  # 16-byte vectors x 32 lanes x ranks: keeps every rank's share aligned
        self.numel = (int(numel) + pad - 1) // pad * pad
        # [gradient arena | tail | reduced slices (two-shot exchange: rank r writes slice r of its copy)]
        self.storage = symm_mem.empty(2 * self.numel + self.TAIL, dtype=torch.float32, device=device)
        self.storage.zero_()
        self.buffer = self.storage[:self.numel]    # the gradient arena the engine writes into
        self.loss_view = self.storage[self.numel:self.numel + 2].view(torch.float64)   # this rank's loss accumulator
        self.handle = None
        self.peers = None

    def rendezvous(self):
        """Collective half: exchange the peer mappings.  Only entered once EVERY rank has allocated successfully."""
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        self.handle = symm_mem.rendezvous(self.storage, self.group_name)
        # raw peer addresses for the fused all-reduce + AdamW kernel (lv_dp_adamw_pack_step); None if the handle does
        # not expose them (then the engine keeps the two-shot all-reduce + separate optimizer launch)
        try:
            bases = [int(p) + int(getattr(self.handle, 'offset', 0)) for p in self.handle.buffer_ptrs]
            if len(bases) == self.world and bases[self.rank] == self.storage.data_ptr():
                self.peers = PeerArena(self, bases, C)
        except Exception:  # noqa: BLE001
            self.peers = None

    def allreduce_(self):
        torch.ops.symm_mem.two_shot_all_reduce_(self.buffer, 'sum', self.group_name)

    @staticmethod
    def _agree(ok, device, group):
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return int(flag.item()) == 1

    @staticmethod
    def try_create(numel, device, group=None):
        """Collective: returns an exchange on every rank or None on every rank.  Two agreement rounds (plain NCCL MIN
        all-reduces) bracket the symmetric-memory collectives, so a rank whose LOCAL allocation fails (OOM, unsupported)
        never leaves its peers blocked inside the rendezvous: nobody enters it unless everybody allocated."""
        import sys
        ex, why = None, None
        if os.environ.get('LARVANET_B200_SYMM_ALLREDUCE', '1') == '0':
            why = 'disabled by LARVANET_B200_SYMM_ALLREDUCE=0'
        else:
            try:
                ex = SymmetricGradExchange(numel, device, group)
            except Exception as e:  # noqa: BLE001 -- any failure means "use NCCL", agreed on collectively below
                ex, why = None, f'{type(e).__name__}: {e}'
        if not SymmetricGradExchange._agree(ex is not None, device, group):
            if why is not None and not why.startswith('disabled'):
                print(f'[larvanet_b200] symmetric-memory allocation failed here ({why}); every rank uses NCCL', file=sys.stderr)
            return None
        ok = True
        try:
            ex.rendezvous()
            ex.allreduce_()                       # one warm-up exchange (zeros) also proves the kernel runs here
            torch.cuda.synchronize(device)
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f'[larvanet_b200] symmetric-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL', file=sys.stderr)
        if not SymmetricGradExchange._agree(ok, device, group):
            return None
        # the fused exchange+optimizer kernel is used only if EVERY rank has the peer addresses
        if not SymmetricGradExchange._agree(ex.peers is not None and ex.peers.supported, device, group):
            ex.peers = None
        return ex


class PeerArena:
    """Raw peer-mapped addresses of every rank's gradient arena / loss slot / flag block (ctypes arrays, index = rank)
    plus this rank's device-local control words: the arguments of `ops.dp_adamw_pack_step`."""

    def __init__(self, ex, bases, C):
        self.world, self.rank = ex.world, ex.rank
        n = ex.numel
        self.grad_ptrs = (C.c_void_p * self.world)(*bases)
        self.loss_ptrs = (C.c_void_p * self.world)(*[b + n * 4 for b in bases])
        self.flag_ptrs = (C.c_void_p * self.world)(*[b + (n + 16) * 4 for b in bases])
        self.reduced_ptrs = (C.c_void_p * self.world)(*[b + (n + ex.TAIL) * 4 for b in bases])
        self.slice = n // self.world                      # numel is padded to a multiple of 128 * world
        # two-shot at 8 ranks (one-shot would read 7 x 3.3 MB per rank: ~30 us; measured step 0.582 vs 0.600 ms); at 4 ranks
        # the extra barrier costs more than the 6.6 MB it saves (0.576 vs 0.568 ms).  LARVANET_B200_DP_TWO_SHOT=0/1 forces.
        mode = os.environ.get('LARVANET_B200_DP_TWO_SHOT', 'auto')
        self.two_shot = (self.world >= 8) if mode == 'auto' else mode == '1'
        self.loss_out = torch.zeros(1, dtype=torch.float64, device=ex.storage.device)   # sum over ranks, written by the kernel
        self.ctl = torch.zeros(8, dtype=torch.int32, device=ex.storage.device)
        self.supported = self.world in (2, 4, 8)


def frames_for_rank(num_frames, rank, world):
    """Round-robin frame indices for batch-sharded inference (no collective)."""
    return list(range(rank, num_frames, world))
