"""world_size-2 gloo test of the data-parallel contract on CPU (host logic only; the per-rank gradient comes from the
numpy oracle): two ranks, each with half of the global batch and `grad_scale`, all-reduce SUM == single-process
gradient of the reference's mean-reduced multi-exit loss on the whole batch (invariant vi)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from larvanet_b200 import dist as lvdist
from larvanet_b200 import synth
from oracle import larva_oracle as O

BLOCKS = [1, 1]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = lvdist.init_from_env('gloo')
    assert (r, w) == (rank, world)
    params = synth.make_larva_params(BLOCKS, seed=21, bias_std=0.02)
    lr, hr = synth.make_images(4, 8, 8, seed=22)
    b, e = lvdist.shard_range(4, rank, world)
    # per-rank oracle gradient of the LOCAL mean loss, then rescaled to the DP contract
    loss, grads, _ = O.larvanet_train_step(params, lr[b:e], hr[b:e], BLOCKS)
    local_numel = hr[b:e].size
    m = len(BLOCKS)
    # oracle grads are d(mean over local numel)/dp = sum(sign)/(local_numel*m); the kernels emit sum(sign)*grad_scale
    factor = lvdist.grad_scale(local_numel, world, m) * (local_numel * m)
    flat = torch.cat([torch.from_numpy(grads[k].ravel() * factor) for k in params])
    loss_sum = torch.tensor([loss * local_numel * m], dtype=torch.float64)
    for wk in lvdist.allreduce_gradients(flat, loss_sum):
        wk.wait()
    if rank == 0:
        np.savez(out_path, flat=flat.numpy(), loss=loss_sum.numpy() / (hr.size * m))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process(tmp_path):
    out = str(tmp_path / 'dp.npz')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    params = synth.make_larva_params(BLOCKS, seed=21, bias_std=0.02)
    lr, hr = synth.make_images(4, 8, 8, seed=22)
    loss, grads, _ = O.larvanet_train_step(params, lr, hr, BLOCKS)
    ref = np.concatenate([grads[k].ravel() for k in params])
    np.testing.assert_allclose(got['flat'], ref, rtol=1e-10, atol=1e-14)
    assert abs(float(got['loss'][0]) - loss) < 1e-12 * loss


def test_sharding_helpers():
    assert [lvdist.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert lvdist.shard_range(2, 3, 4) == (2, 2)
    assert lvdist.frames_for_rank(7, 1, 3) == [1, 4]
    assert lvdist.grad_scale(100, 4, 5) == 1.0 / 2000.0


def _symm_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    lvdist.init_from_env('gloo')
    # no CUDA here: creating the symmetric allocation fails on every rank, the ranks AGREE on that through the MIN
    # all-reduce and everybody keeps the plain all-reduce path (nobody may end up alone in a collective)
    ex = lvdist.SymmetricGradExchange.try_create(1000, torch.device('cpu'))
    flat = torch.full((8,), float(rank + 1), dtype=torch.float32)
    for wk in lvdist.allreduce_gradients(flat):
        wk.wait()
    if rank == 0:
        np.savez(out_path, none=np.array([ex is None]), flat=flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_symmetric_exchange_falls_back_collectively(tmp_path):
    out = str(tmp_path / 'symm.npz')
    mp.spawn(_symm_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    assert bool(got['none'][0])
    np.testing.assert_array_equal(got['flat'], np.full(8, 3.0, dtype=np.float32))


def _symm_asym_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    lvdist.init_from_env('gloo')
    entered = []

    # rank 0 "allocates" fine, rank 1's local allocation throws: rank 0 must NOT enter the rendezvous alone
    def fake_init(self, numel, device, group=None):
        if rank == 1:
            raise MemoryError('simulated symm_mem.empty failure on one rank')
        self.buffer = torch.zeros(8)

    def fake_rendezvous(self):
        entered.append(rank)
        raise AssertionError('rendezvous entered although a peer failed to allocate')

    lvdist.SymmetricGradExchange.__init__ = fake_init
    lvdist.SymmetricGradExchange.rendezvous = fake_rendezvous
    ex = lvdist.SymmetricGradExchange.try_create(1000, torch.device('cpu'))
    ok = torch.tensor([1 if (ex is None and not entered) else 0], dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.savez(out_path, ok=ok.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_symmetric_exchange_one_rank_allocation_failure_does_not_deadlock(tmp_path):
    out = str(tmp_path / 'asym.npz')
    mp.spawn(_symm_asym_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert int(np.load(out)['ok'][0]) == 1
