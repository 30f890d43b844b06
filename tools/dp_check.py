"""Multi-GPU parity check (run under torchrun): G ranks x (B/G) patches with the NCCL gradient all-reduce must equal
one rank on the global batch B.  Prints the worst relative gradient difference on rank 0."""
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import dist as lvdist, synth

def make(blocks, v2):
    m = importlib.import_module('models.LarvaNetV2' if v2 else 'models.LarvaNet').create_model()
    m.parse_args([f'--num_modules={len(blocks)}', '--num_blocks=' + ','.join(map(str, blocks))])
    m.prepare(is_training=True, scales=[4])
    return m

def main():
    rank, world, local = lvdist.init_from_env('nccl')
    torch.cuda.set_device(local)
    for v2 in (False, True):
        blocks = [2, 2]
        params = synth.make_larva_params(blocks, v2=v2, seed=7, bias_std=0.02)
        lr, hr = synth.make_images(8, 32, 32, seed=8)
        sd = {k: torch.from_numpy(v) for k, v in params.items()}
        m = make(blocks, v2); m.get_model().load_state_dict(sd)
        eng = m._engine(); eng.set_data_parallel(world)
        b, e = lvdist.shard_range(8, rank, world)
        loss = eng.train_step(torch.from_numpy(lr[b:e]).cuda(), torch.from_numpy(hr[b:e]).cuda()).item()
        g_dp = eng.arena.grad[:eng.arena.total].clone()
        # single-rank reference on the whole batch (same process, world_size 1)
        m1 = make(blocks, v2); m1.get_model().load_state_dict(sd)
        e1 = m1._engine()
        loss1 = e1.train_step(torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda()).item()
        g1 = e1.arena.grad
        rel = ((g_dp - g1).norm() / g1.norm()).item()
        if rank == 0:
            print(f'v2={v2} world={world}: loss dp={loss:.6f} single={loss1:.6f} rel grad diff={rel:.3e}', flush=True)
        assert abs(loss - loss1) <= 1e-6 * abs(loss1) and rel < 2e-3, (loss, loss1, rel)
    dist.barrier(); dist.destroy_process_group()

main()
