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
    # phase 1 checks the gradients right after engine.train_step(): use the stand-alone exchange (two-shot peer kernel or
    # NCCL); the exchange fused into the optimizer kernel is phase 2
    fused_env = os.environ.get('LARVANET_B200_DP_FUSED')
    os.environ['LARVANET_B200_DP_FUSED'] = '0'
    # (v2, global batch): 7 patches do not divide over the ranks -> uneven shards, scaled by the all-reduced global count
    for v2, total in ((False, 8), (True, 8), (False, 7)):
        blocks = [2, 2]
        params = synth.make_larva_params(blocks, v2=v2, seed=7, bias_std=0.02)
        lr, hr = synth.make_images(total, 32, 32, seed=8)
        sd = {k: torch.from_numpy(v) for k, v in params.items()}
        m = make(blocks, v2); m.get_model().load_state_dict(sd)
        eng = m._engine(); eng.set_data_parallel(world)
        b, e = lvdist.shard_range(total, rank, world)
        loss = eng.train_step(torch.from_numpy(lr[b:e]).cuda(), torch.from_numpy(hr[b:e]).cuda()).item()
        g_dp = eng.arena.grad[:eng.arena.total].clone()
        # single-rank reference on the whole batch (same process, world_size 1)
        m1 = make(blocks, v2); m1.get_model().load_state_dict(sd)
        e1 = m1._engine()
        loss1 = e1.train_step(torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda()).item()
        g1 = e1.arena.grad
        rel = ((g_dp - g1).norm() / g1.norm()).item()
        if rank == 0:
            print(f'v2={v2} batch={total} world={world}: loss dp={loss:.6f} single={loss1:.6f} rel grad diff={rel:.3e}', flush=True)
        assert abs(loss - loss1) <= 1e-6 * abs(loss1) and rel < 2e-3, (loss, loss1, rel)
    if fused_env is None:
        del os.environ['LARVANET_B200_DP_FUSED']
    else:
        os.environ['LARVANET_B200_DP_FUSED'] = fused_env
    # ---- optimizer steps: the fused exchange + AdamW + re-pack kernel (peer memory) against a single process on the
    # global batch, and bit-identical replicas across ranks
    import types
    blocks = [2, 2]
    params = synth.make_larva_params(blocks, seed=9, bias_std=0.02)
    sd = {k: torch.from_numpy(v) for k, v in params.items()}
    batches = [synth.make_smooth_images(8, 32, 32, seed=30 + i) for i in range(4)]
    ns = types.SimpleNamespace(train_path='/tmp')

    def prep(m):
        m.volume_per_step, m.global_step = 1, 5
        m.args.val_volume = 1e30
        return m

    mdp = prep(make(blocks, False)); mdp.get_model().load_state_dict(sd)
    edp = mdp._engine(); edp.set_data_parallel(world)
    m1 = prep(make(blocks, False)); m1.get_model().load_state_dict(sd)
    m1._engine()
    b, e = lvdist.shard_range(8, rank, world)
    for lr, hr in batches:
        l_dp = mdp.train_step_larva(ns, None, torch.from_numpy(lr[b:e]).cuda(), torch.from_numpy(hr[b:e]).cuda())
        l_1 = m1.train_step_larva(ns, None, torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda())
        assert abs(l_dp - l_1) <= 1e-5 * abs(l_1), (l_dp, l_1)
    p_dp, p_1 = edp.arena.flat, m1._engine().arena.flat
    moved = (p_1 - torch.cat([sd[k].flatten() for k in sd]).cuda()).norm().item()
    drift = (p_dp - p_1).norm().item() / moved
    hi, lo = p_dp.clone(), p_dp.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    same = bool(torch.equal(hi, lo))
    if rank == 0:
        print(f'optimizer: fused={edp._dp_fused()} 4 steps, |dp - single| / |update| = {drift:.3e}, replicas bit-identical: {same}',
              flush=True)
    assert same and drift < 5e-2, (same, drift)
    dist.barrier(); dist.destroy_process_group()

main()
