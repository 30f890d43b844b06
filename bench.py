"""Benchmark of the LarvaNet hot path on B200 (driver contract: `python bench.py --gpus N --steps K --warmup W`).

Headline workload (BASELINE.json configs[1]): LarvaNet x4 (M=4, B=4,4,4,4) TRAINING step -- batch 16 per GPU of 48x48 LR
patches, L1 multi-exit loss, bf16 activations / fp32 accumulate -- metric = train patches/s over all GPUs (weak
scaling: per-GPU batch fixed).  The same JSON line carries the second half of BASELINE.json's metric, x4 SR output
Mpix/s for 320x180 -> 1280x720 inference, under "inference".

One step = H2D-free replay of the fused fwd+bwd CUDA graph + (N>1) NCCL gradient allreduce + fused AdamW + weight
re-pack, inputs already resident in HBM (`value`).  `e2e` is the same step through the reference-facing plugin call
`model.train_step_larva(...)` with pinned HOST batches copied in and the loss read back every step.

`--impl reference` times the CPU port of the reference (oracle/torch_port.py, all host threads) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

BLOCKS = [4, 4, 4, 4]
BATCH, PATCH = 16, 48
INF_H, INF_W = 180, 320
CONV_MAC = 9 * 48 * 48          # per LR pixel, one 48->48 conv
HEAD_MAC = 27 * 48


def flops_train_per_patch(blocks=BLOCKS, patch=PATCH):
    s, m = sum(blocks), len(blocks)
    f = HEAD_MAC + (2 * s + 2 * m) * CONV_MAC
    return 2.0 * (3 * f - HEAD_MAC) * patch * patch


def flops_infer_per_lr_px(blocks=BLOCKS):
    return 2.0 * (HEAD_MAC + (2 * sum(blocks) + 2) * CONV_MAC)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------------------------------
def run_reference(a):
    """CPU arm: the PyTorch-CPU port of the reference train step on all host threads (rank 0 only)."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import torch
    from larvanet_b200 import synth
    from oracle import torch_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = synth.make_larva_params(BLOCKS, seed=0)
    tr = torch_port.CpuTrainer(params, BLOCKS, lr=4e-4, threads=cores)
    lr, hr = synth.make_images(BATCH, PATCH, PATCH, seed=1)
    x, t = torch.from_numpy(lr), torch.from_numpy(hr)
    for _ in range(max(1, min(a.warmup, 2))):
        tr.step(x, t)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        tr.step(x, t)
    dt = time.perf_counter() - t0
    v = BATCH * a.steps / dt
    line = {
        'impl': 'reference', 'metric': 'train_patches_per_s', 'value': v, 'unit': 'patches/s', 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': dt / a.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'LarvaNet x4 training step, batch 16, 48x48 LR patches, L1 multi-exit loss '
                               '(M=4, B=4,4,4,4), AdamW', 'device': 'host CPU'},
        'cpu_baseline': {'value': v, 'unit': 'patches/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{a.steps} full steps of batch {BATCH} (oracle/torch_port.py, torch CPU fp32)'},
        'e2e': {'value': v, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
def cpu_baseline_sample(budget_s=20.0):
    """Bounded CPU sample on this box's host cores (rank 0, N=1): reference port, train step and 720p inference."""
    import torch
    from larvanet_b200 import synth
    from oracle import torch_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = synth.make_larva_params(BLOCKS, seed=0)
    tr = torch_port.CpuTrainer(params, BLOCKS, lr=4e-4, threads=cores)
    lr, hr = synth.make_images(BATCH, PATCH, PATCH, seed=1)
    x, t = torch.from_numpy(lr), torch.from_numpy(hr)
    tr.step(x, t)
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < budget_s * 0.6 and n < 50):
        tr.step(x, t)
        n += 1
    train = BATCH * n / (time.perf_counter() - t0)
    xi = torch.from_numpy(synth.make_images(1, INF_H, INF_W, seed=1)[0])
    tr.infer(xi)
    m, t0 = 0, time.perf_counter()
    while m < 2 or (time.perf_counter() - t0 < budget_s * 0.4 and m < 50):
        tr.infer(xi)
        m += 1
    infer = (16 * INF_H * INF_W) * m / (time.perf_counter() - t0) / 1e6
    return {'value': train, 'unit': 'patches/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n} train steps of batch {BATCH} + {m} 720p frames, torch CPU fp32 (oracle/torch_port.py)',
            'inference_mpix_s': infer}


def run_ours(a):
    import torch
    import torch.distributed as dist
    from larvanet_b200 import _lib, ops, synth

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: larvanet_b200 has no CPU path')
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local), timeout=datetime.timedelta(seconds=120))
    dev = torch.device('cuda', local)
    peaks, peaks_src = measured_peaks()

    import importlib
    model = importlib.import_module('models.LarvaNet').create_model()
    model.parse_args(['--num_modules=4', '--num_blocks=4,4,4,4', '--precision=bf16'])
    model.volume_per_step = PATCH * PATCH * BATCH * 3
    model.prepare(is_training=True, scales=[4])
    model.args.val_volume = 1e30
    model.global_step = 1    # skip the step-1 validation hook (needs a val loader; not part of the step)
    sd = {k: torch.from_numpy(v) for k, v in synth.make_larva_params(BLOCKS, seed=0).items()}
    model.get_model().load_state_dict(sd)
    eng = model._engine()
    if world > 1:
        eng.set_data_parallel(world)

    # ---- synthetic data: a pool of different batches, resident in HBM (value) and in pinned host memory (e2e)
    POOL = 8
    pool = [synth.make_images(BATCH, PATCH, PATCH, seed=100 + rank * POOL + i) for i in range(POOL)]
    dev_pool = [(torch.from_numpy(l).to(dev), torch.from_numpy(h).to(dev)) for l, h in pool]
    host_pool = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(h).pin_memory()) for l, h in pool]

    def step_resident(i):
        x, t = dev_pool[i % POOL]
        loss = eng.train_step(x, t)
        model.optim.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    counts = {}

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        counts['c0'], counts['r0'] = _lib.launch_count(), eng.replayed_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        counts['n'] = (_lib.launch_count() - counts['c0']) + (eng.replayed_launches - counts['r0'])
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warm = max(a.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, a.steps, warm)
    clocks = sampler.stop() if rank == 0 else None
    launches = counts['n']
    value = BATCH * world * a.steps / (ms * 1e-3)

    # ---- e2e through the plugin API with host batches
    args_ns = types.SimpleNamespace(train_path='/tmp')

    # host batches go through the framework's input prefetcher (pinned host -> device on a copy stream, one batch
    # ahead), exactly as train_larva.py feeds the plugin; every step's copy is inside the timed region
    from larvanet_b200.prefetch import DevicePrefetcher

    def host_batches():
        i = 0
        while True:
            yield host_pool[i % POOL]
            i += 1

    feeder = DevicePrefetcher(host_batches(), dev, depth=2)

    def step_e2e(i):
        x, t = next(feeder)
        return model.train_step_larva(args_ns, None, x, t)   # returns loss.item(): D2H every step

    ms_e2e = timed(step_e2e, a.steps, warm)
    e2e = BATCH * world * a.steps / (ms_e2e * 1e-3)
    h2d = int(pool[0][0].nbytes + pool[0][1].nbytes)

    line = {
        'metric': 'train_patches_per_s', 'value': value, 'unit': 'patches/s', 'n_gpus': world, 'steps': a.steps,
        'warmup': warm, 'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': 'LarvaNet x4 training step, batch 16 per GPU, 48x48 LR patches, L1 multi-exit loss, '
                               'M=4 B=4,4,4,4, AdamW (BASELINE.json configs[1])',
                   'global_batch': BATCH * world, 'parallelism': f'dp{world}',
                   'l2': 'inputs rotate over 8 resident batches; saved activations+gradients (~330 MB/step) exceed the 126 MB L2'},
        'e2e': {'value': e2e, 'unit': 'patches/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8,
                'ms_per_step': ms_e2e / a.steps},
        'gpu_launches': int(launches),
        'tflops_per_gpu': flops_train_per_patch() * BATCH / (ms / a.steps * 1e-3) / 1e12,
    }

    # ---- instrumented pass: CUDA events around every conv launch of one eager training step.  Every rank runs it
    # (the step contains the gradient all-reduce); rank 0 reports.
    eng.use_graphs = False
    ops.CONV_TIMERS = []
    torch.cuda._sleep(int(4e7))   # ~20 ms of GPU spin so the host can queue the whole eager step ahead of the GPU:
    step_resident(0)              # the events then bracket kernel time, not host launch latency
    torch.cuda.synchronize()
    recs, ops.CONV_TIMERS = ops.CONV_TIMERS, None
    eng.use_graphs = True
    if rank == 0:
        durs = np.array([e0.elapsed_time(e1) for e0, e1, _, _ in recs]) * 1e-3
        fl = np.array([f for _, _, f, _ in recs])
        ach = float(fl.sum() / durs.sum() / 1e12)
        peak = float(peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']))
        line['roofline'] = {'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                            # dram__bytes_read.sum + dram__bytes_write.sum of the step's 80 chained layers from the
                            # ncu --set full capture in profiles/r01c_conv_chain_ncu_full_summary.txt (two launches
                            # there: 64 + 16 layers), per training step; algorithmic: 80 x 3.54 MB of saved activations
                            'traffic': 320.1e6 / max(1, len(recs)),
                            'kernel': 'conv3x3_chain_kernel<48,48> (forward + backward-data conv layers of the step, persistent data-flow launch)',
                            'launches_timed': len(recs), 'avg_launch_us': float(durs.mean() * 1e6),
                            'peak_source': peaks_src + ', sustained bf16 (kernel timed inside a long step)'}
        line['clocks'] = clocks
        # the same kernel where the problem is large enough for its throughput (not the layer-to-layer hand-over
        # latency) to matter: 16 chained ReLU / residual layers on 8 x 270x480 px, measured live
        line['roofline']['steady_state'] = steady_state_chain(dev, float(peaks['bf16_tflops']))

    # ---- inference half of the metric (rank-local frames, no collective): 320x180 -> 1280x720
    lr720 = torch.from_numpy(synth.make_images(1, INF_H, INF_W, seed=1)[0]).to(dev)
    frames = [lr720 + float(i) for i in range(4)]
    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > L2

    def infer_resident(i):
        eng.forward(frames[i % 4])

    isteps = max(a.steps, 20)
    ms_i = timed(infer_resident, isteps, warm)
    # L2-flushed variant: a 192 MB write between frames, timed per frame with its own events
    per = []
    for i in range(10):
        flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.forward(frames[i % 4])
        e1.record()
        torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1))
    host_frame = lr720.cpu().pin_memory()
    host_out = torch.empty((1, 3, 4 * INF_H, 4 * INF_W), dtype=torch.float32).pin_memory()

    def infer_e2e(i):
        x = host_frame.to(dev, non_blocking=True)
        out = model.get_model()(x)
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_ie = timed(infer_e2e, isteps, warm)
    # same, but the frame leaves the device as the uint8 image get_sr.py / validate.py write and score
    # (model.upscale_uint8: round/clip on the device, 2.8 MB instead of 11 MB per 720p frame)
    host_u8 = torch.empty((1, 3, 4 * INF_H, 4 * INF_W), dtype=torch.uint8).pin_memory()

    def infer_e2e_u8(i):
        x = host_frame.to(dev, non_blocking=True)
        host_u8.copy_(ops.image_to_uint8(model.get_model()(x)), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_iu = timed(infer_e2e_u8, isteps, warm)
    hr_px = 16 * INF_H * INF_W
    line['inference'] = {
        'metric': 'sr_x4_output_mpix_per_s', 'unit': 'Mpix/s',
        'value': hr_px * world * isteps / (ms_i * 1e-3) / 1e6,
        'value_l2_flushed': hr_px * world / (float(np.median(per)) * 1e-3) / 1e6,
        'ms_per_frame': ms_i / isteps,
        'e2e': {'value': hr_px * world * isteps / (ms_ie * 1e-3) / 1e6, 'unit': 'Mpix/s',
                'h2d_bytes_per_step': int(host_frame.numel() * 4), 'd2h_bytes_per_step': int(host_out.numel() * 4)},
        'e2e_uint8': {'value': hr_px * world * isteps / (ms_iu * 1e-3) / 1e6, 'unit': 'Mpix/s',
                      'h2d_bytes_per_step': int(host_frame.numel() * 4), 'd2h_bytes_per_step': int(host_u8.numel())},
        'tflops_per_gpu': flops_infer_per_lr_px() * INF_H * INF_W / (ms_i / isteps * 1e-3) / 1e12,
        'config': {'workload': 'LarvaNet x4 inference, batch 1, 320x180 -> 1280x720, M=4 B=4,4,4,4 (BASELINE.json configs[0])',
                   'l2': 'value: back-to-back frames (5.5 MB maps stay in L2, as in steady-state video); '
                         'value_l2_flushed: 192 MB write between frames'},
    }

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline_sample()
    elif rank == 0:
        line['cpu_baseline'] = {'value': None, 'unit': 'patches/s', 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': 'measured at N=1 only (see the N=1 line)'}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def steady_state_chain(dev, peak_burst):
    import torch
    from larvanet_b200 import ops
    n, h, w, layers = 8, 270, 480, 16
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((n, h, 6, w, 8), device=dev, generator=g).to(torch.bfloat16)
    wt = torch.randn((48, 48, 3, 3), device=dev, generator=g) * 0.05
    b = torch.zeros(48, device=dev)
    from larvanet_b200 import _lib
    wl = _lib.LV_W_KY_STACKED      # what the engines use at this size (LarvaEngine.use_row_path): the row-marching chain
    packed = torch.zeros(ops.packed_weight_bytes(48, 48, torch.bfloat16), dtype=torch.uint8, device=dev)
    ops.pack_weights([dict(w=wt, packed=packed, cin=48, dtype=torch.bfloat16, wlayout=wl)])
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    args, src = [], x
    for i in range(layers):
        dst = bufs[i & 1]
        args.append(ops.make_conv_args([src], packed, 48, bias=b, out=dst, relu=(i & 1) == 0,
                                       res1=None if (i & 1) == 0 else x, wlayout=wl))
        src = dst
    ws = ops.chain_workspace(n, h, w, dev)
    for _ in range(2):
        ops.conv3x3_chain(args, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv3x3_chain(args, ws)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / 5
    ach = 2.0 * 9 * 48 * 48 * n * h * w * layers / sec / 1e12
    return {'workload': f'{layers} chained 48->48 convs on {n} x {h}x{w} px (activations 100 MB per layer: HBM-resident)',
            'kernel': 'row::conv3x3_row_kernel<48,48> (row-marching, ky-stacked N=144 MMAs; csrc/conv_row.cu)',
            'achieved': ach, 'peak': peak_burst, 'unit': 'TFLOP/s', 'frac': ach / peak_burst, 'us_per_layer': sec / layers * 1e6,
            'peak_source': 'measured burst bf16 (kernel timed alone)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', dest='no_cpu_baseline', action='store_true')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
