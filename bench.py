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


def roofline_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` summary of the final build
    (profiles/roofline_traffic.json, written from profiles/r02_step_ncu_full_summary.txt); None when there is no record."""
    path = os.path.join(REPO, 'profiles', 'roofline_traffic.json')
    try:
        with open(path) as f:
            for rec in json.load(f):
                if rec.get('key') == key:
                    return rec.get('dram_bytes_per_launch')
    except (OSError, ValueError):
        pass
    return None


def measured_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------------------------------
def headline_config(world):
    """The `config` object of the headline line -- shared by both arms so that the driver's same-config check holds."""
    return {'workload': 'LarvaNet x4 training step, batch 16 per GPU, 48x48 LR patches, L1 multi-exit loss, '
                        'M=4 B=4,4,4,4, AdamW (BASELINE.json configs[1])',
            'global_batch': BATCH * world, 'parallelism': f'dp{world}',
            'l2': 'inputs rotate over 8 resident batches; saved activations+gradients (~330 MB/step) exceed the 126 MB L2'}


def _cpu_trainer(blocks, v2=False, lr=4e-4, threads=None):
    """The reference's train step on the host: the UNMODIFIED reference modules when oracle/_ref is staged (kind
    "reference"), else the golden-pinned restatement oracle/torch_port.py (kind "port")."""
    from larvanet_b200 import synth
    from oracle import ref_loader, torch_port
    params = synth.make_larva_params(blocks, v2=v2, seed=0)
    if ref_loader.load() is not None:
        return ref_loader.RefTrainer(params, blocks, v2=v2, lr=lr), 'reference', 'reference modules via oracle/ref_loader.py'
    return torch_port.CpuTrainer(params, blocks, v2=v2, lr=lr, threads=threads), 'port', 'oracle/torch_port.py'


def run_reference(a):
    """CPU arm: the reference's own train step (its modules + torch.optim.AdamW, fp32) on all host threads, rank 0 only,
    on the headline config (global batch = 16 x --gpus, like our arm's weak scaling)."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import torch
    from larvanet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr, kind, what = _cpu_trainer(BLOCKS, threads=cores)
    gb = BATCH * max(1, a.gpus)
    lr, hr = synth.make_images(gb, PATCH, PATCH, seed=1)
    x, t = torch.from_numpy(lr), torch.from_numpy(hr)
    for _ in range(max(1, min(a.warmup, 2))):
        tr.step(x, t)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        tr.step(x, t)
    dt = time.perf_counter() - t0
    v = gb * a.steps / dt
    cfg = headline_config(max(1, a.gpus))
    line = {
        'impl': 'reference', 'metric': 'train_patches_per_s', 'value': v, 'unit': 'patches/s', 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': dt / a.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': cfg,
        'cpu_baseline': {'value': v, 'unit': 'patches/s', 'cores': cores, 'kind': kind,
                         'sample': f'{a.steps} full steps of batch {gb} on the host CPU ({what}, torch CPU fp32)'},
        'e2e': {'value': v, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
def cpu_baseline_sample(budget_s=20.0):
    """Bounded CPU sample on this box's host cores (rank 0, N=1): the reference's train step and 720p inference."""
    import torch
    from larvanet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr, kind, what = _cpu_trainer(BLOCKS, threads=cores)
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
    return {'value': train, 'unit': 'patches/s', 'cores': cores, 'kind': kind,
            'sample': f'{n} train steps of batch {BATCH} + {m} 720p frames, torch CPU fp32 ({what})',
            'inference_mpix_s': infer}


def gpu_library_baseline(dev, steps=20):
    """The library kernels to beat: the reference's modules on THIS B200 through stock PyTorch / cuDNN --
    eager fp32 (TF32 off), TF32, and bf16 autocast + channels_last -- CUDA-event timed like reference runtime.py:61-67
    (synchronise, time, synchronise).  cfg2 train step (patches/s) and cfg1 720p inference (Mpix/s)."""
    import torch
    from larvanet_b200 import synth
    from oracle import ref_loader, torch_port
    params = synth.make_larva_params(BLOCKS, seed=0)
    lr, hr = synth.make_images(BATCH, PATCH, PATCH, seed=1)
    x720 = torch.from_numpy(synth.make_images(1, INF_H, INF_W, seed=1)[0]).to(dev)
    out = {'what': 'reference modules (oracle/_ref) on the same GPU, stock torch ' + torch.__version__ + ' / cuDNN kernels, '
                   'cudnn.benchmark on' if ref_loader.load() is not None else 'oracle/torch_port.py on the same GPU',
           'modes': {}}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32, ac, cl in (('fp32', False, None, False), ('tf32', True, None, False),
                                   ('bf16_autocast_channels_last', True, torch.bfloat16, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            if ref_loader.load() is not None:
                tr = ref_loader.RefTrainer(params, BLOCKS, lr=4e-4, device=dev, autocast_dtype=ac, channels_last=cl)
                step = lambda a, b: tr.step(a, b, sync=False)
                infer = tr.infer
            else:
                if ac is not None:
                    continue
                tp = torch_port.CpuTrainer({k: v for k, v in params.items()}, BLOCKS, lr=4e-4)
                tp.p = {k: v.detach().to(dev).requires_grad_(True) for k, v in tp.p.items()}
                tp.optim = torch.optim.AdamW(list(tp.p.values()), lr=4e-4)
                step, infer = tp.step, tp.infer
            x, t = torch.from_numpy(lr).to(dev), torch.from_numpy(hr).to(dev)
            xi = x720
            if cl:
                x, xi = x.contiguous(memory_format=torch.channels_last), x720.contiguous(memory_format=torch.channels_last)
            res = {}
            for what, fn, units in (('train_patches_per_s', lambda: step(x, t), BATCH),
                                    ('infer_720p_mpix_per_s', lambda: infer(xi), 16 * INF_H * INF_W / 1e6)):
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[what] = units * steps / (e0.elapsed_time(e1) * 1e-3)
            out['modes'][name] = res
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out


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
    engines = [eng]

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        counts['c0'], counts['r0'] = _lib.launch_count(), sum(e.replayed_launches for e in engines)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        counts['n'] = (_lib.launch_count() - counts['c0']) + (sum(e.replayed_launches for e in engines) - counts['r0'])
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
    # the driver may ask for a very short timed region (20 steps = 11 ms): also time a long one and report both
    long_steps = max(200, a.steps)
    ms_long = timed(step_resident, long_steps, 2) if long_steps > a.steps else ms

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

    feeder = DevicePrefetcher(host_batches(), dev, depth=2, defer=True)

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
        'config': headline_config(world),
        'e2e': {'value': e2e, 'unit': 'patches/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8,
                'ms_per_step': ms_e2e / a.steps},
        'gpu_launches': int(launches),
        'tflops_per_gpu': flops_train_per_patch() * BATCH / (ms / a.steps * 1e-3) / 1e12,
        'long_run': {'steps': long_steps, 'value': BATCH * world * long_steps / (ms_long * 1e-3),
                     'ms_per_step': ms_long / long_steps},
        'dp_exchange': (None if world == 1 else
                        'fused peer-memory all-reduce + AdamW + re-pack kernel (lv_dp_adamw_pack_step)' if eng._dp_fused()
                        else 'symmetric-memory two-shot all-reduce' if eng._symm is not None else 'NCCL all-reduce'),
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
        # the launch is timed ALONE in an otherwise idle eager step (SM clocks at their maximum, no power cap): the
        # burst figure is the right denominator (the sustained one was measured at 1245 MHz under the 1 kW cap)
        peak = float(peaks['bf16_tflops'])
        line['roofline'] = {'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                            # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, read from the
                            # committed summary of the final build's `ncu --set full` capture (tools/ncu_summary.py
                            # writes profiles/roofline_traffic.json); null when no capture of this build exists
                            'traffic': roofline_traffic('cfg2_train_chain'),
                            'kernel': 'conv3x3_chain_kernel<48,48> (forward + backward-data conv layers of the step, persistent data-flow launch)',
                            'launches_timed': len(recs), 'avg_launch_us': float(durs.mean() * 1e6),
                            'peak_source': peaks_src + ', burst bf16 (kernel timed alone)'}
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
    # (model.upscale_uint8: round/clip in the exit conv's epilogue, 2.8 MB instead of 11 MB per 720p frame)
    host_u8 = torch.empty((1, 3, 4 * INF_H, 4 * INF_W), dtype=torch.uint8).pin_memory()

    def infer_e2e_u8(i):
        x = host_frame.to(dev, non_blocking=True)
        host_u8.copy_(eng.forward(x, uint8=True), non_blocking=True)   # uint8 frame from the exit conv's epilogue
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

    # ---- the other BASELINE.json configurations, each with its own `config.workload`
    del dev_pool, host_pool, feeder
    eng._train.clear()
    eng._infer.clear()
    torch.cuda.empty_cache()
    legs = {}
    if world == 1:
        legs['cfg3_edsr_1080p'] = leg_cfg3(dev, timed, engines, float(peaks['bf16_tflops']))
    legs['cfg4_v2_global128'] = leg_cfg4(world, rank, dev, timed, engines)
    legs['cfg5_256_frames'] = leg_cfg5(world, rank, dev, model, eng, barrier)
    line['configs'] = legs
    if rank == 0 and world == 1 and not a.no_gpu_baseline:
        line['gpu_library_baseline'] = gpu_library_baseline(dev)
        best = max(m['train_patches_per_s'] for m in line['gpu_library_baseline']['modes'].values())
        besti = max(m['infer_720p_mpix_per_s'] for m in line['gpu_library_baseline']['modes'].values())
        line['gpu_library_baseline']['ours_over_best_library'] = {'train': value / best,
                                                                  'infer_720p': line['inference']['value'] / besti}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline_sample()
    elif rank == 0:
        line['cpu_baseline'] = {'value': None, 'unit': 'patches/s', 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': 'measured at N=1 only (see the N=1 line)'}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _plugin(kind, flags, training):
    import importlib
    m = importlib.import_module('models.' + kind).create_model()
    m.parse_args(flags)
    m.prepare(is_training=training, scales=[4])
    return m


def leg_cfg3(dev, timed, engines, peak_burst):
    """BASELINE.json configs[2]: EDSR-baseline x4 (16 resblocks, 64 ch) inference, 480x270 -> 1920x1080."""
    import torch
    from larvanet_b200 import synth
    m = _plugin('edsr', ['--edsr_conv_features=64', '--edsr_res_blocks=16', '--precision=bf16'], False)
    m.get_model().load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_edsr_params(64, 16, 4, seed=0).items()})
    eng = m.get_model().engine()
    engines.append(eng)
    frames = [torch.from_numpy(synth.make_images(1, 270, 480, seed=7 + i)[0]).to(dev) for i in range(2)]
    steps = 20
    ms = timed(lambda i: eng.forward(frames[i & 1]), steps, 4)
    engines.remove(eng)
    hr_px = 16 * 270 * 480
    flops = 2.0 * 1983321 * 270 * 480          # 1,983,321 MAC per LR pixel
    sec = ms * 1e-3 / steps
    out = {'config': {'workload': 'EDSR-baseline x4 (16 resblocks, 64 ch) inference, batch 1, 480x270 -> 1920x1080 '
                                  '(BASELINE.json configs[2])',
                      'l2': 'back-to-back frames; 64-channel maps at 2x/4x resolution (66 / 265 MB) exceed the L2'},
           'metric': 'sr_x4_output_mpix_per_s', 'unit': 'Mpix/s', 'value': hr_px / sec / 1e6, 'ms_per_frame': sec * 1e3,
           'roofline': {'bound': 'tensor', 'achieved': flops / sec / 1e12, 'peak': peak_burst, 'unit': 'TFLOP/s',
                        'frac': flops / sec / 1e12 / peak_burst, 'traffic': None,   # several kernels: see profiles/
                        'kernel': 'whole network: row-marching chain (64->64 body) + conv3x3_tc (64->256 PixelShuffle, 64->3)'},
           'row_path': bool(eng.use_row_path(1, 270, 480))}
    del m, eng
    torch.cuda.empty_cache()
    return out


def leg_cfg4(world, rank, dev, timed, engines):
    """BASELINE.json configs[3]: LarvaNetV2 x4 data-parallel training, GLOBAL batch 128 of 64x64 patches (strong scaling:
    128/N patches per GPU), gradient exchange as in the headline."""
    import torch
    from larvanet_b200 import dist as lvdist, synth
    GB, P4 = 128, 64
    m = _plugin('LarvaNetV2', ['--num_modules=4', '--num_blocks=4,4,4,4', '--precision=bf16'], True)
    m.volume_per_step, m.global_step = P4 * P4 * GB * 3, 1
    m.args.val_volume = 1e30
    m.get_model().load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_larva_params(BLOCKS, v2=True, seed=0).items()})
    eng = m._engine()
    if world > 1:
        eng.set_data_parallel(world)
    engines.append(eng)
    b, e = lvdist.shard_range(GB, rank, world)
    pool = []
    for i in range(2):
        lr, hr = synth.make_images(e - b, P4, P4, seed=300 + 17 * rank + i)
        pool.append((torch.from_numpy(lr).to(dev), torch.from_numpy(hr).to(dev)))

    def step(i):
        x, t = pool[i & 1]
        loss = eng.train_step(x, t)
        m.optim.step()
        return loss

    steps = 10
    ms = timed(step, steps, 3)
    engines.remove(eng)
    sec = ms * 1e-3 / steps
    flops = 2.0 * 2864160 * P4 * P4 * GB        # 2,864,160 MAC per LR pixel per step
    out = {'config': {'workload': 'LarvaNetV2 x4 training step, GLOBAL batch 128 of 64x64 LR patches, M=4 B=4,4,4,4, L1 '
                                  'multi-exit + tail loss, AdamW (BASELINE.json configs[3])',
                      'global_batch': GB, 'per_gpu_batch': e - b, 'parallelism': f'dp{world}', 'scaling': 'strong'},
           'metric': 'train_patches_per_s', 'unit': 'patches/s', 'value': GB / sec, 'ms_per_step': sec * 1e3,
           'tflops_total': flops / sec / 1e12, 'row_path': bool(eng.use_row_path(e - b, P4, P4))}
    del m, eng, pool
    torch.cuda.empty_cache()
    return out


def leg_cfg5(world, rank, dev, model, eng, barrier):
    """BASELINE.json configs[4]: 256 synthetic 480x270 frames sharded round-robin over the GPUs (dist.frames_for_rank),
    no collective; sweep over the per-launch batch size."""
    import torch
    import torch.distributed as dist
    from larvanet_b200 import dist as lvdist, synth
    frames, H5, W5 = 256, 270, 480
    mine = lvdist.frames_for_rank(frames, rank, world)
    hr_px = 16 * H5 * W5
    sweep = {}
    for bs in (1, 8, 32):
        if bs > len(mine):
            continue
        x = torch.from_numpy(synth.make_images(bs, H5, W5, seed=500 + rank)[0]).to(dev)
        nb = len(mine) // bs
        for _ in range(3):
            eng.forward(x)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nb):
            eng.forward(x)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        done = nb * bs * world
        sweep[f'batch{bs}'] = {'mpix_per_s': done * hr_px / (float(ms.item()) * 1e-3) / 1e6, 'frames': done,
                               'ms_total': float(ms.item()), 'row_path': bool(eng.use_row_path(bs, H5, W5))}
        eng._infer.clear()
        del x
        torch.cuda.empty_cache()
    best = max(sweep.values(), key=lambda v: v['mpix_per_s'])
    return {'config': {'workload': 'LarvaNet x4 batch-sharded inference of 256 synthetic 480x270 frames -> 1920x1080, frames '
                                   'dealt round-robin to the GPUs, no collective (BASELINE.json configs[4])',
                       'frames_per_gpu': len(mine), 'parallelism': f'frames/{world}'},
            'metric': 'sr_x4_output_mpix_per_s', 'unit': 'Mpix/s', 'value': best['mpix_per_s'], 'sweep': sweep,
            'tflops_per_gpu': best['mpix_per_s'] * 1e6 / 16 * flops_infer_per_lr_px() / 1e12 / world}


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
            'traffic_per_layer': (roofline_traffic('steady_state_row_chain') or 0) / 4 or None,
            'achieved': ach, 'peak': peak_burst, 'unit': 'TFLOP/s', 'frac': ach / peak_burst, 'us_per_layer': sec / layers * 1e6,
            'peak_source': 'measured burst bf16 (kernel timed alone)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', dest='no_cpu_baseline', action='store_true')
    ap.add_argument('--no-gpu-baseline', dest='no_gpu_baseline', action='store_true')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
