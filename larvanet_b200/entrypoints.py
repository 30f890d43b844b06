"""Host-side bodies of the entry points (`train_larva.py`, `train_larvaV2.py`, `validate.py`, `get_sr.py`, `runtime.py`).

The scripts at the repo root keep the reference's names and command lines (reference README.md:28-63) and are thin
wrappers around the functions here.  What the reference's scripts have in common -- the flag chain (script parser ->
data loader parser -> model parser, reference train_larva.py:49,60,71), plugin discovery by module name
(`importlib.import_module('models.' + name).create_model()`), CUDA_VISIBLE_DEVICES handling -- lives in `_flag_parser`,
`_open_loader` and `_open_model`; flags are declared in tables instead of one `add_argument` line each.
Differences from the reference, all on the host side: the default loaders are the synthetic ones (DIV2K readers are out
of scope), `--max_steps` is honoured (the reference loops until Ctrl-C), TensorBoard is optional, input batches are
prefetched to the device, uint8 conversion and validation PSNR run on the device when the plugin offers it.
"""
import argparse
import importlib
import json
import os
import time

import numpy as np

# (flag, type, default, help) -- `bool` means store_true, default None with required=True is spelled REQUIRED
REQUIRED = object()
COMMON_FLAGS = [
    ('model', str, 'LarvaNet', 'Name of the model.'),
    ('cuda_device', str, '0', 'Value for CUDA_VISIBLE_DEVICES.'),
]
SCALES_FLAG = [('scales', str, '4', 'Comma separated scales.')]
RESTORE_FLAGS = [
    ('restore_path', str, REQUIRED, 'Checkpoint to use.'),
    ('restore_target', str, None, 'Target of the restoration.'),
    ('restore_global_step', int, 0, 'Global step of the checkpoint.'),
]
TRAIN_FLAGS = [
    ('dataloader', str, 'synthetic_loader', 'Name of the data loader.'),
    ('val_dataloader', str, 'synthetic_val_loader', 'Name of the validation loader.'),
    ('batch_size', int, 16, 'Patches per training step.'),
    ('input_patch_size', int, 48, 'LR patch edge.'),
    ('train_path', str, '/tmp/larvanet_b200/train/', 'Checkpoint/summary directory.'),
    ('max_steps', int, 300000, 'Stop after this many steps.'),
    ('log_freq', int, 10, 'Print every N steps.'),
    ('summary_freq', int, 1000, 'TensorBoard summary period.'),
    ('save_freq', int, 10000, 'Checkpoint period (the model also saves by volume).'),
    ('sleep_ratio', float, 0.05, 'Idle fraction per step (0 disables).'),
    ('restore_path', str, None, 'Checkpoint to start from.'),
    ('restore_target', str, None, 'Target of the restoration.'),
    ('global_step', int, 0, 'Initial global step when resuming.'),
]
VALIDATE_FLAGS = [
    ('dataloader', str, 'synthetic_val_loader', 'Name of the data loader.'),
    ('save_path', str, None, 'Write the upscaled PNGs below this directory.'),
    ('chop_forward', bool, False, 'Upscale in four overlapping quadrants.'),
    ('chop_overlap_size', int, 20, 'Quadrant overlap in LR pixels (even).'),
    ('exact_bands', int, 0, 'Upscale in K horizontal bands with a receptive-field halo: bounded memory, result identical to '
                            'the full frame (larvanet_b200/tiling.py; --chop_forward is the reference\'s inexact scheme).'),
]
GET_SR_FLAGS = [
    ('exact_bands', int, 0, 'Upscale in K horizontal bands with a receptive-field halo (exact; bounds memory on big frames).'),
    ('scale', int, 4, 'Upscaling factor.'),
    ('input_path', str, 'LR', 'Directory of input PNGs.'),
    ('output_path', str, 'SR', 'Directory for the upscaled PNGs.'),
]
RUNTIME_FLAGS = [('dataloader', str, 'synthetic_val_loader', 'Name of the data loader.')]


def _flag_parser(*tables):
    parser = argparse.ArgumentParser()
    for table in tables:
        for name, kind, default, text in table:
            if kind is bool:
                parser.add_argument('--' + name, action='store_true', help=text)
            elif default is REQUIRED:
                parser.add_argument('--' + name, type=kind, required=True, help=text)
            else:
                parser.add_argument('--' + name, type=kind, default=default, help=text)
    return parser


def _select_device(cuda_device):
    if 'LOCAL_RANK' not in os.environ:          # under torchrun the launcher has already assigned the GPU
        os.environ['CUDA_VISIBLE_DEVICES'] = cuda_device


def _open_loader(name, scales, argv=None):
    """Create + prepare a data loader plugin; returns (loader, its parsed args, unconsumed argv)."""
    print('prepare data loader - %s' % name)
    loader = importlib.import_module('dataloaders.' + name).create_loader()
    parsed, rest = (None, argv)
    if argv is not None:
        parsed, rest = loader.parse_args(argv)
    loader.prepare(scales=scales)
    return loader, parsed, rest


def _open_model(name, argv, prepare_kwargs, before_prepare=None, restore=None):
    """Create a model plugin, let it parse its flags, prepare it, optionally restore a checkpoint."""
    print('prepare model - %s' % name)
    model = importlib.import_module('models.' + name).create_model()
    parsed, rest = model.parse_args(argv)
    if before_prepare is not None:
        before_prepare(model)
    model.prepare(**prepare_kwargs)
    if rest:
        print('WARNING: found unhandled arguments: %s' % rest)
    if restore is not None and restore[0] is not None:
        model.restore(ckpt_path=restore[0], target=restore[1])
        print('restored the model')
    return model, parsed


def _image_to_uint8(image):
    """Round + clip to uint8 (reference validate.py:17-18)."""
    return np.clip(np.round(image), a_min=0, a_max=255).astype(np.uint8)


def _fit_truth_image_size(output_image, truth_image):
    """Crop the truth to the output's size (reference validate.py:20-21)."""
    return truth_image[:, 0:output_image.shape[1], 0:output_image.shape[2]]


def _image_psnr(output_image, truth_image):
    """10*log10(255^2 / MSE) over all RGB pixels (reference validate.py:23-27)."""
    diff = np.float32(truth_image) - np.float32(output_image)
    mse = np.mean(np.power(diff, 2))
    return 10.0 * np.log10(255.0 ** 2 / mse)


def _upscale_banded_uint8(model, image, bands):
    """uint8 CHW frame through larvanet_b200.tiling (exact band-sharded inference); None if the plugin has no engine."""
    net = model.get_model()
    if not hasattr(net, 'engine'):
        return None
    import torch
    from larvanet_b200 import tiling
    x = torch.as_tensor(np.asarray([image]), dtype=torch.float32, device=model.device)
    return tiling.upscale_banded(net.engine(), x, bands, exit_leg=getattr(net, 'leg', None), uint8=True)[0].cpu().numpy()


def _save_image(image, path):
    import cv2 as cv
    cv.imwrite(path, cv.cvtColor(np.transpose(image, [1, 2, 0]), cv.COLOR_RGB2BGR))


def validate_main(argv=None):
    """reference validate.py:30-107: PSNR of every validation image, optional PNG dump / chop-forward."""
    import torch
    from utils import image_utils

    args, rest = _flag_parser(COMMON_FLAGS, SCALES_FLAG, RESTORE_FLAGS, VALIDATE_FLAGS).parse_known_args(argv)
    _select_device(args.cuda_device)
    scale_list = [int(s) for s in args.scales.split(',')]
    dataloader, _, _ = _open_loader(args.dataloader, scale_list)
    model, _ = _open_model(args.model, rest, dict(is_training=False, scales=scale_list, global_step=args.restore_global_step),
                           restore=(args.restore_path, args.restore_target))

    print('begin validation')
    results = {}
    for scale in scale_list:
        durations, psnrs = [], []
        count = dataloader.get_num_images()
        with torch.no_grad():
            for image_index in range(count):
                input_image, truth_image, image_name = dataloader.get_image_pair(image_index=image_index, scale=scale)
                start = time.perf_counter()
                banded = _upscale_banded_uint8(model, input_image, args.exact_bands) if args.exact_bands > 0 else None
                if banded is not None:
                    output_image = banded
                elif args.chop_forward:
                    output_image = image_utils.upscale_with_chop_forward(model=model, input_image=input_image, scale=scale,
                                                                         overlap_size=args.chop_overlap_size)
                elif hasattr(model, 'upscale_uint8'):   # round/clip on the device (== _image_to_uint8 below)
                    output_image = model.upscale_uint8(input_list=[input_image], scale=scale)[0]
                else:
                    output_image = model.upscale(input_list=[input_image], scale=scale)[0]
                durations.append(time.perf_counter() - start)
                output_image = _image_to_uint8(output_image)
                if args.save_path is not None:
                    target_dir = os.path.join(args.save_path, 'x%d' % scale)
                    os.makedirs(target_dir, exist_ok=True)
                    _save_image(output_image, os.path.join(target_dir, image_name + '.png'))
                truth_image = _fit_truth_image_size(output_image=output_image, truth_image=_image_to_uint8(truth_image))
                psnrs.append(_image_psnr(output_image=output_image, truth_image=truth_image))
                print('x%d, %d/%d, psnr=%.2f, duration=%.4f' % (scale, image_index + 1, count, psnrs[-1], durations[-1]))
        results[scale] = (float(np.mean(psnrs)), float(np.mean(durations)))
        print('x%d, psnr=%.2f, duration=%.4f' % (scale, results[scale][0], results[scale][1]))
    print('finished')
    return results


def get_sr_main(argv=None):
    """reference get_sr.py:30-92: every PNG of --input_path -> x4 PNG in --output_path."""
    import cv2 as cv
    import torch

    args, rest = _flag_parser(COMMON_FLAGS, RESTORE_FLAGS, GET_SR_FLAGS).parse_known_args(argv)
    _select_device(args.cuda_device)
    os.makedirs(args.output_path, exist_ok=True)
    model, _ = _open_model(args.model, rest, dict(is_training=False, scales=[args.scale], global_step=args.restore_global_step),
                           restore=(args.restore_path, args.restore_target))

    names = sorted(f for f in os.listdir(args.input_path) if f.lower().endswith('.png'))
    durations = []
    with torch.no_grad():
        for i, name in enumerate(names):
            image = np.transpose(cv.cvtColor(cv.imread(os.path.join(args.input_path, name)), cv.COLOR_BGR2RGB), [2, 0, 1])
            start = time.perf_counter()
            banded = _upscale_banded_uint8(model, image, args.exact_bands) if args.exact_bands > 0 else None
            if banded is not None:
                output = banded
            elif hasattr(model, 'upscale_uint8'):   # round/clip on the device, 4x smaller device->host copy
                output = model.upscale_uint8(input_list=[image], scale=args.scale)[0]
            else:
                output = _image_to_uint8(model.upscale(input_list=[image], scale=args.scale)[0])
            durations.append(time.perf_counter() - start)
            _save_image(output, os.path.join(args.output_path, name))
            print('%d/%d, %s, duration=%.4f' % (i + 1, len(names), name, durations[-1]))
    if durations:
        print('average duration=%.4f' % float(np.mean(durations)))
    print('finished')


def runtime_main(argv=None):
    """reference runtime.py:57-73: per validation image, time `model.fwd_runtime` between device synchronisations.
    Needs no checkpoint (random-init weights), like the reference."""
    import torch

    args, rest = _flag_parser(COMMON_FLAGS, SCALES_FLAG, RUNTIME_FLAGS).parse_known_args(argv)
    _select_device(args.cuda_device)
    scale_list = [int(s) for s in args.scales.split(',')]
    dataloader, _, rest = _open_loader(args.dataloader, scale_list, rest)
    model, _ = _open_model(args.model, rest, dict(is_training=False, scales=scale_list))

    print('begin runtime check')
    averages = {}
    for scale in scale_list:
        runtimes = []
        count = dataloader.get_num_images()
        with torch.no_grad():
            for image_index in range(count):
                input_image, _, _ = dataloader.get_image_pair(image_index=image_index, scale=scale)
                input_tensor = torch.as_tensor(np.asarray([input_image]), dtype=torch.float32, device='cuda')
                torch.cuda.synchronize()
                start = time.perf_counter()
                model.fwd_runtime(input_tensor=input_tensor)
                torch.cuda.synchronize()
                runtimes.append(time.perf_counter() - start)
                print(f'{image_index+1}/{count}, runtime={runtimes[-1]:.4f}')
        averages[scale] = float(np.mean(runtimes))
        print(f'runtime={averages[scale]:.4f}')
    print('finished')
    return averages


def _summary_writer(path):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir=path)
    except Exception:  # noqa: BLE001  (tensorboard is optional)
        return None


def train_main(argv=None, epoch_bookkeeping=False):
    """reference train_larva.py:20-140 (and train_larvaV2.py's epoch bookkeeping with `epoch_bookkeeping=True`)."""
    import torch

    tables = [COMMON_FLAGS, SCALES_FLAG, TRAIN_FLAGS]
    if epoch_bookkeeping:
        tables.append([('steps_per_epoch', int, 1000, 'Steps per epoch (V2 trainer).')])
    args, rest = _flag_parser(*tables).parse_known_args(argv)
    _select_device(args.cuda_device)
    scale_list = [int(s) for s in args.scales.split(',')]
    os.makedirs(args.train_path, exist_ok=True)

    # Data parallel (new; the reference is single-process): under torchrun every rank owns one GPU and a contiguous shard
    # of `--batch_size` patches; the engine all-reduces the gradient arena so that the update equals the reference's
    # single-process step on the global batch (models/LarvaNet.py:102-114).  Without torchrun this is a no-op.
    from larvanet_b200 import dist as lvdist
    rank, world, local = lvdist.init_from_env()
    if world > 1:
        torch.cuda.set_device(local)
        if args.batch_size < world:
            raise ValueError(f'--batch_size={args.batch_size} cannot be sharded over WORLD_SIZE={world} ranks')
    shard_begin, shard_end = lvdist.shard_range(args.batch_size, rank, world)
    local_batch = shard_end - shard_begin
    log = print if rank == 0 else (lambda *a, **k: None)

    dataloader, dataloader_args, rest = _open_loader(args.dataloader, scale_list, rest)
    val_dataloader, _, _ = _open_loader(args.val_dataloader, scale_list)
    if world > 1 and hasattr(dataloader, 'rs'):
        # decorrelate the ranks' random crops (each rank draws its own shard of the global batch)
        dataloader.rs = np.random.RandomState(int(getattr(dataloader_args, 'synthetic_seed', 0)) + 12345 + 7919 * rank)

    def before_prepare(m):
        m.volume_per_step = (args.input_patch_size ** 2) * args.batch_size * 3
        if epoch_bookkeeping:
            m.steps_per_epoch = args.steps_per_epoch

    model, model_args = _open_model(args.model, rest, dict(is_training=True, scales=scale_list, global_step=args.global_step),
                                    before_prepare=before_prepare, restore=(args.restore_path, args.restore_target))

    if world > 1:
        # refuse to train N silent replicas: the plugin must be able to shard the step
        engine = model._engine() if hasattr(model, '_engine') else None
        if engine is None or not hasattr(engine, 'set_data_parallel'):
            raise RuntimeError(f'--model={args.model} cannot run data-parallel (WORLD_SIZE={world}); launch it without torchrun')
        engine.set_data_parallel(world)
        model.dp_rank = rank
        log(f'data parallel: {world} ranks x {local_batch} patches (global batch {args.batch_size}), gradient exchange: '
            f'{"symmetric-memory peer kernel" if engine._symm is not None else "NCCL all-reduce"}')

    summary_writers = {s: (_summary_writer(os.path.join(args.train_path, 'x%d' % s)) if rank == 0 else None) for s in scale_list}
    if rank == 0:
        with open(os.path.join(args.train_path, 'arguments.json'), 'w') as f:
            f.write(json.dumps({**vars(args), **vars(dataloader_args), **vars(model_args)}, sort_keys=True, indent=2))

    if dataloader.is_threaded:
        dataloader.start_training_queue_runner(batch_size=local_batch, input_patch_size=args.input_patch_size)

    log('begin training')
    log(f'volume {model.volume_per_step/1e6:.2f}M for 1 step.')
    log(f'needs {model_args.val_volume/model.volume_per_step:.0f}steps to validate for {model_args.val_volume/1e9:.1f}G volume.')
    loss = float('nan')
    from larvanet_b200.prefetch import DevicePrefetcher

    def host_batches():
        # same loader calls as the reference loop (train_larva.py:112-121 there); batches are staged in pinned memory
        while True:
            sc = model.get_next_train_scale()
            if dataloader.is_threaded:
                input_list, truth_list = dataloader.get_queue_data(scale=sc)
            else:
                input_list, truth_list = dataloader.get_patch_batch(batch_size=local_batch, scale=sc,
                                                                    input_patch_size=args.input_patch_size)
            if torch.is_tensor(input_list):      # tensor loaders (reference div2k_train_loader_tensor.py) hand out tensors
                yield input_list, truth_list
            else:
                yield (torch.from_numpy(np.asarray(input_list, dtype=np.float32)).pin_memory(),
                       torch.from_numpy(np.asarray(truth_list, dtype=np.float32)).pin_memory())

    if getattr(dataloader, 'device', None) is not None and getattr(dataloader, 'device').type == 'cuda':
        feeder = host_batches()   # GPU-resident loader (device-side crop / rot90 / flip): batches are born on the device
    else:
        feeder = DevicePrefetcher(host_batches(), model.device, depth=2, defer=True)
    try:
        while model.global_step < args.max_steps:
            scale = model.get_next_train_scale()
            summary = summary_writers[scale] if (model.global_step % args.summary_freq == 0) else None
            start_time = time.time()
            # the next batch's host->device copy was issued on the copy stream while the previous step computed
            input_tensor, truth_tensor = next(feeder)
            dataload_time = time.time() - start_time
            np2ts_time = 0.0
            check_time = time.time()
            loss = model.train_step_larva(args=args, val_dataloader=val_dataloader, input_tensor=input_tensor,
                                          truth_tensor=truth_tensor, summary=summary)
            train_time = time.time() - check_time
            duration = time.time() - start_time
            if args.sleep_ratio > 0 and duration > 0:
                time.sleep(min(10.0, duration * args.sleep_ratio))
            if rank == 0 and model.global_step < 1000 and model.global_step % args.log_freq == 0:
                print('step %d, lr %.10f, loss %.6f (%.3f sec/batch)' % (model.global_step, model.get_lr(), loss, duration))
                print(f'dataload_time:{dataload_time:.4f}s, np2ts_time:{np2ts_time:.4f}s, train_time: {train_time:.4f}s')
    except KeyboardInterrupt:
        print('interrupted (KeyboardInterrupt)')

    log('finished')
    for w in summary_writers.values():
        if w is not None:
            w.close()
    if dataloader.is_threaded:
        dataloader.stop_queue_runners()
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    return loss
