"""Training entry point for LarvaNet (drop-in for reference train_larva.py; `train_larvaV2.py` adds epoch bookkeeping).

    python train_larva.py --model=LarvaNet --num_modules=4 --num_blocks=4,4,4,4 --batch_size=16 --input_patch_size=48 \
        --train_path=/tmp/larva --max_steps=100 --sleep_ratio=0 [--dataloader=synthetic_loader]

Same flag chain as the reference: this parser takes what it knows, the data loader parses the remainder, the model
parses what is left (reference train_larva.py:49,60,71).  Differences, all on the host side: the default loaders are the
synthetic ones (DIV2K readers are out of scope), `--max_steps` is honoured (the reference loops until Ctrl-C), and
TensorBoard is optional.
"""
import argparse
import importlib
import json
import os
import time

import torch


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--dataloader', type=str, default='synthetic_loader', help='Name of the data loader.')
    parser.add_argument('--val_dataloader', type=str, default='synthetic_val_loader', help='Name of the validation loader.')
    parser.add_argument('--model', type=str, default='LarvaNet', help='Name of the model.')
    parser.add_argument('--batch_size', type=int, default=16, help='Patches per training step.')
    parser.add_argument('--input_patch_size', type=int, default=48, help='LR patch edge.')
    parser.add_argument('--scales', type=str, default='4', help='Comma separated scales.')
    parser.add_argument('--cuda_device', type=str, default='0', help='Value for CUDA_VISIBLE_DEVICES.')
    parser.add_argument('--train_path', type=str, default='/tmp/larvanet_b200/train/', help='Checkpoint/summary directory.')
    parser.add_argument('--max_steps', type=int, default=300000, help='Stop after this many steps.')
    parser.add_argument('--log_freq', type=int, default=10, help='Print every N steps.')
    parser.add_argument('--summary_freq', type=int, default=1000, help='TensorBoard summary period.')
    parser.add_argument('--save_freq', type=int, default=10000, help='Checkpoint period (the model also saves by volume).')
    parser.add_argument('--sleep_ratio', type=float, default=0.05, help='Idle fraction per step (0 disables).')
    parser.add_argument('--restore_path', type=str, help='Checkpoint to start from.')
    parser.add_argument('--restore_target', type=str, help='Target of the restoration.')
    parser.add_argument('--global_step', type=int, default=0, help='Initial global step when resuming.')
    return parser


def _summary_writer(path):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir=path)
    except Exception:  # noqa: BLE001  (tensorboard is optional)
        return None


def main(argv=None, epoch_bookkeeping=False):
    parser = build_parser()
    if epoch_bookkeeping:
        parser.add_argument('--steps_per_epoch', type=int, default=1000, help='Steps per epoch (V2 trainer).')
    args, remaining_args = parser.parse_known_args(argv)

    if 'LOCAL_RANK' not in os.environ:
        os.environ['CUDA_VISIBLE_DEVICES'] = args.cuda_device
    scale_list = [int(s) for s in args.scales.split(',')]
    os.makedirs(args.train_path, exist_ok=True)

    print('prepare data loader - %s' % args.dataloader)
    dataloader = importlib.import_module('dataloaders.' + args.dataloader).create_loader()
    dataloader_args, remaining_args = dataloader.parse_args(remaining_args)
    dataloader.prepare(scales=scale_list)
    val_dataloader = importlib.import_module('dataloaders.' + args.val_dataloader).create_loader()
    val_dataloader.prepare(scales=scale_list)

    print('prepare model - %s' % args.model)
    model = importlib.import_module('models.' + args.model).create_model()
    model_args, remaining_args = model.parse_args(remaining_args)
    model.volume_per_step = (args.input_patch_size ** 2) * args.batch_size * 3
    if epoch_bookkeeping:
        model.steps_per_epoch = args.steps_per_epoch
    model.prepare(is_training=True, scales=scale_list, global_step=args.global_step)
    if remaining_args:
        print('WARNING: found unhandled arguments: %s' % remaining_args)
    if args.restore_path is not None:
        model.restore(ckpt_path=args.restore_path, target=args.restore_target)
        print('restored the model')

    summary_writers = {s: _summary_writer(os.path.join(args.train_path, 'x%d' % s)) for s in scale_list}
    with open(os.path.join(args.train_path, 'arguments.json'), 'w') as f:
        f.write(json.dumps({**vars(args), **vars(dataloader_args), **vars(model_args)}, sort_keys=True, indent=2))

    if dataloader.is_threaded:
        dataloader.start_training_queue_runner(batch_size=args.batch_size, input_patch_size=args.input_patch_size)

    print('begin training')
    print(f'volume {model.volume_per_step/1e6:.2f}M for 1 step.')
    print(f'needs {model_args.val_volume/model.volume_per_step:.0f}steps to validate for {model_args.val_volume/1e9:.1f}G volume.')
    loss = float('nan')
    import numpy as np
    from larvanet_b200.prefetch import DevicePrefetcher

    def host_batches():
        # same loader calls as the reference loop (train_larva.py:112-121 there); batches are staged in pinned memory
        while True:
            sc = model.get_next_train_scale()
            if dataloader.is_threaded:
                input_list, truth_list = dataloader.get_queue_data(scale=sc)
            else:
                input_list, truth_list = dataloader.get_patch_batch(batch_size=args.batch_size, scale=sc,
                                                                    input_patch_size=args.input_patch_size)
            yield (torch.from_numpy(np.asarray(input_list, dtype=np.float32)).pin_memory(),
                   torch.from_numpy(np.asarray(truth_list, dtype=np.float32)).pin_memory())

    feeder = DevicePrefetcher(host_batches(), model.device, depth=2)
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
            if model.global_step < 1000 and model.global_step % args.log_freq == 0:
                print('step %d, lr %.10f, loss %.6f (%.3f sec/batch)' % (model.global_step, model.get_lr(), loss, duration))
                print(f'dataload_time:{dataload_time:.4f}s, np2ts_time:{np2ts_time:.4f}s, train_time: {train_time:.4f}s')
    except KeyboardInterrupt:
        print('interrupted (KeyboardInterrupt)')

    print('finished')
    for w in summary_writers.values():
        if w is not None:
            w.close()
    if dataloader.is_threaded:
        dataloader.stop_queue_runners()
    return loss


if __name__ == '__main__':
    main()
