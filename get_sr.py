"""Inference entry point: directory of PNGs -> x4 super-resolved PNGs (drop-in for reference get_sr.py).

    python get_sr.py --model=LarvaNet --num_modules=4 --num_blocks=4,4,4,4 --restore_path=ckpt.pth \
        --input_path=lr/ --output_path=sr/
"""
import argparse
import importlib
import os
import time

import numpy as np


def main(argv=None):
    import cv2 as cv
    import torch

    parser = argparse.ArgumentParser()
    parser.add_argument('--model', type=str, default='LarvaNet', help='Name of the model.')
    parser.add_argument('--scale', type=int, default=4, help='Upscaling factor.')
    parser.add_argument('--cuda_device', type=str, default='0', help='Value for CUDA_VISIBLE_DEVICES.')
    parser.add_argument('--restore_path', type=str, required=True, help='Checkpoint to use.')
    parser.add_argument('--restore_target', type=str, help='Target of the restoration.')
    parser.add_argument('--restore_global_step', type=int, default=0, help='Global step of the checkpoint.')
    parser.add_argument('--input_path', type=str, default='LR', help='Directory of input PNGs.')
    parser.add_argument('--output_path', type=str, default='SR', help='Directory for the upscaled PNGs.')
    args, remaining_args = parser.parse_known_args(argv)

    os.environ['CUDA_VISIBLE_DEVICES'] = args.cuda_device
    os.makedirs(args.output_path, exist_ok=True)

    print('prepare model - %s' % args.model)
    model = importlib.import_module('models.' + args.model).create_model()
    _, remaining_args = model.parse_args(remaining_args)
    model.prepare(is_training=False, scales=[args.scale], global_step=args.restore_global_step)
    if remaining_args:
        print('WARNING: found unhandled arguments: %s' % remaining_args)
    model.restore(ckpt_path=args.restore_path, target=args.restore_target)
    print('restored the model')

    names = sorted(f for f in os.listdir(args.input_path) if f.lower().endswith('.png'))
    durations = []
    with torch.no_grad():
        for i, name in enumerate(names):
            image = cv.cvtColor(cv.imread(os.path.join(args.input_path, name)), cv.COLOR_BGR2RGB)
            image = np.transpose(image, [2, 0, 1])
            start = time.perf_counter()
            if hasattr(model, 'upscale_uint8'):   # round/clip on the device, 4x smaller device->host copy
                output = model.upscale_uint8(input_list=[image], scale=args.scale)[0]
            else:
                output = np.clip(np.round(model.upscale(input_list=[image], scale=args.scale)[0]), 0, 255).astype(np.uint8)
            durations.append(time.perf_counter() - start)
            cv.imwrite(os.path.join(args.output_path, name), cv.cvtColor(np.transpose(output, [1, 2, 0]), cv.COLOR_RGB2BGR))
            print('%d/%d, %s, duration=%.4f' % (i + 1, len(names), name, durations[-1]))
    if durations:
        print('average duration=%.4f' % float(np.mean(durations)))
    print('finished')


if __name__ == '__main__':
    main()
