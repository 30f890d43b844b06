"""Validation entry point + the PSNR helpers the model plugins import (drop-in for reference validate.py).

    python validate.py --model=LarvaNet --num_modules=4 --num_blocks=4,4,4,4 --restore_path=ckpt.pth \
        [--dataloader=synthetic_val_loader] [--save_path=out/] [--chop_forward]
"""
import argparse
import importlib
import os
import time

import numpy as np


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


def _save_image(image, path):
    import cv2 as cv
    cv.imwrite(path, cv.cvtColor(np.transpose(image, [1, 2, 0]), cv.COLOR_RGB2BGR))


def main(argv=None):
    import torch
    from utils import image_utils

    parser = argparse.ArgumentParser()
    parser.add_argument('--dataloader', type=str, default='synthetic_val_loader', help='Name of the data loader.')
    parser.add_argument('--model', type=str, default='LarvaNet', help='Name of the model.')
    parser.add_argument('--scales', type=str, default='4', help='Comma separated scales.')
    parser.add_argument('--cuda_device', type=str, default='0', help='Value for CUDA_VISIBLE_DEVICES.')
    parser.add_argument('--restore_path', type=str, required=True, help='Checkpoint to evaluate.')
    parser.add_argument('--restore_target', type=str, help='Target of the restoration.')
    parser.add_argument('--restore_global_step', type=int, default=0, help='Global step of the checkpoint.')
    parser.add_argument('--save_path', type=str, help='Write the upscaled PNGs below this directory.')
    parser.add_argument('--chop_forward', action='store_true', help='Upscale in four overlapping quadrants.')
    parser.add_argument('--chop_overlap_size', type=int, default=20, help='Quadrant overlap in LR pixels (even).')
    args, remaining_args = parser.parse_known_args(argv)

    os.environ['CUDA_VISIBLE_DEVICES'] = args.cuda_device
    scale_list = [int(s) for s in args.scales.split(',')]

    print('prepare data loader - %s' % args.dataloader)
    dataloader = importlib.import_module('dataloaders.' + args.dataloader).create_loader()
    dataloader.prepare(scales=scale_list)

    print('prepare model - %s' % args.model)
    model = importlib.import_module('models.' + args.model).create_model()
    _, remaining_args = model.parse_args(remaining_args)
    model.prepare(is_training=False, scales=scale_list, global_step=args.restore_global_step)
    if remaining_args:
        print('WARNING: found unhandled arguments: %s' % remaining_args)
    model.restore(ckpt_path=args.restore_path, target=args.restore_target)
    print('restored the model')

    print('begin validation')
    results = {}
    for scale in scale_list:
        durations, psnrs = [], []
        with torch.no_grad():
            for image_index in range(dataloader.get_num_images()):
                input_image, truth_image, image_name = dataloader.get_image_pair(image_index=image_index, scale=scale)
                start = time.perf_counter()
                if args.chop_forward:
                    output_image = image_utils.upscale_with_chop_forward(model=model, input_image=input_image, scale=scale,
                                                                         overlap_size=args.chop_overlap_size)
                elif hasattr(model, 'upscale_uint8'):   # round/clip on the device (== _image_to_uint8 below)
                    output_image = model.upscale_uint8(input_list=[input_image], scale=scale)[0]
                else:
                    output_image = model.upscale(input_list=[input_image], scale=scale)[0]
                durations.append(time.perf_counter() - start)
                truth_image = _image_to_uint8(truth_image)
                output_image = _image_to_uint8(output_image)
                if args.save_path is not None:
                    os.makedirs(os.path.join(args.save_path, 'x%d' % scale), exist_ok=True)
                    _save_image(output_image, os.path.join(args.save_path, 'x%d' % scale, image_name + '.png'))
                truth_image = _fit_truth_image_size(output_image=output_image, truth_image=truth_image)
                psnrs.append(_image_psnr(output_image=output_image, truth_image=truth_image))
                print('x%d, %d/%d, psnr=%.2f, duration=%.4f' % (scale, image_index + 1, dataloader.get_num_images(),
                                                                psnrs[-1], durations[-1]))
        results[scale] = (float(np.mean(psnrs)), float(np.mean(durations)))
        print('x%d, psnr=%.2f, duration=%.4f' % (scale, results[scale][0], results[scale][1]))
    print('finished')
    return results


if __name__ == '__main__':
    main()
