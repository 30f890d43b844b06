"""GPU runtime check (drop-in for reference runtime.py:57-73): per validation image, time `model.fwd_runtime` between
`torch.cuda.synchronize()` calls.  Needs no checkpoint (random-init weights), like the reference."""
import argparse
import importlib
import os
import time

import numpy as np


def main(argv=None):
    import torch

    parser = argparse.ArgumentParser()
    parser.add_argument('--dataloader', type=str, default='synthetic_val_loader', help='Name of the data loader.')
    parser.add_argument('--model', type=str, default='LarvaNet', help='Name of the model.')
    parser.add_argument('--scales', type=str, default='4', help='Comma separated scales.')
    parser.add_argument('--cuda_device', type=str, default='0', help='Value for CUDA_VISIBLE_DEVICES.')
    args, remaining_args = parser.parse_known_args(argv)

    os.environ['CUDA_VISIBLE_DEVICES'] = args.cuda_device
    scale_list = [int(s) for s in args.scales.split(',')]

    print('prepare data loader - %s' % args.dataloader)
    dataloader = importlib.import_module('dataloaders.' + args.dataloader).create_loader()
    _, remaining_args = dataloader.parse_args(remaining_args)
    dataloader.prepare(scales=scale_list)

    print('prepare model - %s' % args.model)
    model = importlib.import_module('models.' + args.model).create_model()
    _, remaining_args = model.parse_args(remaining_args)
    model.prepare(is_training=False, scales=scale_list)
    if remaining_args:
        print('WARNING: found unhandled arguments: %s' % remaining_args)

    print('begin runtime check')
    averages = {}
    for scale in scale_list:
        runtimes = []
        with torch.no_grad():
            for image_index in range(dataloader.get_num_images()):
                input_image, _, _ = dataloader.get_image_pair(image_index=image_index, scale=scale)
                input_tensor = torch.as_tensor(np.asarray([input_image]), dtype=torch.float32, device='cuda')
                torch.cuda.synchronize()
                start = time.perf_counter()
                model.fwd_runtime(input_tensor=input_tensor)
                torch.cuda.synchronize()
                runtimes.append(time.perf_counter() - start)
                print(f'{image_index+1}/{dataloader.get_num_images()}, runtime={runtimes[-1]:.4f}')
        averages[scale] = float(np.mean(runtimes))
        print(f'runtime={averages[scale]:.4f}')
    print('finished')
    return averages


if __name__ == '__main__':
    main()
