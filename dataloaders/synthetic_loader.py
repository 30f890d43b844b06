"""Synthetic training/validation loader: lets train_larva.py / validate.py / runtime.py run with no dataset.

The reference's loaders read DIV2K PNGs from hard-coded Windows paths (dataloaders/div2k_val_loader.py:28,108,125) and
are out of scope (host I/O).  This plugin implements the same contract (dataloaders/base.py) on band-limited random
images from `larvanet_b200.synth`, with the reference's patch augmentation (random crop, rot90, horizontal flip --
dataloaders/div2k_train_loader.py:72-98).
"""
import argparse
import copy

import numpy as np

from dataloaders.base import BaseLoader
from larvanet_b200 import synth


def create_loader():
    return SyntheticLoader()


class SyntheticLoader(BaseLoader):
    def __init__(self):
        super().__init__()
        self.args = argparse.Namespace(synthetic_images=4, synthetic_height=96, synthetic_width=128, synthetic_seed=1)

    def parse_args(self, args):
        parser = argparse.ArgumentParser()
        parser.add_argument('--synthetic_images', type=int, default=4, help='Number of synthetic images.')
        parser.add_argument('--synthetic_height', type=int, default=96, help='LR height of each synthetic image.')
        parser.add_argument('--synthetic_width', type=int, default=128, help='LR width of each synthetic image.')
        parser.add_argument('--synthetic_seed', type=int, default=1, help='Seed of the image generator.')
        self.args, remaining = parser.parse_known_args(args=args)
        return copy.deepcopy(self.args), remaining

    def prepare(self, scales):
        self.scale_list = scales
        a = self.args
        self.lr, self.hr = {}, {}
        for scale in scales:
            lr, hr = synth.make_smooth_images(a.synthetic_images, a.synthetic_height, a.synthetic_width, scale=scale,
                                              seed=a.synthetic_seed)
            self.lr[scale], self.hr[scale] = np.round(lr), np.round(hr)   # what a PNG would hold
        self.rs = np.random.RandomState(a.synthetic_seed + 12345)

    def get_num_images(self):
        return self.args.synthetic_images

    def get_patch_batch(self, batch_size, scale, input_patch_size):
        pairs = [self.get_random_image_patch_pair(scale, input_patch_size) for _ in range(batch_size)]
        return [p[0] for p in pairs], [p[1] for p in pairs]

    def get_random_image_patch_pair(self, scale, input_patch_size):
        return self.get_image_patch_pair(self.rs.randint(self.get_num_images()), scale, input_patch_size)

    def get_image_patch_pair(self, image_index, scale, input_patch_size):
        lr, hr = self.lr[scale][image_index], self.hr[scale][image_index]
        _, h, w = lr.shape
        y = self.rs.randint(h - input_patch_size + 1)
        x = self.rs.randint(w - input_patch_size + 1)
        p = lr[:, y:y + input_patch_size, x:x + input_patch_size]
        t = hr[:, y * scale:(y + input_patch_size) * scale, x * scale:(x + input_patch_size) * scale]
        k = self.rs.randint(4)
        p, t = np.rot90(p, k, axes=(1, 2)), np.rot90(t, k, axes=(1, 2))
        if self.rs.randint(2):
            p, t = p[:, :, ::-1], t[:, :, ::-1]
        return np.ascontiguousarray(p), np.ascontiguousarray(t)

    def get_image_pair(self, image_index, scale):
        return self.lr[scale][image_index], self.hr[scale][image_index], 'synthetic_%04d' % image_index
