"""EDSR-baseline model plugin -- drop-in for reference models/edsr.py (inference) on the B200 conv kernels.

Same flags (`--edsr_conv_features`, `--edsr_res_blocks`, `--edsr_res_weight`, learning-rate flags), same sub-module
names and therefore the same `state_dict()` keys (`mean_shift`, `first_conv`, `res_blocks.{j}.body.{0,2}`,
`after_res_conv`, `upsample.body.{0,2}`, `final_conv`, `mean_inverse_shift`).  Note the reference quirk kept here:
`MeanShift.__init__` stores its identity/mean tensors in unused attributes, so the two 1x1 convs keep their frozen
default-initialised weights, which live in the checkpoint (reference models/edsr.py:129-136) -- they are run as general
1x1 convs (fused into the first and last kernels), not as +-mean.

BASELINE.json lists EDSR for inference only (config 3); `train_step` is not provided on this path.
"""
import argparse
import copy
import math
import os

import numpy as np
import torch
import torch.nn as nn

from models.base import BaseModel
from models.LarvaNet import Conv3x3


def create_model():
    return EDSR()


class MeanShift(nn.Module):
    """Frozen 1x1 conv parameter holder (weight [3,3,1,1], bias [3]) -- reference models/edsr.py:129-136."""

    def __init__(self, rgb_mean, sign):
        super().__init__()
        bound = 1.0 / math.sqrt(3)
        self.weight = nn.Parameter(torch.empty(3, 3, 1, 1).uniform_(-bound, bound), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(3).uniform_(-bound, bound), requires_grad=False)
        self.weight_data = torch.eye(3).view(3, 3, 1, 1)           # unused, like the reference
        self.bias_data = sign * torch.Tensor(rgb_mean)


class ResidualBlock(nn.Module):
    def __init__(self, num_channels, weight=1.0):
        super().__init__()
        self.body = nn.Sequential(Conv3x3(num_channels, num_channels), nn.ReLU(inplace=True),
                                  Conv3x3(num_channels, num_channels))
        self.weight = weight


class UpsampleBlock(nn.Module):
    def __init__(self, num_channels, scale):
        super().__init__()
        layers = []
        if scale in (2, 4, 8):
            for _ in range(int(math.log(scale, 2))):
                layers.append(Conv3x3(num_channels, 4 * num_channels))
                layers.append(nn.PixelShuffle(2))
        elif scale == 3:
            layers.append(Conv3x3(num_channels, 9 * num_channels))
            layers.append(nn.PixelShuffle(3))
        self.body = nn.Sequential(*layers)


class EDSRModule(nn.Module):
    def __init__(self, args, scale):
        super().__init__()
        f = args.edsr_conv_features
        self.features, self.num_res_blocks, self.res_weight, self.scale = f, args.edsr_res_blocks, args.edsr_res_weight, scale
        self.mean_shift = MeanShift([114.4, 111.5, 103.0], sign=1.0)
        self.first_conv = Conv3x3(3, f)
        self.res_blocks = nn.Sequential(*[ResidualBlock(f, weight=args.edsr_res_weight) for _ in range(args.edsr_res_blocks)])
        self.after_res_conv = Conv3x3(f, f)
        self.upsample = UpsampleBlock(f, scale)
        self.final_conv = Conv3x3(f, 3)
        self.mean_inverse_shift = MeanShift([114.4, 111.5, 103.0], sign=-1.0)
        self.precision = getattr(args, 'precision', 'bf16')
        self._engine = None

    def engine(self):
        if self._engine is None:
            from larvanet_b200._lib import LarvaNetB200Error
            from larvanet_b200.engine import EdsrEngine
            if not torch.cuda.is_available():
                raise LarvaNetB200Error('EDSR on larvanet_b200 needs a CUDA device (sm_100a); there is no CPU path')
            dt = {'bf16': torch.bfloat16, 'fp32': torch.float32}[self.precision]
            self._engine = EdsrEngine(self, self.features, self.num_res_blocks, self.res_weight, self.scale, act_dtype=dt)
        return self._engine

    def _apply(self, fn, *a, **k):
        if self._engine is not None:
            probe = fn(torch.empty(0, device=self._engine.device))
            if probe.device == self._engine.device and probe.dtype == torch.float32:
                return self
            from larvanet_b200._lib import LarvaNetB200Error
            raise LarvaNetB200Error('cannot move/cast an EDSRModule after its engine has been created')
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        return self.engine().forward(x).clone()


class EDSR(BaseModel):
    def parse_args(self, args):
        parser = argparse.ArgumentParser()
        parser.add_argument('--edsr_conv_features', type=int, default=64, help='The number of convolutional features.')
        parser.add_argument('--edsr_res_blocks', type=int, default=16, help='The number of residual blocks.')
        parser.add_argument('--edsr_res_weight', type=float, default=1.0, help='The scaling factor.')
        parser.add_argument('--edsr_learning_rate', type=float, default=1e-4, help='Initial learning rate.')
        parser.add_argument('--edsr_learning_rate_decay', type=float, default=0.5, help='Learning rate decay factor.')
        parser.add_argument('--edsr_learning_rate_decay_steps', type=int, default=200000, help='Decay period in steps.')
        parser.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'fp32'],
                            help='bf16: tcgen05 tensor-core path; fp32: CUDA-core validation mode.')
        self.args, remaining_args = parser.parse_known_args(args=args)
        return copy.deepcopy(self.args), remaining_args

    def prepare(self, is_training, scales, global_step=0):
        self.global_step = global_step
        self.scale_list = scales
        for scale in self.scale_list:
            if scale not in (2, 3, 4):
                raise ValueError('Unsupported scale is provided.')
        if len(self.scale_list) != 1:
            raise ValueError('Only one scale should be provided.')
        self.scale = self.scale_list[0]
        self.model = EDSRModule(args=self.args, scale=self.scale)
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.model = self.model.to(self.device)

    def save(self, base_path):
        save_path = os.path.join(base_path, 'model_%d.pth' % self.global_step)
        torch.save({k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}, save_path)

    def restore(self, ckpt_path, target=None):
        self.model.load_state_dict(torch.load(ckpt_path, map_location=self.device))

    def get_model(self):
        return self.model

    def get_next_train_scale(self):
        return self.scale_list[np.random.randint(len(self.scale_list))]

    def train_step(self, input_list, scale, truth_list, summary=None):
        raise NotImplementedError('EDSR training is outside the larvanet_b200 hot path (BASELINE.json config 3 is '
                                  'inference only); use --model=LarvaNet / LarvaNetV2 with train_larva.py')

    def upscale(self, input_list, scale):
        x = input_list if torch.is_tensor(input_list) else torch.as_tensor(np.asarray(input_list), dtype=torch.float32)
        return self.model(x.to(device=self.device, dtype=torch.float32)).detach().cpu().numpy()

    def fwd_runtime(self, input_tensor):
        return self.model(input_tensor)

    def _get_learning_rate(self):
        return self.args.edsr_learning_rate * (
            self.args.edsr_learning_rate_decay ** (self.global_step // self.args.edsr_learning_rate_decay_steps))
