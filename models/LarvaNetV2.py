"""LarvaNetV2 model plugin -- drop-in for reference models/LarvaNetV2.py on B200 kernels.

V2 = the LarvaNet trunk plus `LarvaTail`: cat(all body outputs) -> 3x3 merge conv (48*M -> 48) -> conv-ReLU-conv ->
PixelShuffle(4) -> + bicubic base (reference models/LarvaNetV2.py:314-334).  Inference uses the tail only (:355-365);
training sums the M leg losses and the tail loss and divides by M+1 (:105-118).  The merge conv never materialises the
concatenation: the conv kernel loops its K dimension over the M source tensors.
"""
import torch
import torch.nn as nn

from models.LarvaNet import (NUM_FILTERS, Conv3x3, LarvaBody, LarvaHead, LarvaLeg, LarvaNet, LarvaNetModule as _V1Module,
                             ResidualBlock, _engine_of, initialize_weights)

__all__ = ['create_model', 'LarvaNetV2', 'LarvaNetModule', 'LarvaTail', 'LarvaBody', 'LarvaHead', 'LarvaLeg',
           'ResidualBlock', 'initialize_weights']


def create_model():
    return LarvaNetV2()


class LarvaTail(nn.Module):
    def __init__(self, num_modules):
        super().__init__()
        self.merge_conv = Conv3x3(NUM_FILTERS * num_modules, NUM_FILTERS)
        self.recon_block = nn.Sequential(Conv3x3(NUM_FILTERS, NUM_FILTERS), nn.ReLU(inplace=True),
                                         Conv3x3(NUM_FILTERS, NUM_FILTERS))
        initialize_weights([self.recon_block, self.merge_conv], 0.1)
        self.upsample = nn.PixelShuffle(4)

    def forward(self, features, base):
        return _engine_of(self).run_tail(features, base)


class LarvaNetModule(_V1Module):
    V2 = True

    def _build_extra(self):
        self.tail = LarvaTail(self.len)


class LarvaNetV2(LarvaNet):
    MODULE = LarvaNetModule

    # reference models/LarvaNetV2.py:196-206: keep only keys the model has, so V1 checkpoints warm-start V2
    def restore(self, ckpt_path, target=None):
        pretrained = torch.load(ckpt_path, map_location=self.device)
        model_dict = self.model.state_dict()
        model_dict.update({k: v for k, v in pretrained.items() if k in model_dict})
        self.model.load_state_dict(model_dict)
        self.model.to(self.device)

    # reference models/LarvaNetV2.py:47-66: other defaults than V1, no --lr_step / --cooldown (scheduler cooldown 0)
    DEFAULTS = dict(val_volume=3e9, lr=1e-4, min_lr=1e-7)
    HAS_COOLDOWN = False
