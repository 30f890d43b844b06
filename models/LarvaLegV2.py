"""LarvaLegV2 model plugin -- drop-in for reference models/LarvaLegV2.py: LarvaNetV2 with `--leg=k` early-exit inference.

Inference ignores the tail: `leg == 0` returns the bicubic base, otherwise bodies 0..k-1 run and body k-1's leg
reconstructs the image (reference models/LarvaLegV2.py:357-368).  Same state_dict and training step as LarvaNetV2.
"""
from models.LarvaNetV2 import (LarvaBody, LarvaHead, LarvaLeg, LarvaNetModule as _Module, LarvaNetV2 as _LarvaNetV2,  # noqa: F401
                               LarvaTail, ResidualBlock, initialize_weights)


def create_model():
    return LarvaNet()


class LarvaNetModule(_Module):
    EARLY_EXIT = True


class LarvaNet(_LarvaNetV2):
    MODULE = LarvaNetModule
    HAS_LEG = True
