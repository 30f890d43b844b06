"""LarvaLeg model plugin -- drop-in for reference models/LarvaLeg.py: LarvaNet with `--leg=k` early-exit inference.

`leg == 0` returns the bicubic base only; otherwise bodies 0..k-1 run and body k-1's leg reconstructs the image
(reference models/LarvaLeg.py:289-299).  Same state_dict as LarvaNet, so LarvaNet checkpoints load unchanged; the
training step is the multi-exit step of LarvaNet (the reference file differs only in the flags and the forward).
The shorter chain runs on the same kernels (`LarvaEngine.forward(exit_leg=k)`).
"""
from models.LarvaNet import (NUM_FILTERS, Conv3x3, LarvaBody, LarvaHead, LarvaLeg, LarvaNet as _LarvaNet,  # noqa: F401
                             LarvaNetModule as _Module, ResidualBlock, initialize_weights)


def create_model():
    return LarvaNet()


class LarvaNetModule(_Module):
    EARLY_EXIT = True


class LarvaNet(_LarvaNet):
    MODULE = LarvaNetModule
    # reference models/LarvaLeg.py:47-64
    DEFAULTS = dict(val_volume=3e9, lr=1e-4, min_lr=1e-7)
    HAS_COOLDOWN = False
    HAS_LEG = True
