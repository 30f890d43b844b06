"""Seeded synthetic weights and images for parity tests, smoke and benchmarks.

There is no dataset or checkpoint for this path, so every test and
benchmark uses weights/inputs produced here.  The generator is numpy's frozen legacy
`RandomState` (bit-stable across numpy versions and machines), so the small fixtures under
`tests/golden/` only need to store a seed + config, not the weights.

Weight statistics follow the reference initialisers:
  * LarvaNet / LarvaNetV2: `initialize_weights(.., 0.1)` = kaiming-normal(fan_in) * 0.1, zero
    bias (reference models/LarvaNet.py:22-39).  `bias_std > 0` replaces the zero bias by small
    normal noise so parity tests exercise the bias path.
  * EDSR: torch's default Conv2d init, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
    (reference models/edsr.py builds plain nn.Conv2d; MeanShift keeps the default-initialised
    1x1 weights, models/edsr.py:129-136).
"""
from __future__ import annotations

import math

import numpy as np

NUM_FILTERS = 48


def larva_param_shapes(blocks, v2=False):
    """state_dict key -> shape, in the reference's registration order."""
    c = NUM_FILTERS
    shapes = {'head.feature_extraction.weight': (c, 3, 3, 3), 'head.feature_extraction.bias': (c,)}
    for i, nb in enumerate(blocks):
        for j in range(nb):
            for k in (0, 2):
                shapes[f'body_{i}.res_blocks.{j}.body.{k}.weight'] = (c, c, 3, 3)
                shapes[f'body_{i}.res_blocks.{j}.body.{k}.bias'] = (c,)
        for k in (0, 2):
            shapes[f'body_{i}.leg.recon_block.{k}.weight'] = (c, c, 3, 3)
            shapes[f'body_{i}.leg.recon_block.{k}.bias'] = (c,)
    if v2:
        shapes['tail.merge_conv.weight'] = (c, c * len(blocks), 3, 3)
        shapes['tail.merge_conv.bias'] = (c,)
        for k in (0, 2):
            shapes[f'tail.recon_block.{k}.weight'] = (c, c, 3, 3)
            shapes[f'tail.recon_block.{k}.bias'] = (c,)
    return shapes


def edsr_param_shapes(features=64, res_blocks=16, scale=4):
    f = features
    shapes = {'mean_shift.weight': (3, 3, 1, 1), 'mean_shift.bias': (3,),
              'first_conv.weight': (f, 3, 3, 3), 'first_conv.bias': (f,)}
    for j in range(res_blocks):
        for k in (0, 2):
            shapes[f'res_blocks.{j}.body.{k}.weight'] = (f, f, 3, 3)
            shapes[f'res_blocks.{j}.body.{k}.bias'] = (f,)
    shapes['after_res_conv.weight'] = (f, f, 3, 3)
    shapes['after_res_conv.bias'] = (f,)
    for s in range({2: 1, 4: 2, 8: 3}[scale]):
        shapes[f'upsample.body.{2 * s}.weight'] = (4 * f, f, 3, 3)
        shapes[f'upsample.body.{2 * s}.bias'] = (4 * f,)
    shapes['final_conv.weight'] = (3, f, 3, 3)
    shapes['final_conv.bias'] = (3,)
    shapes['mean_inverse_shift.weight'] = (3, 3, 1, 1)
    shapes['mean_inverse_shift.bias'] = (3,)
    return shapes


def make_larva_params(blocks, v2=False, seed=0, bias_std=0.0, weight_gain=0.1):
    rs = np.random.RandomState(seed)
    out = {}
    for key, shape in larva_param_shapes(blocks, v2).items():
        if key.endswith('weight'):
            fan_in = shape[1] * shape[2] * shape[3]
            std = weight_gain * math.sqrt(2.0 / fan_in)
            out[key] = (rs.standard_normal(shape) * std).astype(np.float32)
        else:
            out[key] = (rs.standard_normal(shape) * bias_std).astype(np.float32)
    return out


def make_edsr_params(features=64, res_blocks=16, scale=4, seed=0):
    rs = np.random.RandomState(seed)
    out = {}
    shapes = edsr_param_shapes(features, res_blocks, scale)
    for key, shape in shapes.items():
        wshape = shape if key.endswith('weight') else shapes[key[:-4] + 'weight']
        bound = 1.0 / math.sqrt(wshape[1] * wshape[2] * wshape[3])
        out[key] = rs.uniform(-bound, bound, size=shape).astype(np.float32)
    return out


def make_images(n, h, w, scale=4, seed=1, quantize=False):
    """LR input [n,3,h,w] and HR truth [n,3,scale*h,scale*w], float32 on the 0..255 scale.
    `quantize=True` rounds to integers (what a uint8 PNG would hold; exact in bf16)."""
    rs = np.random.RandomState(seed)
    lr = rs.uniform(0.0, 255.0, size=(n, 3, h, w)).astype(np.float32)
    hr = rs.uniform(0.0, 255.0, size=(n, 3, h * scale, w * scale)).astype(np.float32)
    if quantize:
        lr, hr = np.round(lr), np.round(hr)
    return lr, hr


def make_smooth_images(n, h, w, scale=4, seed=1):
    """Band-limited synthetic pair (HR = smooth random field, LR = its 4x4 box average) so that
    PSNR against the truth is finite and meaningful in the PSNR-delta parity tests."""
    rs = np.random.RandomState(seed)
    hh, ww = h * scale, w * scale
    yy, xx = np.meshgrid(np.arange(hh, dtype=np.float64), np.arange(ww, dtype=np.float64), indexing='ij')
    hr = np.zeros((n, 3, hh, ww), dtype=np.float64)
    for i in range(n):
        for c in range(3):
            acc = np.zeros((hh, ww))
            for _ in range(6):
                fy, fx = rs.uniform(0.005, 0.12, size=2)
                ph = rs.uniform(0, 2 * math.pi, size=2)
                acc += rs.uniform(0.3, 1.0) * np.sin(2 * math.pi * fy * yy + ph[0]) * np.cos(2 * math.pi * fx * xx + ph[1])
            acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-12)
            hr[i, c] = 16.0 + 224.0 * acc
    lr = hr.reshape(n, 3, h, scale, w, scale).mean(axis=(3, 5))
    return lr.astype(np.float32), hr.astype(np.float32)
