"""Exact tile-sharded inference: cut a frame into horizontal bands with a halo as wide as the network's receptive field,
upscale the bands independently (one after another on one GPU to bound memory, or one band per rank on several GPUs)
and stitch the band interiors on the device.

The reference's chop-forward (`utils/image_utils.py:7-66`, `validate.py:51-52,95-96`) overlaps its four quadrants by
`overlap_size // 2` = 10 LR pixels by default and hard-crops at the cut, although a LarvaNet output pixel depends on
1 (head) + 2*sum(num_blocks) + 2 (leg) conv rings + 2 px of bicubic support = 37 LR pixels for the 16-resblock
network: its result differs from the full-frame one near every cut.  Here the halo is the full receptive field, so every
stitched pixel sees exactly the inputs it sees in the full frame -- and, because each kernel accumulates a pixel's taps in
a fixed order that does not depend on where the pixel sits in a tile, the stitched frame equals the full-frame result
BIT FOR BIT (tests/test_gpu_network.py::test_tile_sharded_inference_is_exact).

No collective is needed to compute; `gather_bands` (optional) assembles the frame on one rank with torch.distributed.
"""
from __future__ import annotations

import torch

from . import dist as lvdist


def receptive_halo(blocks, v2=False, exit_leg=None):
    """LR pixels an output pixel depends on in each direction: head conv + 2 convs per residual block of every body that
    runs + 2 leg convs (V2 tail: merge conv + 2) and, for the bicubic base, 2."""
    k = len(blocks) if exit_leg is None else int(exit_leg)
    if k == 0:
        return 2
    convs = 1 + 2 * sum(blocks[:k]) + (3 if (v2 and exit_leg is None) else 2)
    return convs + 2      # 35 + 2 = 37 for M=4, B=4,4,4,4: conv rings, plus the bicubic taps' margin


def band_ranges(height, bands, halo):
    """[(y0, y1, lo, hi)]: band b owns LR rows [y0, y1) and reads rows [lo, hi) = the band plus `halo` rows each side,
    clipped at the frame (where the network's own zero padding / clamped bicubic taps apply, as in the full frame)."""
    bands = max(1, min(int(bands), height))
    out = []
    for b in range(bands):
        y0, y1 = lvdist.shard_range(height, b, bands)
        out.append((y0, y1, max(0, y0 - halo), min(height, y1 + halo)))
    return out


def upscale_banded(engine, x, bands, halo=None, exit_leg=None, uint8=False, only=None):
    """x: [1 or n, 3, H, W] fp32 CUDA tensor.  Returns the stitched [n, 3, 4H, 4W] frame (fp32, or uint8 with `uint8`).
    `only`: iterable of band indices to compute (the others are left untouched in the returned buffer) -- a rank of a
    multi-GPU run passes its own bands."""
    n, _, h, w = (int(v) for v in x.shape)
    if halo is None:
        halo = receptive_halo(engine.blocks, engine.v2, exit_leg)
    out = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.uint8 if uint8 else torch.float32, device=x.device)
    ranges = band_ranges(h, bands, halo)
    todo = range(len(ranges)) if only is None else only
    for b in todo:
        y0, y1, lo, hi = ranges[b]
        if y1 <= y0:
            continue
        piece = engine.forward(x[:, :, lo:hi, :].contiguous(), exit_leg=exit_leg, uint8=uint8)
        out[:, :, 4 * y0:4 * y1, :] = piece[:, :, 4 * (y0 - lo):4 * (y1 - lo), :]
    return out


def bands_for_rank(bands, rank, world):
    """Round-robin band indices of `rank` (no collective)."""
    return lvdist.frames_for_rank(bands, rank, world)


def gather_bands(local, height, bands, rank, world, dst=0, group=None):
    """Assemble the full frame on `dst` from every rank's `upscale_banded(..., only=bands_for_rank(...))` buffer.
    The only communication of the tile-sharded path; skipped entirely when every rank writes its own bands to disk."""
    import torch.distributed as tdist
    if world == 1:
        return local
    ranges = band_ranges(height, bands, 0)
    parts = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    tdist.gather(local, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = parts[dst].clone()
    for r in range(world):
        for b in bands_for_rank(len(ranges), r, world):
            y0, y1, _, _ = ranges[b]
            out[:, :, 4 * y0:4 * y1, :] = parts[r][:, :, 4 * y0:4 * y1, :]
    return out
