"""Chop-forward upscaling (drop-in for reference utils/image_utils.py:7-66): split the LR image into four overlapping
quadrants, upscale each through `model.upscale`, stitch by hard crop.  Like the reference this is NOT exact unless
the overlap covers the network's receptive field (35 LR px for S=16); SURVEY.md section 8(f) ranks an exact,
GPU-stitched variant as follow-up work."""
import numpy as np


def _split_image(image, chop, overlap_size):
    if not chop:
        return [image]
    _, height, width = image.shape
    sh, sw, half = height // 2, width // 2, overlap_size // 2
    return [np.array(image[:, :sh + half, :sw + half]), np.array(image[:, :sh + half, sw - half:]),
            np.array(image[:, sh - half:, :sw + half]), np.array(image[:, sh - half:, sw - half:])]


def _combine_images(images, input_image, scale, chop, overlap_size):
    if len(images) == 1:
        return images[0]
    _, height, width = input_image.shape
    nh, nw = (height // 2) * scale, (width // 2) * scale
    nhalf = (overlap_size // 2) * scale
    out = np.zeros([3, height * scale, width * scale])
    out[:, :nh, :nw] = images[0][:, :nh, :nw]
    out[:, :nh, nw:] = images[1][:, :nh, nhalf:]
    out[:, nh:, :nw] = images[2][:, nhalf:, :nw]
    out[:, nh:, nw:] = images[3][:, nhalf:, nhalf:]
    return out


def upscale_with_chop_forward(model, input_image, scale, overlap_size):
    parts = _split_image(input_image, chop=True, overlap_size=overlap_size)
    outs = [model.upscale(input_list=[p], scale=scale)[0] for p in parts]
    return _combine_images(outs, input_image=input_image, scale=scale, chop=True, overlap_size=overlap_size)
