"""GPU-resident synthetic training loader: the images live in HBM and the patch pipeline (random crop, rot90, flip) runs
on the device, writing the batch tensors the training step reads -- no per-step host->device copy of the batch.

Mirrors the contract of reference dataloaders/div2k_train_loader_tensor.py:57-97 (`get_patch_batch` returns TENSORS, not
lists of numpy arrays) with the same augmentation semantics (crop, `torch.rot90(k, dims=(1,2))` with k in 1..4, flip of the
last axis with probability 1/2).  The random draws stay on the host (numpy RandomState, 6 ints per patch = one small pinned
copy per step); the pixels never leave the device (`lv_crop_augment`).
"""
import ctypes as C

import numpy as np
import torch

from dataloaders.synthetic_loader import SyntheticLoader
from larvanet_b200 import _lib, ops


def create_loader():
    return SyntheticTensorLoader()


class SyntheticTensorLoader(SyntheticLoader):
    def prepare(self, scales):
        super().prepare(scales)
        if not torch.cuda.is_available():
            raise _lib.LarvaNetB200Error('synthetic_loader_tensor keeps its images on a CUDA device; there is no CPU path')
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.dev_lr = {s: [torch.from_numpy(np.ascontiguousarray(im)).to(self.device) for im in self.lr[s]] for s in scales}
        self.dev_hr = {s: [torch.from_numpy(np.ascontiguousarray(im)).to(self.device) for im in self.hr[s]] for s in scales}
        self._slots = {}     # (batch, patch, scale) -> ring of (pinned items, device items, out_lr, out_hr)
        self._turn = 0

    def draw(self, batch_size, scale, input_patch_size):
        """The step's random draws: [(image_index, y, x, rot, flip)] (host RNG, reference div2k_train_loader.py:78-93)."""
        out = []
        for _ in range(batch_size):
            idx = self.rs.randint(self.get_num_images())
            _, h, w = self.lr[scale][idx].shape
            y = self.rs.randint(h - input_patch_size + 1)
            x = self.rs.randint(w - input_patch_size + 1)
            out.append((idx, y, x, int(self.rs.randint(4)) + 1, int(self.rs.uniform() < 0.5)))
        return out

    def get_patch_batch(self, batch_size, scale, input_patch_size, draws=None):
        """-> (input_tensor [B,3,p,p], truth_tensor [B,3,scale*p,scale*p]) float32 CUDA tensors.  The tensors belong to a
        ring of 3 slots: a batch stays valid while the next two are being produced."""
        key = (batch_size, input_patch_size, scale)
        ring = self._slots.get(key)
        if ring is None:
            p, ph = input_patch_size, input_patch_size * scale
            ring = [(torch.empty(batch_size * C.sizeof(_lib.PatchItem), dtype=torch.uint8).pin_memory(),
                     torch.empty(batch_size * C.sizeof(_lib.PatchItem), dtype=torch.uint8, device=self.device),
                     torch.empty((batch_size, 3, p, p), dtype=torch.float32, device=self.device),
                     torch.empty((batch_size, 3, ph, ph), dtype=torch.float32, device=self.device)) for _ in range(3)]
            self._slots[key] = ring
        host, dev, out_lr, out_hr = ring[self._turn % 3]
        self._turn += 1
        draws = self.draw(batch_size, scale, input_patch_size) if draws is None else draws
        items = (_lib.PatchItem * batch_size).from_buffer(host.numpy())
        for b, (idx, y, x, rot, flip) in enumerate(draws):
            lr, hr = self.dev_lr[scale][idx], self.dev_hr[scale][idx]
            it = items[b]
            it.lr, it.hr = lr.data_ptr(), hr.data_ptr()
            it.h, it.w, it.y, it.x, it.rot, it.flip = int(lr.shape[1]), int(lr.shape[2]), int(y), int(x), int(rot), int(flip)
        dev.copy_(host, non_blocking=True)
        ops.crop_augment(dev, batch_size, out_lr, out_hr, input_patch_size, scale)
        return out_lr, out_hr
