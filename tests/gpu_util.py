"""Helpers shared by the `-m gpu` tests: oracle-side reference computations on bf16-rounded operands."""
import numpy as np
import torch

from oracle import larva_oracle as O


def bf16_round(a):
    """numpy float -> values representable in bf16 (round-to-nearest-even), returned as float64."""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32)
    return t.numpy().astype(np.float64)


def to_nhwc(x_nchw, dtype, device='cuda'):
    return torch.from_numpy(np.ascontiguousarray(x_nchw.transpose(0, 2, 3, 1), dtype=np.float32)).to(device).to(dtype).contiguous()


def from_nhwc(t):
    return t.detach().to(torch.float32).cpu().numpy().transpose(0, 3, 1, 2).astype(np.float64)


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def load_params(module, params):
    sd = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()}
    module.load_state_dict(sd)
