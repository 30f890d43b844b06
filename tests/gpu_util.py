"""Helpers shared by the `-m gpu` tests: oracle-side reference computations on bf16-rounded operands."""
import numpy as np
import torch

from oracle import larva_oracle as O


def bf16_round(a):
    """numpy float -> values representable in bf16 (round-to-nearest-even), returned as float64."""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32)
    return t.numpy().astype(np.float64)


def to_nhwc(x_nchw, dtype, device='cuda'):
    """NCHW numpy -> the library's planar-8 activation layout [n][h][c/8][w][8] (name kept from the NHWC days)."""
    n, c, h, w = x_nchw.shape
    a = np.ascontiguousarray(x_nchw.reshape(n, c // 8, 8, h, w).transpose(0, 3, 1, 4, 2), dtype=np.float32)
    return torch.from_numpy(a).to(device).to(dtype).contiguous()


def from_nhwc(t):
    """planar-8 activation tensor -> NCHW float64 numpy."""
    n, h, ch, w, _ = t.shape
    a = t.detach().to(torch.float32).cpu().numpy()
    return a.transpose(0, 2, 4, 1, 3).reshape(n, ch * 8, h, w).astype(np.float64)


def act_empty(n, h, w, c, dtype, device='cuda'):
    return torch.empty((n, h, c // 8, w, 8), dtype=dtype, device=device)


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def load_params(module, params):
    sd = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()}
    module.load_state_dict(sd)
