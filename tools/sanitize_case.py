"""Tiny workload for compute-sanitizer (racecheck / synccheck / memcheck): both chain kernels on (3,20,9) with 5 CTAs and on
(2,24,40), the batched weight-gradient kernel and the per-layer conv kernel.  Small on purpose: the sanitizer slows the
flag-spinning persistent kernels down by orders of magnitude."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops  # noqa: E402
from tests.test_gpu_kernels import _act, _chain_layers, _pack, _rand_conv  # noqa: E402


def main():
    rs = np.random.RandomState(5)
    for (n, h, w), ctas in (((3, 20, 9), 5), ((2, 24, 40), 0)):
        for wl in (0, 1):
            make_args, make_bufs = _chain_layers(rs, n, h, w, depth=2, wlayout=wl)
            b = make_bufs()
            ws = ops.chain_workspace(n, h, w, 'cuda')
            ops.conv3x3_chain(make_args(b), ws, max_ctas=ctas)
            torch.cuda.synchronize()
            print('chain ok', (n, h, w), 'row' if wl else 'tile', flush=True)
    dtype = torch.bfloat16
    wt, bb, _ = _rand_conv(rs, 48, 48, dtype)
    x, _ = _act(rs, 2, 48, 19, 13, dtype)
    dy, _ = _act(rs, 2, 48, 19, 13, dtype)
    out = torch.empty_like(x)
    ops.conv3x3([x], _pack(wt, dtype, 48), 48, bias=torch.from_numpy(bb).cuda(), out=out, relu=True, res1=x)
    dw = torch.zeros((48, 48, 3, 3), device='cuda')
    db = torch.zeros(48, device='cuda')
    ops.WgradBatch([dict(x=x, dy=dy, dw=dw, db=db, overwrite=True)], splits=3, device='cuda').launch()
    torch.cuda.synchronize()
    print('sanitize workload done', flush=True)


if __name__ == '__main__':
    main()
