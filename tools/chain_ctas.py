"""Sweep the grid size of the tile chain kernel: us per layer for several CTA counts per shape (a grid that divides the
tile count keeps every CTA on the same number of tiles per layer).

    python tools/chain_ctas.py [layers]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops  # noqa: E402
from tools.row_vs_tile import build  # noqa: E402

CASES = {(16, 48, 48): [148, 144, 96, 72], (16, 64, 64): [148, 128, 103], (1, 180, 320): [148, 120, 96, 80],
         (128, 64, 64): [148, 128], (1, 270, 480): [148, 128, 146]}


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device('cuda', 0)
    for (n, h, w), grids in CASES.items():
        tiles = n * ((h + 15) // 16) * ((w + 7) // 8)
        ws = ops.chain_workspace(n, h, w, dev)
        args, keep = build(n, h, w, layers, _lib.LV_W_TAP_MAJOR, dev)
        out = []
        for g in grids:
            for _ in range(2):
                ops.conv3x3_chain(args, ws, max_ctas=g)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv3x3_chain(args, ws, max_ctas=g)
            e1.record()
            torch.cuda.synchronize()
            out.append(f'{g} CTAs ({tiles / g:.2f} tiles each): {e0.elapsed_time(e1) * 1e3 / 5 / layers:.2f} us')
        print(f'{(n, h, w)} {tiles} tiles | ' + ' | '.join(out), flush=True)


if __name__ == '__main__':
    main()
