"""Time the two chain kernels on the same layer list: 16x8-tile data-flow chain (tap-major weights, csrc/conv_chain.cu) vs
row-marching chain (ky-stacked weights, csrc/conv_row.cu).  Prints us per layer and TFLOP/s per shape; the numbers behind
LarvaEngine.use_row_path's threshold.

    python tools/row_vs_tile.py [layers] [n,h,w ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops  # noqa: E402

SHAPES = [(16, 48, 48), (1, 180, 320), (16, 64, 64), (1, 270, 480), (32, 64, 64), (2, 270, 480), (64, 64, 64),
          (4, 270, 480), (8, 270, 480), (32, 270, 480)]


def build(n, h, w, layers, wlayout, dev, c=48):
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((n, h, c // 8, w, 8), device=dev, generator=g).to(torch.bfloat16)
    wt = torch.randn((c, c, 3, 3), device=dev, generator=g) * 0.05
    b = torch.zeros(c, device=dev)
    packed = torch.zeros(ops.packed_weight_bytes(c, c, torch.bfloat16), dtype=torch.uint8, device=dev)
    ops.pack_weights([dict(w=wt, packed=packed, cin=c, dtype=torch.bfloat16, wlayout=wlayout)])
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    args, src = [], x
    for i in range(layers):
        dst = bufs[i & 1]
        args.append(ops.make_conv_args([src], packed, c, bias=b, out=dst, relu=(i & 1) == 0,
                                       res1=None if (i & 1) == 0 else x, wlayout=wlayout))
        src = dst
    return args, (x, wt, b, packed, bufs)


def time_chain(args, ws, reps):
    for _ in range(2):
        ops.conv3x3_chain(args, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv3x3_chain(args, ws)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    shapes = [tuple(int(v) for v in a.split(',')) for a in sys.argv[2:]] or SHAPES
    dev = torch.device('cuda', 0)
    print(f'{"shape":>16} {"px":>9} | {"tile us/layer":>13} {"TF/s":>7} | {"row us/layer":>12} {"TF/s":>7} | row/tile')
    for n, h, w in shapes:
        ws = ops.chain_workspace(n, h, w, dev)
        res = []
        for wl in (_lib.LV_W_TAP_MAJOR, _lib.LV_W_KY_STACKED):
            args, keep = build(n, h, w, layers, wl, dev)
            sec = time_chain(args, ws, 5 if n * h * w < 2e6 else 2)
            res.append(sec)
            del args, keep
        fl = 2.0 * 9 * 48 * 48 * n * h * w * layers
        print(f'{str((n, h, w)):>16} {n * h * w:9d} | {res[0] / layers * 1e6:13.2f} {fl / res[0] / 1e12:7.1f} | '
              f'{res[1] / layers * 1e6:12.2f} {fl / res[1] / 1e12:7.1f} | {res[0] / res[1]:.2f}x', flush=True)


if __name__ == '__main__':
    main()
