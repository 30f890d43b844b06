"""Single-kernel micro-benchmark / ncu target: one 48->48 conv (bias+ReLU+residual epilogue) on a given shape."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=8)
    ap.add_argument('--h', type=int, default=270)
    ap.add_argument('--w', type=int, default=480)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--ctas', type=int, default=296)
    ap.add_argument('--mode', default='fwd', choices=['fwd', 'plain', 'ps', 'wgrad'])
    ap.add_argument('--splits', type=int, default=296)
    ap.add_argument('--ky', type=int, default=0)
    ap.add_argument('--chain', type=int, default=0, help='time ONE lv_conv3x3_chain launch of this many dependent convs')
    ap.add_argument('--graph', type=int, default=0, help='time a CUDA graph of this many DEPENDENT convs (ping-pong)')
    a = ap.parse_args()
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn((a.n, a.h, 6, a.w, 8), device='cuda', generator=g).to(torch.bfloat16)
    wt = torch.randn((48, 48, 3, 3), device='cuda', generator=g) * 0.05
    b = torch.zeros(48, device='cuda')
    packed = torch.zeros(ops.packed_weight_bytes(48, 48, torch.bfloat16), dtype=torch.uint8, device='cuda')
    ops.pack_weights([dict(w=wt, packed=packed, cin=48, dtype=torch.bfloat16, wlayout=a.ky)])
    o = torch.empty_like(x)
    if a.mode == 'fwd':
        fn = lambda: ops.conv3x3([x], packed, 48, bias=b, out=o, relu=True, res1=x, max_ctas=a.ctas, wlayout=a.ky)
    elif a.mode == 'plain':
        fn = lambda: ops.conv3x3([x], packed, 48, bias=b, out=o, max_ctas=a.ctas, wlayout=a.ky)
    elif a.mode == 'ps':
        hr = torch.empty((a.n, 3, 4 * a.h, 4 * a.w), device='cuda')
        base = torch.zeros_like(hr)
        fn = lambda: ops.conv3x3([x], packed, 48, bias=b, epilogue=1, out_hr=hr, base_hr=base, max_ctas=a.ctas)
    else:
        dw = torch.zeros((48, 48, 3, 3), device='cuda')
        db = torch.zeros(48, device='cuda')
        batch = ops.WgradBatch([dict(x=x, dy=o.copy_(x), dw=dw, db=db)], splits=a.splits, device='cuda')
        fn = batch.launch
    if a.chain:
        o2 = torch.empty_like(x)
        bufs = [o, o2]
        L, src = [], x
        for i in range(a.chain):
            dst = bufs[i & 1]
            L.append(ops.make_conv_args([src], packed, 48, bias=b, out=dst, relu=(i & 1) == 0,
                                        res1=None if (i & 1) == 0 else x))
            src = dst
        ws = ops.chain_workspace(a.n, a.h, a.w, 'cuda')
        fn = lambda: ops.conv3x3_chain(L, ws, max_ctas=a.ctas)
        a.graph = 0
    if a.graph:
        o2 = torch.empty_like(x)
        bufs = [o, o2]
        def chain():
            src = x
            for i in range(a.graph):
                dst = bufs[i & 1]
                ops.conv3x3([src], packed, 48, bias=b, out=dst, relu=(i & 1) == 0, res1=None if (i & 1) == 0 else x,
                            max_ctas=a.ctas, wlayout=a.ky)
                src = dst
        chain()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            chain()
        fn = gr.replay
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / a.iters * 1e3 / max(a.graph, a.chain, 1)
    fl = 2 * 20736 * a.n * a.h * a.w
    print(f'{a.mode} n={a.n} {a.h}x{a.w} ctas={a.ctas}: {us:.1f} us/launch  {fl / us * 1e-6:.1f} TFLOP/s')


if __name__ == '__main__':
    main()
