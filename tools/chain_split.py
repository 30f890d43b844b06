"""Developer tool: time the forward and backward-data conv chains of one training step separately, and per-layer-kind
micro chains (relu / res / mask / res1+res2 / PixelShuffle+loss) on the training shape."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import ops, synth, _lib
from models.LarvaNet import LarvaNet


def main():
    n, h, w = 16, 48, 48
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn((n, h, 6, w, 8), device='cuda', generator=g).to(torch.bfloat16)
    wt = torch.randn((48, 48, 3, 3), device='cuda', generator=g) * 0.05
    b = torch.zeros(48, device='cuda')
    packed = torch.zeros(ops.packed_weight_bytes(48, 48, torch.bfloat16), dtype=torch.uint8, device='cuda')
    ops.pack_weights([dict(w=wt, packed=packed, cin=48, dtype=torch.bfloat16)])
    bufs = [torch.empty_like(x) for _ in range(3)]
    base = torch.zeros((n, 3, 4 * h, 4 * w), device='cuda')
    truth = torch.zeros_like(base)
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    ws = ops.chain_workspace(n, h, w, 'cuda')
    kinds = {
        'plain': lambda s, d: dict(out=d),
        'relu': lambda s, d: dict(out=d, relu=True),
        'res1': lambda s, d: dict(out=d, res1=x),
        'mask': lambda s, d: dict(out=d, mask=x),
        'res1+res2': lambda s, d: dict(out=d, res1=x, res2=bufs[2]),
        'ps4+loss': lambda s, d: dict(epilogue=_lib.LV_EPI_PS4_NCHW, base_hr=base, truth_hr=truth, loss_sum=loss, grad_sign=d),
    }
    for name, mk in kinds.items():
        L, src = [], x
        for i in range(24):
            d = bufs[i & 1]
            L.append(ops.make_conv_args([src], packed, 48, bias=b, **mk(src, d)))
            src = d
        for _ in range(3):
            ops.conv3x3_chain(L, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv3x3_chain(L, ws)
        e1.record()
        torch.cuda.synchronize()
        print(f'{name:10s}: {e0.elapsed_time(e1) / 10 / 24 * 1e3:.2f} us/layer')


main()
