"""Developer diagnostic for a B200 box (not a test, not a product path): exercises the tensor-core kernels against the
CUDA-core kernels with structured operands so a wrong UMMA descriptor / layout shows up as a pattern, and prints quick
CUDA-event timings.  `torch.nn.functional.conv2d` is used here ONLY as a third opinion for the CUDA-core kernel.

    python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops  # noqa: E402


def pack(w, dtype, cin, **kw):
    O_, I_ = w.shape[:2]
    transpose = kw.get('transpose', 0)
    i_cnt = kw.get('i_cnt', I_)
    pc, pt = (i_cnt, O_) if transpose else (O_, i_cnt)
    packed = torch.zeros(ops.packed_weight_bytes(pc, pt, dtype), dtype=torch.uint8, device='cuda')
    ops.pack_weights([dict(w=w, packed=packed, cin=cin, dtype=dtype, **kw)])
    return packed


def nchw(x):
    """planar-8 activation [n,h,c/8,w,8] -> NCHW"""
    n, h, ch, w, _ = x.shape
    return x.permute(0, 2, 4, 1, 3).reshape(n, ch * 8, h, w).contiguous()


def run(fn, what):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f'[{what}] EXCEPTION: {type(e).__name__}: {e}', flush=True)
        return False
    return True


def conv_check(n, h, w, cin=48, cout=48, nsrc=1, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    wt = torch.randn((cout, cin * nsrc, 3, 3), device='cuda', generator=g) * 0.05
    b = torch.randn(cout, device='cuda', generator=g)
    xs = [torch.randn((n, h, cin // 8, w, 8), device='cuda', generator=g).to(torch.bfloat16) for _ in range(nsrc)]
    packed = pack(wt, torch.bfloat16, cin)
    o_tc = torch.empty((n, h, cout // 8, w, 8), dtype=torch.bfloat16, device='cuda')
    o_si = torch.empty_like(o_tc)
    ok = run(lambda: ops.conv3x3(xs, packed, cout, bias=b, out=o_si, simt=True), 'simt')
    ref = F.conv2d(torch.cat([nchw(x.float()) for x in xs], 1), wt.to(torch.bfloat16).float(), b, padding=1)
    e_si = (nchw(o_si.float()) - ref).abs().max().item()
    ok2 = run(lambda: ops.conv3x3(xs, packed, cout, bias=b, out=o_tc), 'tc')
    e_tc = (nchw(o_tc.float()) - ref).abs().max().item() if ok2 else float('nan')
    print(f'conv n={n} h={h} w={w} cin={cin}x{nsrc} cout={cout}: simt_vs_torch={e_si:.4g} tc_vs_torch={e_tc:.4g} '
          f'(ref max {ref.abs().max().item():.3g})', flush=True)
    return ok and ok2 and e_tc < 0.1


def conv_pattern():
    """weights non-zero for a single (tap, 16-channel K step): a 9x3 table of max errors."""
    n, h, w = 1, 16, 8
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn((n, h, 6, w, 8), device='cuda', generator=g).to(torch.bfloat16)
    full = torch.randn((48, 48, 3, 3), device='cuda', generator=g) * 0.1
    print('per (tap, kstep) max error of the tensor-core conv (rows: tap 0..8, cols: channels 0-15,16-31,32-47)')
    for tap in range(9):
        row = []
        for ks in range(3):
            wt = torch.zeros_like(full)
            wt[:, 16 * ks:16 * ks + 16, tap // 3, tap % 3] = full[:, 16 * ks:16 * ks + 16, tap // 3, tap % 3]
            packed = pack(wt, torch.bfloat16, 48)
            o = torch.empty((n, h, 6, w, 8), dtype=torch.bfloat16, device='cuda')
            if not run(lambda: ops.conv3x3([x], packed, 48, out=o), f'pattern tap{tap} ks{ks}'):
                return
            ref = F.conv2d(nchw(x.float()), wt.to(torch.bfloat16).float(), None, padding=1)
            row.append((nchw(o.float()) - ref).abs().max().item())
        print(f'  tap {tap}: ' + ' '.join(f'{v:9.4f}' for v in row), flush=True)


def wgrad_check(n, h, w, splits):
    g = torch.Generator(device='cuda').manual_seed(2)
    x = torch.randn((n, h, 6, w, 8), device='cuda', generator=g).to(torch.bfloat16)
    dy = torch.randn((n, h, 6, w, 8), device='cuda', generator=g).to(torch.bfloat16)
    res = {}
    for simt in (True, False):
        dw = torch.zeros((48, 48, 3, 3), device='cuda')
        db = torch.zeros(48, device='cuda')
        batch = ops.WgradBatch([dict(x=x, dy=dy, dw=dw, db=db)], splits=splits, device='cuda')
        if not run(lambda: batch.launch(simt=simt), f'wgrad simt={simt}'):
            return False
        res[simt] = (dw, db)
    xr = nchw(x.float()).requires_grad_(False)
    wt = torch.zeros((48, 48, 3, 3), device='cuda', requires_grad=True)
    out = F.conv2d(xr, wt, torch.zeros(48, device='cuda', requires_grad=True), padding=1)
    gw = torch.autograd.grad(out, wt, nchw(dy.float()))[0]
    gb = nchw(dy.float()).sum((0, 2, 3))
    for simt in (True, False):
        dw, db = res[simt]
        print(f'wgrad n={n} h={h} w={w} splits={splits} simt={simt}: rel dw={((dw - gw).norm() / gw.norm()).item():.3g} '
              f'rel db={((db - gb).norm() / gb.norm()).item():.3g}', flush=True)
    dw = res[False][0]
    if ((dw - gw).norm() / gw.norm()).item() > 1e-3:
        err = (dw - gw).abs()
        print('  tc wgrad err by tap:', [f'{err[:, :, t // 3, t % 3].max().item():.3g}' for t in range(9)])
        print('  tc wgrad err by cout block of 8:', [f'{err[8 * k:8 * k + 8].max().item():.3g}' for k in range(6)])
        print('  tc wgrad err by cin block of 8:', [f'{err[:, 8 * k:8 * k + 8].max().item():.3g}' for k in range(6)])
    return True


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def conv_timing():
    for (n, h, w) in [(1, 180, 320), (16, 48, 48), (1, 270, 480), (32, 270, 480)]:
        g = torch.Generator(device='cuda').manual_seed(3)
        x = torch.randn((n, h, 6, w, 8), device='cuda', generator=g).to(torch.bfloat16)
        wt = torch.randn((48, 48, 3, 3), device='cuda', generator=g) * 0.05
        b = torch.zeros(48, device='cuda')
        packed = pack(wt, torch.bfloat16, 48)
        o = torch.empty_like(x)
        for ctas in (0, 296):
            us = timeit(lambda: ops.conv3x3([x], packed, 48, bias=b, out=o, relu=True, res1=x, max_ctas=ctas))
            fl = 2 * 20736 * n * h * w
            print(f'conv48 tc n={n} {h}x{w} max_ctas={ctas}: {us:.1f} us  {fl / us * 1e-6:.1f} TFLOP/s', flush=True)


def main():
    print(torch.cuda.get_device_name(0), 'sm count', ops.device_check(0), flush=True)
    t0 = time.time()
    ok = conv_check(1, 16, 8)
    if not ok:
        conv_pattern()
    conv_check(2, 19, 13)
    conv_check(1, 40, 24, nsrc=4)
    conv_check(1, 33, 17, cin=64, cout=64)
    conv_check(1, 20, 9, cin=64, cout=256)
    wgrad_check(1, 16, 8, 1)
    wgrad_check(2, 35, 21, 4)
    conv_timing()
    print('probe done in %.1fs, launches=%d' % (time.time() - t0, _lib.launch_count()), flush=True)


if __name__ == '__main__':
    main()
