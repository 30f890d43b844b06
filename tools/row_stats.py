"""Per-role cycle breakdown of the row-marching conv kernel (CTA 0): where a row's time goes.
    python tools/row_stats.py n h w [layers]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops  # noqa: E402
from tools.row_vs_tile import build  # noqa: E402


def main():
    n, h, w = (int(v) for v in sys.argv[1:4])
    layers = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    ctas = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    lib.lv_debug_set_timeline.argtypes = [C.c_void_p]
    args, keep = build(n, h, w, layers, _lib.LV_W_KY_STACKED, dev)
    ws = ops.chain_workspace(n, h, w, dev)
    for mode in ('chain', 'single'):
        stats = torch.zeros(64, dtype=torch.int64, device=dev)
        run = (lambda: ops.conv3x3_chain(args, ws, max_ctas=ctas)) if mode == 'chain' else (lambda: [ops.conv3x3_launch(a, ctas) for a in args])
        run()
        torch.cuda.synchronize()
        lib.lv_debug_set_timeline(C.c_void_p(stats.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        lib.lv_debug_set_timeline(None)
        s = stats.cpu().tolist()
        us = e0.elapsed_time(e1) * 1e3
        rows = max(s[3], 1)
        print(f'== {mode}: {n}x{h}x{w}, {layers} layers: {us / layers:.1f} us/layer; CTA0 input rows {s[3]}, '
              f'{us * 1e-6 * 1.965e9 / rows:.0f} clk per input row (at 1965 MHz)')
        print(f'   scheduler   per row: probes+setup {s[0] / rows:6.0f}  barrier-wait {s[4] / rows:6.0f}  slot-wait {s[1] / rows:6.0f}  '
              f'write+publish {s[5] / rows:6.0f}  rows not staged {s[13] / rows:4.2f}  blocks not drained {s[14] / rows:4.2f}')
        print(f'   issuer      per row: cmd-wait {s[7] / rows:6.0f}  mma-issue {s[2] / rows:6.0f}  commits {s[6] / rows:6.0f}  '
              f'polls {s[12] / rows:5.2f}')
        pr = max(s[11], 1)
        print(f'   producer    per row: flag-wait {s[8] / pr:7.0f}  empty-wait {s[9] / pr:7.0f}  issue {s[10] / pr:7.0f}   (rows {s[11]})')
        r0 = max(s[18], 1)
        print(f'   epilogue g0 detail : issue-loads {s[32] / r0:6.0f}  acc-wait {s[33] / r0:6.0f}  ldtm0 {s[34] / r0:6.0f}  '
              f'half0->ldtm1 {s[35] / r0:6.0f}  whole fn {s[36] / r0:6.0f}')
        for g in (0, 1):
            b = 16 + 8 * g
            r = max(s[b + 2], 1)
            print(f'   epilogue g{g} per row: tfull-wait {s[b] / r:7.0f}  drain+store {s[b + 1] / r:7.0f}   (rows {s[b + 2]})')


if __name__ == '__main__':
    main()
