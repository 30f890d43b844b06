"""Timeline of CTA 0's roles in the row-marching conv kernel for 64 consecutive rows (needs a -DLV_ROW_TRACE=1 build:
python tools/build_variant.py trace -DLV_ROW_TRACE=1; LARVANET_B200_LIB=.../lib_trace.so python tools/row_trace.py n h w).
Prints, per row, clock stamps relative to the first one: producer (buffer free, copies issued), scheduler (start,
accumulator blocks free, row staged, command written), issuer (command read, MMAs issued, commits issued), and per
output row the epilogue group's (start waiting, accumulator complete, stored)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops  # noqa: E402
from tools.row_vs_tile import build  # noqa: E402

FIRST, ROWS = 100, 64


def main():
    n, h, w = (int(v) for v in sys.argv[1:4])
    layers = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    mode = sys.argv[5] if len(sys.argv) > 5 else 'single'
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    lib.lv_debug_set_timeline.argtypes = [C.c_void_p]
    args, keep = build(n, h, w, layers, _lib.LV_W_KY_STACKED, dev)
    ws = ops.chain_workspace(n, h, w, dev)
    run = (lambda: ops.conv3x3_chain(args, ws)) if mode == 'chain' else (lambda: ops.conv3x3_launch(args[0], 0))
    run()
    torch.cuda.synchronize()
    stats = torch.zeros(64 + 8 * ROWS * 4, dtype=torch.int64, device=dev)
    lib.lv_debug_set_timeline(C.c_void_p(stats.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.lv_debug_set_timeline(None)
    t = stats[64:].cpu().view(8, ROWS, 4)
    base = int(t[t > 0].min())
    rel = lambda v: int(v) - base if int(v) > 0 else -1
    print('row   | producer free  issued | sched start  tfree  staged  written | issuer read  mmas  commits')
    for r in range(ROWS):
        p, s, i = t[0, r], t[1, r], t[2, r]
        print(f'{FIRST + r:5d} | {rel(p[0]):8d} {rel(p[1]):8d} | {rel(s[0]):8d} {rel(s[1]):7d} {rel(s[2]):7d} {rel(s[3]):7d} | '
              f'{rel(i[0]):8d} {rel(i[1]):7d} {rel(i[2]):7d}')
    print('row   | issuer: command read, then after each of the 9 accumulate MMAs (delta clk)')
    for r in range(ROWS):
        st = [int(t[2, r, 0])] + [int(t[5 + (i >> 2), r, i & 3]) for i in range(9)] + [int(t[2, r, 2])]
        print(f'{FIRST + r:5d} | ' + ' '.join(f'{b - a:5d}' for a, b in zip(st[:-1], st[1:])))
    print('out row | group | wait-from  complete  stored')
    for r in range(ROWS):
        g = 3 + ((FIRST + r) & 1)
        e = t[g, r]
        print(f'{FIRST + r:7d} | g{g - 3}    | {rel(e[0]):8d} {rel(e[1]):8d} {rel(e[2]):8d}')


if __name__ == '__main__':
    main()
