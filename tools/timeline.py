"""Developer tool: per-role clock64 timeline of CTA 0 of one tensor-core conv launch."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops

def main():
    n, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 270, 480)
    ctas = int(sys.argv[4]) if len(sys.argv) > 4 else 296
    mode = sys.argv[5] if len(sys.argv) > 5 else 'fwd'
    ky = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    lib = _lib.load()
    lib.lv_debug_set_timeline.argtypes = [ctypes.c_void_p]
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn((n, h, 6, w, 8), device='cuda', generator=g).to(torch.bfloat16)
    wt = torch.randn((48, 48, 3, 3), device='cuda', generator=g) * 0.05
    b = torch.zeros(48, device='cuda')
    packed = torch.zeros(ops.packed_weight_bytes(48, 48, torch.bfloat16), dtype=torch.uint8, device='cuda')
    ops.pack_weights([dict(w=wt, packed=packed, cin=48, dtype=torch.bfloat16, wlayout=ky)])
    o = torch.empty_like(x)
    kw = dict(relu=True, res1=x) if mode == 'fwd' else {}
    kw['wlayout'] = ky
    for _ in range(3):
        ops.conv3x3([x], packed, 48, bias=b, out=o, max_ctas=ctas, **kw)
    tl = torch.zeros(9 * 64 * 4, dtype=torch.int64, device='cuda')
    lib.lv_debug_set_timeline(tl.data_ptr())
    ops.conv3x3([x], packed, 48, bias=b, out=o, max_ctas=ctas, **kw)
    torch.cuda.synchronize()
    lib.lv_debug_set_timeline(None)
    t = tl.cpu().view(9, 64, 4)
    t0 = int(t[t > 0].min())
    names = {0: ('prod', ['wait_empty', 'got_empty', 'issued', 'arrived_full']),
             1: ('mma ', ['wait_tempty', 'got_tempty', 'got_full', 'committed']),
             2: ('epi ', ['wait_tfull', 'got_tfull', 'tmem_read', 'stored']),
             3: ('epi2', ['after_barA', 'got_res', 'after_barB', 'computed']),
             4: ('epi3', ['fenced', 'loop_top', 'prefetched', '-']),
             5: ('w0  ', ['afterA', 'computed', 'afterB', '-']), 6: ('w1  ', ['afterA', 'computed', 'afterB', '-']),
             7: ('w2  ', ['afterA', 'computed', 'afterB', '-']), 8: ('w3  ', ['afterA', 'computed', 'afterB', '-'])}
    for k in range(6):
        for role in range(9):
            nm, evs = names[role]
            vals = [int(t[role, k, e]) - t0 if t[role, k, e] > 0 else -1 for e in range(4)]
            print(f'tile {k:2d} {nm}: ' + '  '.join(f'{e}={v}' for e, v in zip(evs, vals)))
    print()

main()
