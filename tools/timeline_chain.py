"""Developer tool: per-role clock64 timeline of CTA 0 of one lv_conv3x3_chain launch (first 64 jobs)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops


def main():
    n, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 180, 320)
    layers = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    show = int(sys.argv[5]) if len(sys.argv) > 5 else 16
    lib = _lib.load()
    lib.lv_debug_set_timeline.argtypes = [ctypes.c_void_p]
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn((n, h, 6, w, 8), device='cuda', generator=g).to(torch.bfloat16)
    wt = torch.randn((48, 48, 3, 3), device='cuda', generator=g) * 0.05
    b = torch.zeros(48, device='cuda')
    packed = torch.zeros(ops.packed_weight_bytes(48, 48, torch.bfloat16), dtype=torch.uint8, device='cuda')
    ops.pack_weights([dict(w=wt, packed=packed, cin=48, dtype=torch.bfloat16)])
    bufs = [torch.empty_like(x), torch.empty_like(x)]
    L, src = [], x
    for i in range(layers):
        dst = bufs[i & 1]
        L.append(ops.make_conv_args([src], packed, 48, bias=b, out=dst, relu=(i & 1) == 0, res1=None if (i & 1) == 0 else x))
        src = dst
    ws = ops.chain_workspace(n, h, w, 'cuda')
    for _ in range(3):
        ops.conv3x3_chain(L, ws)
    tl = torch.zeros(9 * 64 * 4, dtype=torch.int64, device='cuda')
    lib.lv_debug_set_timeline(tl.data_ptr())
    ops.conv3x3_chain(L, ws)
    torch.cuda.synchronize()
    lib.lv_debug_set_timeline(None)
    t = tl.cpu().view(9, 64, 4)
    t0 = int(t[t > 0].min())
    names = {0: ('prod', ['deps_ok/lyr_wait', 'got_empty/lyr_ok', 'issued/w_ok', 'polled/fenced']),
             3: ('pub ', ['arrived', 'released', '-', '-']),
             4: ('st h0', ['q0', 'q1', 'q2', 'q3']), 5: ('st h1', ['q0', 'q1', 'q2', 'q3']),
             6: ('rl h0', ['q0', 'q1', 'q2', 'q3']), 7: ('rl h1', ['q0', 'q1', 'q2', 'q3']),
             1: ('mma ', ['wait_tempty', 'got_tempty', 'got_full', 'committed']),
             2: ('epi ', ['wait_tfull', 'got_tfull', 'tmem_read', 'stored'])}
    for k in range(show):
        for role in (0, 1, 2, 3, 4, 5, 6, 7):
            nm, evs = names[role]
            vals = [int(t[role, k, e]) - t0 if t[role, k, e] > 0 else -1 for e in range(4)]
            print(f'job {k:2d} {nm}: ' + '  '.join(f'{e}={v}' for e, v in zip(evs, vals)))


main()
