"""Developer tool: bisect the chain test's layer list on the strip kernel (first prefix length whose buffers differ)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import ops
from tests.test_gpu_kernels import _chain_layers

n, h, w = (int(v) for v in sys.argv[1:4])
rs = np.random.RandomState(5)
make_args, make_bufs = _chain_layers(rs, n, h, w, depth=5)
full = len(make_args(make_bufs()))
ws = ops.chain_workspace(n, h, w, 'cuda')
for L in range(1, full + 1):
    ref = make_bufs()
    for a in make_args(ref)[:L]:
        ops.conv3x3_launch(a)
    got = make_bufs()
    ops.conv3x3_chain(make_args(got)[:L], ws)
    torch.cuda.synchronize()
    bad = [k for k in ref if k != 'loss' and not torch.equal(got[k], ref[k])]
    a = make_args(got)[L - 1]
    print(f'prefix {L:2d}: last layer relu={a.relu} mask={bool(a.mask)} res1={bool(a.res1)} res2={bool(a.res2)} epi={a.epilogue}  differing buffers: {bad}')
    if bad:
        k = bad[0]
        d = (got[k].float() - ref[k].float()).abs()
        if d.dim() == 5:
            m = (d > 0).any(dim=4).any(dim=2)
            print('   bad images:', m.any(dim=2).any(dim=1).nonzero().flatten().tolist(), ' rows:', m.any(dim=2).any(dim=0).nonzero().flatten().tolist()[:50], ' cols:', m.any(dim=1).any(dim=0).nonzero().flatten().tolist()[:50])
        break
