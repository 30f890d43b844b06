"""Developer tool: compare the cluster-resident strip chain (conv_strip.cu) with per-layer launches, layer by layer."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import ops


def main():
    n, h, w, L = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (1, 16, 8, 3)
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn((n, h, 6, w, 8), device='cuda', generator=g).to(torch.bfloat16)
    packs, biases = [], []
    for i in range(L):
        wt = torch.randn((48, 48, 3, 3), device='cuda', generator=g) * 0.05
        pk = torch.zeros(ops.packed_weight_bytes(48, 48, torch.bfloat16), dtype=torch.uint8, device='cuda')
        ops.pack_weights([dict(w=wt, packed=pk, cin=48, dtype=torch.bfloat16)])
        packs.append(pk)
        biases.append(torch.randn(48, device='cuda', generator=g))

    def build(outs):
        Ls, src = [], x
        for i in range(L):
            Ls.append(ops.make_conv_args([src], packs[i], 48, bias=biases[i], out=outs[i], relu=(i % 2 == 0)))
            src = outs[i]
        return Ls
    ref = [torch.zeros_like(x) for _ in range(L)]
    for a in build(ref):
        ops.conv3x3_launch(a)
    got = [torch.zeros_like(x) for _ in range(L)]
    ws = ops.chain_workspace(n, h, w, 'cuda')
    ops.conv3x3_chain(build(got), ws)
    torch.cuda.synchronize()
    for i in range(L):
        d = (got[i].float() - ref[i].float()).abs()          # [n, h, 6, w, 8]
        bad = (d > 0)
        print(f'layer {i}: max diff {d.max().item():.4f}, mismatching elements {int(bad.sum())} / {bad.numel()}')
        if bad.any():
            rows = bad.any(dim=4).any(dim=2).any(dim=0)     # [h, w]
            print('   bad rows:', sorted(set(rows.any(dim=1).nonzero().flatten().tolist()))[:40])
            print('   bad cols:', sorted(set(rows.any(dim=0).nonzero().flatten().tolist()))[:60])
            chunks = bad.any(dim=4).any(dim=3).any(dim=1).any(dim=0)
            print('   bad chunks:', chunks.nonzero().flatten().tolist())


main()
