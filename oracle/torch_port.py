"""PyTorch-CPU port of the reference hot path, used ONLY as the timed CPU baseline (`bench.py` cpu_baseline leg and
`--impl reference`) and cross-checked against the golden fixtures in tests/test_oracle_golden.py.

TEST / BASELINE INFRASTRUCTURE -- the product path never imports this.  The reference itself is Python that calls
torch.nn ops (oneDNN on CPU); it cannot travel to the GPU box, so this file restates its module graph with
torch.nn.functional on the same state_dict layout: LarvaNetModule.forward (reference models/LarvaNet.py:287-293),
the train step core incl. AdamW (models/LarvaNet.py:102-114), LarvaNetV2 (models/LarvaNetV2.py:105-123,355-365).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _conv(x, p, prefix):
    return F.conv2d(x, p[prefix + '.weight'], p[prefix + '.bias'], padding=1)


def _resblock(x, p, prefix):
    return x + _conv(F.relu(_conv(x, p, prefix + '.body.0')), p, prefix + '.body.2')


def _body(x, p, i, nb):
    a = x
    for j in range(nb):
        a = _resblock(a, p, f'body_{i}.res_blocks.{j}')
    return x + a


def _recon(fea, base, p, prefix):
    v = _conv(F.relu(_conv(fea, p, prefix + '.recon_block.0')), p, prefix + '.recon_block.2')
    return F.pixel_shuffle(v, 4) + base


def forward(p, x, blocks, v2=False):
    fea = _conv(x, p, 'head.feature_extraction')
    feats = []
    for i, nb in enumerate(blocks):
        fea = _body(fea, p, i, nb)
        feats.append(fea)
    base = F.interpolate(x, scale_factor=4, mode='bicubic', align_corners=False)
    if v2:
        mf = _conv(torch.cat(feats, 1), p, 'tail.merge_conv')
        return _recon(mf, base, p, 'tail')
    return _recon(fea, base, p, f'body_{len(blocks) - 1}.leg')


def loss_fn(p, x, truth, blocks, v2=False):
    fea = _conv(x, p, 'head.feature_extraction')
    base = F.interpolate(x, scale_factor=4, mode='bicubic', align_corners=False)
    loss = 0
    feats = []
    for i, nb in enumerate(blocks):
        fea = _body(fea, p, i, nb)
        feats.append(fea)
        loss = loss + F.l1_loss(_recon(fea, base, p, f'body_{i}.leg'), truth)
    if v2:
        mf = _conv(torch.cat(feats, 1), p, 'tail.merge_conv')
        loss = loss + F.l1_loss(_recon(mf, base, p, 'tail'), truth)
        return loss / (len(blocks) + 1)
    return loss / len(blocks)


class CpuTrainer:
    """fwd + bwd + AdamW on the host, like the reference's train_step_larva on a CPU-only machine."""

    def __init__(self, params_np, blocks, v2=False, lr=4e-4, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.blocks, self.v2 = list(blocks), v2
        self.p = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in params_np.items()}
        self.optim = torch.optim.AdamW(list(self.p.values()), lr=lr)

    def step(self, x, truth):
        loss = loss_fn(self.p, x, truth, self.blocks, self.v2)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return float(loss.item())

    @torch.no_grad()
    def infer(self, x):
        return forward(self.p, x, self.blocks, self.v2)
