"""Import the UNMODIFIED reference modules (staged by oracle/make_ref.py under oracle/_ref/, or straight from
/root/reference in the build container) and drive them the way the reference's own scripts do.

TEST / BASELINE INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's baseline legs import this.

The reference's packages are called `models`, `utils`, `dataloaders`, `validate` -- the same names this repo's drop-in
plugins use -- so they are imported with sys.modules swapped and handed back as plain module objects; this repo's own
modules are restored afterwards.

`RefTrainer` restates the 12 lines of the reference's train step (models/LarvaNet.py:102-114, models/LarvaNetV2.py:
105-123) around the reference's own nn.Modules and torch.optim.AdamW: `LarvaNet.prepare(is_training=True)` itself
cannot run on torch >= 2.7 because it passes `verbose=True` to ReduceLROnPlateau.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.path.join(HERE, '_ref'), '/root/reference']
_SHADOWED = ('models', 'utils', 'dataloaders', 'validate')
_cache = None


def ref_root():
    for c in CANDIDATES:
        if os.path.exists(os.path.join(c, 'models', 'LarvaNet.py')):
            return c
    return None


def load():
    """{'LarvaNet': module, 'LarvaNetV2': ..., 'LarvaLeg': ..., 'LarvaLegV2': ..., 'edsr': ..., 'root': path} or None."""
    global _cache
    if _cache is not None:
        return _cache
    root = ref_root()
    if root is None:
        return None
    saved = {k: v for k, v in sys.modules.items() if k in _SHADOWED or k.split('.')[0] in _SHADOWED}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        mods = {n: importlib.import_module('models.' + n) for n in ('LarvaNet', 'LarvaNetV2', 'LarvaLeg', 'LarvaLegV2', 'edsr')}
        mods['validate'] = sys.modules.get('validate')
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k in _SHADOWED or k.split('.')[0] in _SHADOWED]:
            del sys.modules[k]
        sys.modules.update(saved)
    mods['root'] = root
    _cache = mods
    return mods


def larva_args(blocks, leg=None):
    a = types.SimpleNamespace(num_modules=len(blocks), num_blocks=','.join(str(b) for b in blocks), interpolate='bicubic')
    if leg is not None:
        a.leg = int(leg)
    return a


def make_module(blocks, v2=False, params=None, leg=None, device='cpu', dtype=None):
    """The reference's LarvaNetModule (or the LarvaLeg / LarvaLegV2 variant when `leg` is given) with `params` loaded."""
    import torch
    mods = load()
    if mods is None:
        raise RuntimeError('reference modules unavailable (no oracle/_ref and no /root/reference)')
    if leg is None:
        mod = mods['LarvaNetV2' if v2 else 'LarvaNet'].LarvaNetModule(larva_args(blocks))
    else:
        mod = mods['LarvaLegV2' if v2 else 'LarvaLeg'].LarvaNetModule(larva_args(blocks, leg))
    if params is not None:
        mod.load_state_dict({k: torch.as_tensor(v) for k, v in params.items()})
    mod = mod.to(device)
    if dtype is not None:
        mod = mod.to(dtype)
    return mod


def make_edsr(features, res_blocks, params=None, device='cpu', res_weight=1.0):
    import torch
    mods = load()
    args = types.SimpleNamespace(edsr_conv_features=features, edsr_res_blocks=res_blocks, edsr_res_weight=res_weight)
    mod = mods['edsr'].EDSRModule(args, scale=4)
    if params is not None:
        mod.load_state_dict({k: torch.as_tensor(v) for k, v in params.items()})
    return mod.to(device)


class RefTrainer:
    """The reference's multi-exit train step on the reference's own modules (any device / autocast mode)."""

    def __init__(self, params, blocks, v2=False, lr=4e-4, device='cpu', autocast_dtype=None, channels_last=False):
        import torch
        self.torch = torch
        self.blocks, self.v2 = list(blocks), bool(v2)
        self.model = make_module(blocks, v2, params, device=device)
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.loss_fn = torch.nn.L1Loss()
        self.optim = torch.optim.AdamW(filter(lambda p: p.requires_grad, self.model.parameters()), lr=lr)
        self.autocast_dtype = autocast_dtype
        self.device = torch.device(device)

    def _loss(self, x, truth):
        m = self.model
        fea = m.head(x)
        base = m.base(x)
        loss = 0
        feats = []
        for i in range(len(self.blocks)):
            fea = getattr(m, f'body_{i}')(fea)
            feats.append(fea)
            out = getattr(m, f'body_{i}').leg(fea, base)
            loss = loss + self.loss_fn(out.float(), truth)
        if self.v2:
            out = m.tail(feats, base)
            loss = loss + self.loss_fn(out.float(), truth)
            return loss / (len(self.blocks) + 1)
        return loss / len(self.blocks)

    def loss_and_backward(self, x, truth):
        torch = self.torch
        if self.autocast_dtype is not None:
            with torch.autocast(self.device.type, dtype=self.autocast_dtype):
                loss = self._loss(x, truth)
        else:
            loss = self._loss(x, truth)
        self.optim.zero_grad()
        loss.backward()
        return loss

    def step(self, x, truth, sync=True):
        loss = self.loss_and_backward(x, truth)
        self.optim.step()
        return float(loss.item()) if sync else loss

    def infer(self, x):
        torch = self.torch
        with torch.no_grad():
            if self.autocast_dtype is not None:
                with torch.autocast(self.device.type, dtype=self.autocast_dtype):
                    return self.model(x)
            return self.model(x)

    def grads(self):
        return {n: p.grad.detach().cpu().numpy() for n, p in self.model.named_parameters()}

    def params(self):
        return {n: p.detach().cpu().numpy() for n, p in self.model.named_parameters()}
