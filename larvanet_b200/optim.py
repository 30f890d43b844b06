"""AdamW over the engine's flat parameter arena: one `lv_adamw_step` launch per step instead of ~80 foreach ops.

Subclasses torch.optim.AdamW so callers that touch `optim.param_groups[0]['lr']`, attach an LR scheduler
(ReduceLROnPlateau in reference models/LarvaNet.py:90-92) or call `zero_grad()` keep working; hyper-parameters are
read from `param_groups[0]` at every step exactly like torch does.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import LarvaNetB200Error


class FusedAdamW(torch.optim.AdamW):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._engine = None
        self._exp_avg = None
        self._exp_avg_sq = None
        self._steps = 0

    def attach(self, engine):
        """Bind to a LarvaEngine / EdsrEngine whose arena holds exactly this optimizer's parameters."""
        if len(self.param_groups) != 1:
            raise LarvaNetB200Error('FusedAdamW supports a single param group')
        ours = {id(p) for p in self.param_groups[0]['params']}
        theirs = {id(p) for p in engine.arena.params.values() if p.requires_grad}
        if ours != theirs:
            raise LarvaNetB200Error('FusedAdamW: optimizer parameters differ from the engine arena')
        if any(not p.requires_grad for p in engine.arena.params.values()):
            raise LarvaNetB200Error('FusedAdamW: frozen parameters inside the arena are not supported')
        self._engine = engine
        if hasattr(engine, '_dp_optim'):
            engine._dp_optim = self
            engine._train.clear()      # buffers built before the optimizer was known use the unfused exchange
        self._exp_avg = torch.zeros_like(engine.arena.flat)
        self._exp_avg_sq = torch.zeros_like(engine.arena.flat)

    def zero_grad(self, set_to_none=True):
        if self._engine is None:
            return super().zero_grad(set_to_none)
        self._engine.arena.grad.zero_()

    # No autograd graph can arise here (raw-pointer kernels on the arena), so no torch.no_grad() wrapper; and torch's
    # per-class profiler wrapper around step() (~25 us of host time per call) is declined below with `step.hooked`.
    def step(self, closure=None):
        if closure is not None:
            raise LarvaNetB200Error('FusedAdamW does not support closures')
        if self._engine is None:
            raise LarvaNetB200Error('FusedAdamW.step() before attach(engine): there is no unfused fallback')
        g = self.param_groups[0]
        self._steps += 1
        a = self._engine.arena
        if getattr(self._engine, '_dp_fused', lambda: False)():
            # data parallel: gradient all-reduce over peer memory + update + re-pack in ONE kernel
            if not self._engine._dp_pending:
                raise LarvaNetB200Error('FusedAdamW.step() in data-parallel mode needs a preceding engine.train_step()')
            self._engine._dp_pending = False
            ops.dp_adamw_pack_step(a.flat, self._exp_avg, self._exp_avg_sq, g['lr'], g['betas'][0], g['betas'][1], g['eps'],
                                   g['weight_decay'], self._steps, self._engine._fused_convs, self._engine._symm.peers, 1.0)
            self._engine.weights_updated_and_packed()
            return
        if getattr(self._engine, 'fused_update_available', lambda: False)():
            # update + re-pack of the conv operands (forward and backward-data forms) in one kernel
            ops.adamw_pack_step(a.flat, a.grad, self._exp_avg, self._exp_avg_sq, g['lr'], g['betas'][0], g['betas'][1],
                                g['eps'], g['weight_decay'], self._steps, self._engine._fused_convs, 1.0)
            self._engine.weights_updated_and_packed()
            return
        ops.adamw_step(a.flat, a.grad, self._exp_avg, self._exp_avg_sq, g['lr'], g['betas'][0], g['betas'][1], g['eps'],
                       g['weight_decay'], self._steps, 1.0)
        if hasattr(self._engine, 'weights_updated'):
            self._engine.weights_updated()
        else:
            self._engine.mark_weights_changed()


# torch.optim.Optimizer.__init__ wraps `cls.step` in a profiler hook unless the function says it already is
FusedAdamW.step.hooked = True
