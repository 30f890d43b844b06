"""Fused execution engine for LarvaNet / LarvaNetV2: the whole forward (and backward) as a fixed chain of
`lv_*` kernel launches over pre-allocated NHWC buffers, optionally replayed as one CUDA graph.

What the reference does with ~40 Conv2d modules + autograd (models/LarvaNet.py:98-114, :287-293;
models/LarvaNetV2.py:101-123, :355-365) becomes:

  forward   head+bicubic kernel -> one conv kernel per Conv2d with bias/ReLU/residual/body-skip fused in the epilogue
            -> exit conv writes PixelShuffle(4)+base (and, when training, the L1 partial sums and the sign gradient)
  backward  one conv kernel per Conv2d for backward-data (180-degree rotated weights; ReLU mask and branch-gradient
            accumulation fused in the epilogue), then BATCHED weight-gradient launches (one per body) that write
            straight into a flat fp32 gradient arena, so zero_grad is one memset, the data-parallel exchange is one
            allreduce per body slice and AdamW is one kernel.

Activations are planar-8 ([n][h][c/8][w][8], see csrc/lv_common.cuh) in `act_dtype` (bf16 = product path on tcgen05 tensor cores, fp32 = validation mode).
Everything here is plumbing: tensors, streams, graphs.  No arithmetic happens in PyTorch.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops
from ._lib import LV_EPI_NHWC, LV_EPI_PS4_NCHW, LarvaNetB200Error

C = 48  # LarvaNet feature width (reference models/LarvaNet.py:226 -- hard-wired)


class ParamArena:
    """Flat fp32 parameter + gradient storage; every module parameter becomes a view into it (same shapes, same
    state_dict keys), so the engine can hand raw offsets to the kernels and treat all gradients as one buffer."""

    def __init__(self, module, device):
        named = [(n, p) for n, p in module.named_parameters()]
        self.names = [n for n, _ in named]
        total = sum(p.numel() for _, p in named)
        self.flat = torch.empty(total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.offsets = {}
        self.views = {}
        self.grad_views = {}
        self.params = {}
        off = 0
        for n, p in named:
            k = p.numel()
            v = self.flat[off:off + k].view(p.shape)
            v.copy_(p.data.to(device=device, dtype=torch.float32))
            p.data = v
            gv = self.grad[off:off + k].view(p.shape)
            self.offsets[n] = (off, k)
            self.views[n] = v
            self.grad_views[n] = gv
            self.params[n] = p
            off += k
        self.total = total

    def replace_grad(self, flat):
        """Move the gradient arena into `flat` (>= total elements, e.g. a symmetric-memory allocation)."""
        self.grad = flat
        for n, (off, k) in self.offsets.items():
            self.grad_views[n] = flat[off:off + k].view(self.params[n].shape)
            if self.params[n].grad is not None:
                self.params[n].grad = None

    def attach_grads(self):
        for n, p in self.params.items():
            if p.requires_grad and p.grad is not self.grad_views[n]:
                p.grad = self.grad_views[n]

    def version(self):
        return sum(p._version for p in self.params.values())

    def slice_of(self, prefix):
        """(start, end) element range covering all parameters whose name starts with `prefix`."""
        lo, hi = None, None
        for n in self.names:
            if n.startswith(prefix):
                o, k = self.offsets[n]
                lo = o if lo is None else min(lo, o)
                hi = o + k if hi is None else max(hi, o + k)
        return lo, hi


class _Bufs:
    pass


class DeviceScalar:
    """The step's loss: a float64 device accumulator times a host-side factor, read back lazily.  `.item()` / `float()`
    are the only device->host traffic (8 bytes) and no extra kernels are launched for the scaling."""

    def __init__(self, acc, scale):
        self._acc, self._scale = acc, float(scale)

    def item(self):
        return float(self._acc.item()) * self._scale

    __float__ = item

    def tensor(self):
        return self._acc[0] * self._scale

    def __repr__(self):
        return f'DeviceScalar({self.item():.6f})'


def _capture_graph(run):
    """Warm up `run()` eagerly, then capture it into a CUDA graph.  Returns (graph, launches) or (None, 0) when the
    capture was invalidated (seen once after unrelated allocator churn); the caller then keeps launching the same CUDA
    kernels eagerly -- slower to launch, numerically identical, never a different code path."""
    import gc
    import warnings
    run()  # warm-up: module loading, cudaFuncSetAttribute, allocator growth all happen outside the capture
    torch.cuda.current_stream().synchronize()
    for attempt in range(2):
        gc.collect()
        g = torch.cuda.CUDAGraph()
        c0 = _lib.launch_count()
        try:
            with torch.cuda.graph(g, capture_error_mode='thread_local'):
                run()
            return g, _lib.launch_count() - c0
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            err = e
    warnings.warn(f'larvanet_b200: CUDA graph capture failed twice ({type(err).__name__}: {err}); running eagerly')
    return None, 0


class _ShapeCache:
    """Per-input-shape buffer sets + captured CUDA graphs, bounded: least-recently-used shapes are dropped (their buffers
    and graph pool are freed) once more than `capacity` are alive, and a shape is only captured into a graph the second
    time it is seen -- validate.py / get_sr.py walk over images of ~100 different sizes, each seen once per pass, and
    must neither accumulate hundreds of MB per size nor pay a warm-up + capture pass for a one-off shape."""

    def __init__(self, capacity):
        import collections
        self.capacity = max(1, int(capacity))
        self.entries = collections.OrderedDict()   # key -> dict(bufs, graph, launches, hits)

    def get(self, key, build):
        ent = self.entries.get(key)
        if ent is None:
            while len(self.entries) >= self.capacity:
                _, old = self.entries.popitem(last=False)
                old['graph'] = None
                old['bufs'] = None
            ent = dict(bufs=build(), graph=None, launches=0, hits=0)
            self.entries[key] = ent
        else:
            self.entries.move_to_end(key)
        ent['hits'] += 1
        return ent

    def clear(self):
        self.entries.clear()

    def __len__(self):
        return len(self.entries)


def _run_cached(engine, ent, run):
    """Run `run()` for a cached shape: eagerly the first time, through a CUDA graph (captured on the second use) after."""
    if engine.use_graphs and not engine.simt and ent['hits'] >= 2:
        if ent['graph'] is None:
            g, nl = _capture_graph(run)
            ent['graph'] = g if g is not None else False
            ent['launches'] = nl
        if ent['graph']:
            ent['graph'].replay()
            engine.replayed_launches += ent['launches']
            return
    run()


class LarvaEngine:
    def __init__(self, module, blocks, v2=False, act_dtype=torch.bfloat16, device=None, use_graphs=None):
        self.module = module
        self.blocks = list(blocks)
        self.m = len(self.blocks)
        self.v2 = bool(v2)
        self.act_dtype = act_dtype
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != 'cuda':
            raise LarvaNetB200Error('larvanet_b200 runs on CUDA devices only (no CPU fallback)')
        self.sm_count = ops.device_check(self.device.index)
        if use_graphs is None:
            use_graphs = os.environ.get('LARVANET_B200_GRAPHS', '1') != '0'
        self.use_graphs = use_graphs
        self.max_ctas = int(os.environ.get('LARVANET_B200_MAX_CTAS', '0'))
        self.wgrad_splits = int(os.environ.get('LARVANET_B200_WGRAD_SPLITS', '0'))
        self.arena = ParamArena(module, self.device)
        self._conv_layers = self._enumerate_convs()
        self._head_grad_slice = self.arena.slice_of('head.')
        self._alloc_packed()
        self._packed_version = None
        self._seen_version = None
        self._packed_bwd = False
        cap = int(os.environ.get('LARVANET_B200_SHAPE_CACHE', '4'))
        self._infer = _ShapeCache(cap)   # (n, h, w, exit_leg) -> buffers + graph, LRU-bounded
        self._train = _ShapeCache(2)
        self.simt = False  # tests flip this to cross-check the tensor-core kernels on CUDA cores
        # consecutive 48->48 convs run as ONE persistent data-flow launch (ops.conv3x3_chain); LARVANET_B200_CHAIN=0
        # falls back to one launch per conv
        self.use_chain = os.environ.get('LARVANET_B200_CHAIN', '1') != '0'
        self._chain = None      # list of pending ConvArgs while a chain is being recorded
        self._chain_ws = {}     # (n, h, w) -> flag workspace
        self._row_active = False  # the pass being recorded uses the row-marching kernel (ky-stacked operands)
        self.row_min_pixels = int(os.environ.get('LARVANET_B200_ROW_MIN_PIXELS', str(640 * 1024)))
        self.replayed_launches = 0  # kernels executed through CUDA-graph replays (lv_launch_count only sees eager ones)
        # data parallel
        self.world_size = 1
        self.process_group = None
        self._symm = None       # symmetric-memory gradient exchange (data parallel), else NCCL
        self._dp_optim = None   # FusedAdamW bound to this engine (enables the fused exchange+optimizer kernel)
        self._dp_pending = False

    # ------------------------------------------------------------------ layers / packed weights
    def _enumerate_convs(self):
        layers = []  # (prefix, O, I)
        for i, nb in enumerate(self.blocks):
            for j in range(nb):
                for k in (0, 2):
                    layers.append((f'body_{i}.res_blocks.{j}.body.{k}', C, C))
            for k in (0, 2):
                layers.append((f'body_{i}.leg.recon_block.{k}', C, C))
        if self.v2:
            layers.append(('tail.merge_conv', C, C * self.m))
            for k in (0, 2):
                layers.append((f'tail.recon_block.{k}', C, C))
        return layers

    def _alloc_packed(self):
        dt = self.act_dtype
        self._pk = {}      # (prefix, 'fwd') / (prefix, 'bwd', s) -> uint8 view
        items_f, items_b = [], []
        sizes = []
        for prefix, O, I in self._conv_layers:
            sizes.append(((prefix, 'fwd'), ops.packed_weight_bytes(O, I, dt)))
            for s in range(I // C):
                sizes.append(((prefix, 'bwd', s), ops.packed_weight_bytes(C, O, dt)))
        total = sum((b + 255) // 256 * 256 for _, b in sizes)
        self._packed = torch.zeros(total, dtype=torch.uint8, device=self.device)
        off = 0
        for key, b in sizes:
            self._pk[key] = self._packed[off:off + b]
            off += (b + 255) // 256 * 256
        for prefix, O, I in self._conv_layers:
            w = self.arena.views[prefix + '.weight']
            items_f.append(dict(w=w, packed=self._pk[(prefix, 'fwd')], transpose=0, i_off=0, i_cnt=I, cin=C, dtype=dt))
            for s in range(I // C):
                items_b.append(dict(w=w, packed=self._pk[(prefix, 'bwd', s)], transpose=1, i_off=C * s, i_cnt=C, cin=C,
                                    dtype=dt))
        self._pack_items_fwd, self._pack_items_bwd = items_f, items_b
        # ky-stacked copies of the single-source operands for the row-marching kernel (csrc/conv_row.cu); refreshed
        # lazily, only when a pass actually takes the row path
        self._ky = None
        if dt == torch.bfloat16:
            single = [(p, O, I) for p, O, I in self._conv_layers if I == C and O == C]
            nbytes = ops.packed_weight_bytes(C, C, dt)
            step = (nbytes + 255) // 256 * 256
            self._packed_ky = torch.zeros(2 * step * len(single), dtype=torch.uint8, device=self.device)
            kf, kb = [], []
            for k, (prefix, O, I) in enumerate(single):
                w = self.arena.views[prefix + '.weight']
                self._pk[(prefix, 'fwd', 'ky')] = self._packed_ky[2 * k * step:2 * k * step + nbytes]
                self._pk[(prefix, 'bwd', 0, 'ky')] = self._packed_ky[(2 * k + 1) * step:(2 * k + 1) * step + nbytes]
                kf.append(dict(w=w, packed=self._pk[(prefix, 'fwd', 'ky')], transpose=0, i_off=0, i_cnt=C, cin=C, dtype=dt,
                               wlayout=_lib.LV_W_KY_STACKED))
                kb.append(dict(w=w, packed=self._pk[(prefix, 'bwd', 0, 'ky')], transpose=1, i_off=0, i_cnt=C, cin=C, dtype=dt,
                               wlayout=_lib.LV_W_KY_STACKED))
            self._ky = dict(fwd=ops.build_pack_arrays(kf), all=ops.build_pack_arrays(kf + kb), version=None, bwd=False)
        # the optimizer can update the weights and emit both operand forms in ONE kernel (bf16, 48-output convs only)
        self._fused_convs = None
        if dt == torch.bfloat16 and os.environ.get('LARVANET_B200_FUSED_ADAMW', '1') != '0' and len(self._conv_layers) <= 64 \
                and all(O == C and I % C == 0 and I // C <= 4 for _, O, I in self._conv_layers):
            convs = [dict(w_off=self.arena.offsets[prefix + '.weight'][0], cin_total=I, fwd=self._pk[(prefix, 'fwd')],
                          bwd=[self._pk[(prefix, 'bwd', s)] for s in range(I // C)]) for prefix, O, I in self._conv_layers]
            convs.sort(key=lambda c: c['w_off'])
            self._fused_convs = ops.build_fused_convs(convs)
        # marshalled once: re-packing runs every optimizer step and must not cost host time per layer
        self._pack_arr_fwd = ops.build_pack_arrays(items_f)
        self._pack_arr_all = ops.build_pack_arrays(items_f + items_b)

    def repack(self, backward=False, force=False):
        """Refresh the packed bf16/fp32 conv operands from the fp32 master weights when they changed."""
        ver = self._seen_version = self.arena.version()     # one walk over the parameters per step (~15 us of host time)
        if not force and self._packed_version is not None:
            if ver == self._packed_version and (self._packed_bwd or not backward):
                return
        ops.pack_weights_prebuilt(self._pack_arr_all if backward else self._pack_arr_fwd)   # launch first, book-keep after
        self._packed_version = ver
        self._packed_bwd = backward

    def repack_ky(self, backward=False):
        """Refresh the ky-stacked operands of the row-marching kernel (only called by passes that take the row path)."""
        ky = self._ky
        if ky['version'] == self.arena.version() and ky['version'] is not None and (ky['bwd'] or not backward):
            return
        ops.pack_weights_prebuilt(ky['all'] if backward else ky['fwd'])
        ky['version'], ky['bwd'] = self.arena.version(), backward

    def use_row_path(self, n, h, w):
        """Policy: the row-marching kernel (9 MMAs of N=144 per 128 px, tensor bound) needs jobs of several rows per SM to
        amortise its two halo rows and fill its 128-pixel lanes; small problems (one 320x180 frame, 16 patches of 48x48:
        2-3 tiles per SM, bound by the layer-to-layer hand-over latency, not by MMA throughput) stay on the 16x8-tile
        kernel.  LARVANET_B200_ROW=0/1 forces either; default threshold from tools/row_vs_tile.py measurements."""
        if self._ky is None or self.simt:
            return False
        mode = os.environ.get('LARVANET_B200_ROW', 'auto')
        if mode in ('0', '1'):
            return mode == '1'
        return n * h * w >= self.row_min_pixels

    def mark_weights_changed(self):
        self._packed_version = None
        if self._ky is not None:
            self._ky['version'] = None

    def fused_update_available(self):
        return self._fused_convs is not None and not self.simt

    def weights_updated_and_packed(self):
        """Called by FusedAdamW after lv_adamw_pack_step: both operand forms are current.  The kernel updates the arena
        through raw pointers (no tensor version changes), so the version seen by this step's `repack` still stands; a
        parameter modified in place between train_step() and optim.step() only costs one extra re-pack next step."""
        self._packed_version = self._seen_version if self._seen_version is not None else self.arena.version()
        self._packed_bwd = True

    def weights_updated(self):
        """Called by FusedAdamW right after its update kernel: re-pack NOW, on the same stream, in the form the last
        pass used -- the next step's first launch is then not preceded by host-side re-pack work while the GPU idles."""
        self._packed_version = None
        if self._packed_bwd:
            self.repack(backward=True)

    # ------------------------------------------------------------------ helpers
    def _w(self, prefix):
        return self.arena.views[prefix + '.weight'], self.arena.views[prefix + '.bias']

    def _conv(self, srcs, prefix, out=None, **kw):
        _, b = self._w(prefix)
        if self._row_active and len(srcs) == 1:
            self._emit(ops.make_conv_args(srcs, self._pk[(prefix, 'fwd', 'ky')], C, bias=b, out=out,
                                          wlayout=_lib.LV_W_KY_STACKED, **kw))
        else:
            self._emit(ops.make_conv_args(srcs, self._pk[(prefix, 'fwd')], C, bias=b, out=out, **kw))

    def _dgrad(self, dy, prefix, out, s=0, **kw):
        if self._row_active and (prefix, 'bwd', s, 'ky') in self._pk:
            self._emit(ops.make_conv_args([dy], self._pk[(prefix, 'bwd', s, 'ky')], C, bias=None, out=out,
                                          wlayout=_lib.LV_W_KY_STACKED, **kw))
        else:
            self._emit(ops.make_conv_args([dy], self._pk[(prefix, 'bwd', s)], C, bias=None, out=out, **kw))

    def _emit(self, args):
        """Launch one conv, or queue it while a chain is being recorded (flushed by the first conv that cannot join: another
        shape, several sources, or the other weight layout -- a chain runs on ONE of the two chain kernels)."""
        if self._chain is not None and ops.chain_eligible(args) and \
                (not self._chain or self._chain[0].wlayout == args.wlayout):
            self._chain.append(args)
            return
        self._flush_chain()
        ops.conv3x3_launch(args, self.max_ctas, self.simt)

    def _begin_chain(self):
        on = self.use_chain and not self.simt and self.act_dtype == torch.bfloat16
        self._chain = [] if on else None

    def _flush_chain(self, end=False):
        pending = self._chain
        if pending:
            self._chain = []
            if len(pending) == 1:
                ops.conv3x3_launch(pending[0], self.max_ctas, self.simt)
            else:
                a0 = pending[0]
                key = (a0.n, a0.h, a0.w)
                if key not in self._chain_ws:
                    self._chain_ws[key] = ops.chain_workspace(a0.n, a0.h, a0.w, self.device)
                ops.conv3x3_chain(pending, self._chain_ws[key], self.max_ctas)
        if end:
            self._chain = None

    def _act(self, n, h, w, c=C):
        return ops.act_empty(n, h, w, c, self.act_dtype, self.device)

    # ------------------------------------------------------------------ inference
    def _build_infer(self, n, h, w, exit_leg=None):
        b = _Bufs()
        b.x = torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device)
        b.base = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=self.device)
        b.out = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=self.device)
        b.f0 = self._act(n, h, w)
        b.t = self._act(n, h, w)
        b.pp = [self._act(n, h, w), self._act(n, h, w)]
        b.feats = [self._act(n, h, w) for _ in range(self.m)]
        b.u = self._act(n, h, w)
        b.mf = self._act(n, h, w) if self.v2 else None
        b.row = self.use_row_path(n, h, w)
        b.out_u8 = None
        return b

    def _run_infer(self, b, exit_leg=None, u8=False):
        """`u8`: the exit conv's epilogue writes the uint8 frame (round + clip) INSTEAD of the fp32 one."""
        self._row_active = b.row
        outs = dict(out_u8=b.out_u8) if u8 else dict(out_hr=b.out)
        hw, hb = self._w('head.feature_extraction')
        k = self.m if exit_leg is None else exit_leg
        ops.head_bicubic(b.x, hw, hb, b.f0, b.base)
        if k == 0:
            if u8:
                ops.image_to_uint8(b.base, b.out_u8)
            else:
                b.out.copy_(b.base)
            return
        self._begin_chain()
        fin = b.f0
        for i in range(k):
            a = fin
            nb = self.blocks[i]
            for j in range(nb):
                p = f'body_{i}.res_blocks.{j}.body'
                self._conv([a], p + '.0', out=b.t, relu=True)
                last = j == nb - 1
                dst = b.feats[i] if last else b.pp[j & 1]
                self._conv([b.t], p + '.2', out=dst, res1=a, res2=fin if last else None)
                a = dst
            if nb == 0:  # empty body: x + x
                raise LarvaNetB200Error('num_blocks entries must be >= 1')
            fin = b.feats[i]
        if self.v2 and exit_leg is None:
            self._conv(b.feats, 'tail.merge_conv', out=b.mf)
            self._conv([b.mf], 'tail.recon_block.0', out=b.u, relu=True)
            self._conv([b.u], 'tail.recon_block.2', epilogue=LV_EPI_PS4_NCHW, base_hr=b.base, **outs)
        else:
            p = f'body_{k - 1}.leg.recon_block'
            self._conv([fin], p + '.0', out=b.u, relu=True)
            self._conv([b.u], p + '.2', epilogue=LV_EPI_PS4_NCHW, base_hr=b.base, **outs)
        self._flush_chain(end=True)

    def forward(self, x, exit_leg=None, uint8=False):
        """x: fp32 NCHW [n,3,h,w] CUDA tensor on the 0..255 scale -> fp32 NCHW [n,3,4h,4w] (engine-owned buffer), or, with
        `uint8`, the PNG-ready uint8 frame clip(round(.)) written straight from the exit conv's epilogue (reference
        get_sr.py:86-89, validate.py:17-18): no fp32 frame is stored and no conversion kernel runs."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise LarvaNetB200Error(f'expected NCHW input with 3 channels, got {tuple(x.shape)}')
        n, _, h, w = (int(v) for v in x.shape)
        self.repack(backward=False)
        key = (n, h, w, exit_leg, bool(uint8))
        ent = self._infer.get(key, lambda: self._build_infer(n, h, w))
        b = ent['bufs']
        if uint8 and b.out_u8 is None:
            b.out_u8 = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.uint8, device=self.device)
        if b.row:
            self.repack_ky(backward=False)
        b.x.copy_(x.to(dtype=torch.float32), non_blocking=True)
        if n * h * w == 0:
            return b.out_u8 if uint8 else b.out
        _run_cached(self, ent, lambda: self._run_infer(b, exit_leg, uint8))
        return b.out_u8 if uint8 else b.out

    # ------------------------------------------------------------------ training
    def _build_train(self, n, h, w):
        b = _Bufs()
        dev = self.device
        b.x = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
        b.truth = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=dev)
        b.base = torch.empty_like(b.truth)
        # data parallel with the fused exchange+optimizer kernel: the local loss accumulator lives in the symmetric buffer
        # (peers read it), the all-reduced loss lands in peers.loss_out
        b.loss_sum = self._symm.loss_view if self._dp_fused() else torch.zeros(1, dtype=torch.float64, device=dev)
        b.f0 = self._act(n, h, w)
        b.t = [[self._act(n, h, w) for _ in range(nb)] for nb in self.blocks]       # post-ReLU of conv1
        b.a = [[self._act(n, h, w) for _ in range(nb - 1)] for nb in self.blocks]   # block outputs (not the last)
        b.feats = [self._act(n, h, w) for _ in range(self.m)]                       # body outputs
        b.u = [self._act(n, h, w) for _ in range(self.m)]                           # leg post-ReLU
        b.g = [self._act(n, h, w) for _ in range(self.m)]                           # leg sign gradient (dY of recon.2)
        b.du = [self._act(n, h, w) for _ in range(self.m)]                          # dY of leg recon.0
        b.dt = [[self._act(n, h, w) for _ in range(nb)] for nb in self.blocks]      # dY of block conv1
        b.da = [[self._act(n, h, w) for _ in range(nb)] for nb in self.blocks]      # dY of block conv2 (= d block out)
        b.dfin = [self._act(n, h, w) for _ in range(self.m)]                        # gradient wrt each body's input
        if self.v2:
            b.mf, b.ut, b.gt, b.dut, b.dmf = (self._act(n, h, w) for _ in range(5))
            b.dfeat = [self._act(n, h, w) for _ in range(self.m)]
        b.exits = None
        b.head_ws = torch.empty(int(_lib.load().lv_head_wgrad_workspace_bytes(C)), dtype=torch.uint8, device=dev)
        b.row = self.use_row_path(n, h, w)
        # weight-gradient batches, one per body (+ tail), in arena order
        numel = n * 3 * 16 * h * w
        denom = self.m + 1 if self.v2 else self.m
        b.scale = 1.0 / (float(numel) * denom)
        gv = self.arena.grad_views
        splits = self.wgrad_splits
        tiles = n * ((h + 15) // 16) * ((w + 7) // 8)
        per_body = []
        for i, nb in enumerate(self.blocks):
            items = []
            fin = b.f0 if i == 0 else b.feats[i - 1]
            for j in range(nb):
                p = f'body_{i}.res_blocks.{j}.body'
                a_in = fin if j == 0 else b.a[i][j - 1]
                items.append(dict(x=a_in, dy=b.dt[i][j], dw=gv[p + '.0.weight'], db=gv[p + '.0.bias']))
                items.append(dict(x=b.t[i][j], dy=b.da[i][j], dw=gv[p + '.2.weight'], db=gv[p + '.2.bias']))
            p = f'body_{i}.leg.recon_block'
            items.append(dict(x=b.feats[i], dy=b.du[i], dw=gv[p + '.0.weight'], db=gv[p + '.0.bias']))
            items.append(dict(x=b.u[i], dy=b.g[i], dw=gv[p + '.2.weight'], db=gv[p + '.2.bias']))
            per_body.append(items)
        if self.v2:
            items = []
            for s in range(self.m):
                items.append(dict(x=b.feats[s], dy=b.dmf, dw=gv['tail.merge_conv.weight'],
                                  db=gv['tail.merge_conv.bias'] if s == 0 else None, cin_total=C * self.m, cin_off=C * s))
            items.append(dict(x=b.mf, dy=b.dut, dw=gv['tail.recon_block.0.weight'], db=gv['tail.recon_block.0.bias']))
            items.append(dict(x=b.ut, dy=b.gt, dw=gv['tail.recon_block.2.weight'], db=gv['tail.recon_block.2.bias']))
            per_body.append(items)
        # All weight gradients run after the backward-data chain, so bodies can share launches.  Every launch pays a
        # fixed TMEM-drain + split reduction (~25 us) and runs one CTA per SM, so pick the grouping that minimises
        #   launches x (tiles per CTA x ~1.7 us + fixed)      [measured on B200: 1.66 us per 16x8 tile in steady state]
        best = None
        for k in range(1, len(per_body) + 1):
            groups = [sum(per_body[g:g + k], []) for g in range(0, len(per_body), k)]
            cost = 0.0
            for grp in groups:
                ctas = min(self.sm_count, len(grp) * tiles, len(grp) * splits if splits > 0 else self.sm_count)
                cost += -(-len(grp) * tiles // ctas) * 1.4 + 25.0
            if best is None or cost < best[0] - 1e-9:
                best = (cost, groups)
        for grp in best[1]:
            for it in grp:
                it['overwrite'] = True     # every conv weight / bias slice has exactly one writer per step
        b.wgrad = [self._make_wgrad(grp, tiles, splits) for grp in best[1]]
        return b

    def _make_wgrad(self, items, tiles, splits):
        if splits <= 0:
            # the library cuts the launch's tile jobs into min(splits * layers, SMs) equal ranges (1 CTA/SM): ask for
            # at least one range per SM
            splits = max(1, min(tiles, -(-self.sm_count // len(items))))
        return ops.WgradBatch(items, splits, self.device)

    def set_data_parallel(self, world_size, process_group=None):
        """Collective when world_size > 1 (every rank must call it)."""
        self.world_size = int(world_size)
        self.process_group = process_group
        self._train.clear()
        self._symm = None
        if self.world_size > 1 and self.device.type == 'cuda':
            from . import dist as lvdist
            self._symm = lvdist.SymmetricGradExchange.try_create(self.arena.total, self.device, process_group)
            if self._symm is not None:
                self.arena.replace_grad(self._symm.buffer)

    def _dp_fused(self):
        """True when the data-parallel exchange is fused into the optimizer kernel (peer-mapped arenas on every rank, a
        FusedAdamW attached, bf16 fused-update layout)."""
        return (self.world_size > 1 and self._symm is not None and self._symm.peers is not None
                and self._dp_optim is not None and self.fused_update_available()
                and os.environ.get('LARVANET_B200_DP_FUSED', '1') != '0')

    def _run_train(self, b):
        """forward with saved activations + fused losses, then backward-data chain and batched weight gradients."""
        self._row_active = b.row
        hw, hb = self._w('head.feature_extraction')
        scale = b.scale
        b.loss_sum.zero_()
        if self.simt or self.act_dtype != torch.bfloat16:
            self.arena.grad.zero_()           # the CUDA-core weight-gradient kernels accumulate with atomics
        # (bf16: the tensor-core weight-gradient reduction and the head gradient's second pass STORE their results --
        # every gradient has exactly one writer per step -- so nothing needs zeroing)
        ops.head_bicubic(b.x, hw, hb, b.f0, b.base)
        self._begin_chain()
        # ---------------- forward ----------------
        fin = b.f0
        for i, nb in enumerate(self.blocks):
            a = fin
            for j in range(nb):
                p = f'body_{i}.res_blocks.{j}.body'
                self._conv([a], p + '.0', out=b.t[i][j], relu=True)
                last = j == nb - 1
                dst = b.feats[i] if last else b.a[i][j]
                self._conv([b.t[i][j]], p + '.2', out=dst, res1=a, res2=fin if last else None)
                a = dst
            fin = b.feats[i]
            p = f'body_{i}.leg.recon_block'
            self._conv([fin], p + '.0', out=b.u[i], relu=True)
            self._conv([b.u[i]], p + '.2', epilogue=LV_EPI_PS4_NCHW, base_hr=b.base, truth_hr=b.truth,
                       loss_sum=b.loss_sum, grad_sign=b.g[i],
                       out_hr=b.exits[i] if b.exits is not None else None)
        if self.v2:
            self._conv(b.feats, 'tail.merge_conv', out=b.mf)
            self._conv([b.mf], 'tail.recon_block.0', out=b.ut, relu=True)
            self._conv([b.ut], 'tail.recon_block.2', epilogue=LV_EPI_PS4_NCHW, base_hr=b.base, truth_hr=b.truth,
                       loss_sum=b.loss_sum, grad_sign=b.gt,
                       out_hr=b.exits[self.m] if b.exits is not None else None)
        # ---------------- backward ----------------
        # (the weight-gradient batches only read saved activations and dY buffers, so they run after the whole
        # backward-data chain)
        if self.v2:
            self._dgrad(b.gt, 'tail.recon_block.2', out=b.dut, mask=b.ut)
            self._dgrad(b.dut, 'tail.recon_block.0', out=b.dmf)
            for s in range(self.m):
                self._dgrad(b.dmf, 'tail.merge_conv', out=b.dfeat[s], s=s)
        dnext = None
        for i in reversed(range(self.m)):
            nb = self.blocks[i]
            p = f'body_{i}.leg.recon_block'
            self._dgrad(b.g[i], p + '.2', out=b.du[i], mask=b.u[i])
            # gradient wrt the body output = leg branch + next body's input gradient (+ tail branch)
            dfout = b.da[i][nb - 1]
            self._dgrad(b.du[i], p + '.0', out=dfout, res1=dnext, res2=b.dfeat[i] if self.v2 else None)
            for j in reversed(range(nb)):
                pj = f'body_{i}.res_blocks.{j}.body'
                self._dgrad(b.da[i][j], pj + '.2', out=b.dt[i][j], mask=b.t[i][j])
                dst = b.dfin[i] if j == 0 else b.da[i][j - 1]
                self._dgrad(b.dt[i][j], pj + '.0', out=dst, res1=b.da[i][j], res2=dfout if j == 0 else None)
            dnext = b.dfin[i]
        self._flush_chain(end=True)
        # (tried: the head conv's weight gradient on a side stream next to the persistent weight-gradient kernel -- the
        # step got SLOWER, 0.608 vs 0.517 ms: the small kernel's blocks delay the persistent kernel's one-CTA-per-SM wave)
        for wb in b.wgrad:
            wb.launch(simt=self.simt)
        ops.head_wgrad(b.x, b.dfin[0], self.arena.grad_views['head.feature_extraction.weight'],
                       self.arena.grad_views['head.feature_extraction.bias'], scale,
                       overwrite=not (self.simt or self.act_dtype != torch.bfloat16), workspace=b.head_ws)

    def train_step(self, x, truth, keep_exits=False):
        """One forward+backward.  Leaves d(loss)/d(param) in the gradient arena (== every param.grad) and returns the
        multi-exit loss as a lazily read `DeviceScalar` (`.item()` / `float()`; no host sync, no extra launch before that).  Restates models/LarvaNet.py:102-113 /
        models/LarvaNetV2.py:105-120 (everything before optim.step())."""
        n, _, h, w = (int(v) for v in x.shape)
        if tuple(truth.shape) != (n, 3, 4 * h, 4 * w):
            raise LarvaNetB200Error(f'truth shape {tuple(truth.shape)} does not match 4x input {tuple(x.shape)}')
        key = (n, h, w, bool(keep_exits))

        def build():
            b = self._build_train(n, h, w)
            if keep_exits:
                b.exits = [torch.empty_like(b.truth) for _ in range(self.m + (1 if self.v2 else 0))]
            # mean over the GLOBAL batch: 1 / (HR samples on all ranks x exits).  Ranks may hold different numbers of
            # patches (dist.shard_range with total % world != 0), so the global count is all-reduced once per shape
            # (collective: every rank builds its buffers for a new shape in the same step).
            b.global_numel = float(n * 3 * 16 * h * w)
            if self.world_size > 1:
                import torch.distributed as tdist
                cnt = torch.tensor([b.global_numel], dtype=torch.float64, device=self.device)
                tdist.all_reduce(cnt, op=tdist.ReduceOp.SUM, group=self.process_group)
                b.global_numel = float(cnt.item())
            denom = self.m + 1 if self.v2 else self.m
            b.scale = 1.0 / (b.global_numel * denom)
            for wb in b.wgrad:
                wb.set_scale(b.scale)
            return b

        ent = self._train.get(key, build)
        b = ent['bufs']
        b.x.copy_(x, non_blocking=True)          # the GPU starts on the input copies while the host checks the weights
        b.truth.copy_(truth, non_blocking=True)
        self.repack(backward=True)
        if b.row:
            self.repack_ky(backward=True)
        _run_cached(self, ent, lambda: self._run_train(b))
        self.arena.attach_grads()   # host-only book-keeping, after the launches so that the GPU is already busy
        self.last_exits = b.exits
        self._last_train = b
        if self.world_size > 1 and self._dp_fused():
            # the exchange happens inside the optimizer's kernel (ops.dp_adamw_pack_step): gradients AND the loss are
            # global only after `optim.step()`; the returned scalar reads the all-reduced loss lazily
            self._dp_pending = True
            return DeviceScalar(self._symm.peers.loss_out, b.scale)
        if self.world_size > 1:
            from . import dist as lvdist
            if self._symm is not None:
                # 8-byte loss on NCCL's stream, the gradients through peer memory on this one, concurrently
                works = lvdist.allreduce_gradients(b.loss_sum, None, self.process_group)
                self._symm.allreduce_()
            else:
                works = lvdist.allreduce_gradients(self.arena.grad, b.loss_sum, self.process_group)
            for work in works:
                work.wait()   # stream-ordered on NCCL: enqueues a wait on the current stream, no host sync
        return DeviceScalar(b.loss_sum, b.scale)

    def saved_activations(self):
        """The forward activations the last train_step saved for its backward pass, as NCHW float32 CPU tensors keyed
        like oracle.larva_oracle.larvanet_train_step(tapes_from=...).  For parity tests / debugging only."""
        b = self._last_train

        def cv(a):   # planar-8 [n,h,c/8,w,8] -> NCHW
            n, h, ch, w, _ = a.shape
            return a.detach().to(torch.float32).permute(0, 2, 4, 1, 3).reshape(n, ch * 8, h, w).contiguous().cpu().numpy()

        t = {'f0': cv(b.f0)}
        for i, nb in enumerate(self.blocks):
            t[('feat', i)] = cv(b.feats[i])
            t[('u', i)] = cv(b.u[i])
            for j in range(nb):
                t[('t', i, j)] = cv(b.t[i][j])
                if j > 0:
                    t[('a', i, j)] = cv(b.a[i][j - 1])
        if self.v2:
            t['mf'], t['ut'] = cv(b.mf), cv(b.ut)
        return t

    # ------------------------------------------------------------------ individually-callable modules
    # The reference's train step and analysis scripts call sub-modules one by one with NCHW fp32 tensors
    # (models/LarvaNet.py:102-107, validate_tree.py:94-96).  These helpers keep that surface: convert at the
    # boundary, run the same kernels eagerly.  Forward only (no autograd graph is recorded).
    def _to_act(self, x_nchw):
        self._row_active = False   # individually-called modules run per-layer on the tile kernel
        x = x_nchw.detach().to(device=self.device, dtype=torch.float32).contiguous()
        n, c, h, w = (int(v) for v in x.shape)
        a = ops.act_empty(n, h, w, c, self.act_dtype, self.device)
        ops.nchw_to_nhwc(x, a)
        return a

    def _to_nchw(self, a):
        n, h, w, c = ops.act_dims(a)
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=self.device)
        ops.nhwc_to_nchw(a, y)
        return y

    def run_head(self, x):
        self.repack()
        x = x.detach().to(device=self.device, dtype=torch.float32).contiguous()
        n, _, h, w = (int(v) for v in x.shape)
        hw, hb = self._w('head.feature_extraction')
        fea = self._act(n, h, w)
        ops.head_bicubic(x, hw, hb, fea, None)
        return self._to_nchw(fea)

    def run_base(self, x):
        x = x.detach().to(device=self.device, dtype=torch.float32).contiguous()
        n, c, h, w = (int(v) for v in x.shape)
        out = torch.empty((n, c, 4 * h, 4 * w), dtype=torch.float32, device=self.device)
        ops.bicubic_x4(x, out)
        return out

    def _resblock(self, i, j, a, fin=None):
        n, h, w, _ = ops.act_dims(a)
        p = f'body_{i}.res_blocks.{j}.body'
        t, o = self._act(n, h, w), self._act(n, h, w)
        self._conv([a], p + '.0', out=t, relu=True)
        self._conv([t], p + '.2', out=o, res1=a, res2=fin)
        return o

    def run_resblock(self, i, j, x):
        self.repack()
        return self._to_nchw(self._resblock(i, j, self._to_act(x)))

    def run_body(self, i, x):
        self.repack()
        fin = self._to_act(x)
        a = fin
        nb = self.blocks[i]
        for j in range(nb):
            a = self._resblock(i, j, a, fin if j == nb - 1 else None)
        return self._to_nchw(a)

    def _recon(self, prefix, fea, base):
        n, h, w, _ = ops.act_dims(fea)
        base = base.detach().to(device=self.device, dtype=torch.float32).contiguous()
        u = self._act(n, h, w)
        out = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=self.device)
        self._conv([fea], prefix + '.recon_block.0', out=u, relu=True)
        self._conv([u], prefix + '.recon_block.2', epilogue=LV_EPI_PS4_NCHW, out_hr=out, base_hr=base)
        return out

    def run_leg(self, i, fea, base):
        self.repack()
        return self._recon(f'body_{i}.leg', self._to_act(fea), base)

    def run_tail(self, features, base):
        self.repack()
        feats = [self._to_act(f) for f in features]
        n, h, w, _ = ops.act_dims(feats[0])
        mf = self._act(n, h, w)
        self._conv(feats, 'tail.merge_conv', out=mf)
        return self._recon('tail', mf, base)


class EdsrEngine:
    """EDSR-baseline x2/x4 inference on the same conv kernels (reference models/edsr.py:177-207, BASELINE config 3):
    1x1 mean_shift + first_conv in the head kernel, 2 convs per ResidualBlock (res_weight folded into the epilogue's
    res_scale), after_res_conv with the global skip as residual, conv(F->4F) with a PixelShuffle(2) epilogue per
    upsampling stage, final_conv(F->3) with the 1x1 mean_inverse_shift fused into its epilogue.  Inference only."""

    def __init__(self, module, features, num_res_blocks, res_weight, scale, act_dtype=torch.bfloat16, device=None,
                 use_graphs=None):
        self.module = module
        self.f = int(features)
        self.nb = int(num_res_blocks)
        self.res_weight = float(res_weight)
        self.scale = int(scale)
        if self.scale not in (2, 4, 8):
            raise LarvaNetB200Error(f'EDSR scale {scale}: only the PixelShuffle(2) stages (x2/x4/x8) have a kernel')
        self.nup = {2: 1, 4: 2, 8: 3}[self.scale]
        self.act_dtype = act_dtype
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.sm_count = ops.device_check(self.device.index)
        if act_dtype == torch.bfloat16 and self.f % 16 != 0:
            raise LarvaNetB200Error(f'--edsr_conv_features={self.f}: the bf16 tensor-core path needs a multiple of 16 '
                                    '(use --precision=fp32 for other widths)')
        if use_graphs is None:
            use_graphs = os.environ.get('LARVANET_B200_GRAPHS', '1') != '0'
        self.use_graphs = use_graphs
        self.arena = ParamArena(module, self.device)
        self.simt = False
        self.replayed_launches = 0
        f = self.f
        self._layers = [(f'res_blocks.{j}.body.{k}', f, f) for j in range(self.nb) for k in (0, 2)]
        self._layers.append(('after_res_conv', f, f))
        self._layers += [(f'upsample.body.{2 * s}', 4 * f, f) for s in range(self.nup)]
        self._layers.append(('final_conv', 3, f))
        sizes = [(p, ops.packed_weight_bytes(o, i, act_dtype)) for p, o, i in self._layers]
        total = sum((b + 255) // 256 * 256 for _, b in sizes)
        self._packed = torch.zeros(total, dtype=torch.uint8, device=self.device)
        self._pk, off = {}, 0
        for p, b in sizes:
            self._pk[p] = self._packed[off:off + b]
            off += (b + 255) // 256 * 256
        self._pack_items = [dict(w=self.arena.views[p + '.weight'], packed=self._pk[p], transpose=0, i_off=0, i_cnt=i,
                                 cin=i, dtype=act_dtype) for p, o, i in self._layers]
        # the F -> F convs of the body (32 resblock convs + after_res_conv) as ONE row-marching chain launch when the frame
        # is large enough (csrc/conv_row.cu: 9 MMAs of N = 192 per 128 px instead of 27 of N = 64)
        self._body = [p for p, o, i in self._layers if o == f and i == f]
        self._row_ok = act_dtype == torch.bfloat16 and f in (48, 64)
        if self._row_ok:
            nb_ = ops.packed_weight_bytes(f, f, act_dtype)
            step = (nb_ + 255) // 256 * 256
            self._packed_ky = torch.zeros(step * len(self._body), dtype=torch.uint8, device=self.device)
            for k, p in enumerate(self._body):
                self._pk[(p, 'ky')] = self._packed_ky[k * step:k * step + nb_]
            self._pack_items += [dict(w=self.arena.views[p + '.weight'], packed=self._pk[(p, 'ky')], transpose=0, i_off=0,
                                      i_cnt=f, cin=f, dtype=act_dtype, wlayout=_lib.LV_W_KY_STACKED) for p in self._body]
            if f == 64:
                # the last conv (64 -> 3 at the output resolution) on the row kernel as well: 12 MMAs of N = 48 per 128
                # pixels instead of 36 of N = 16 on the tile kernel
                self._pk[('final_conv', 'ky')] = torch.zeros(ops.packed_weight_bytes(3, f, act_dtype), dtype=torch.uint8,
                                                            device=self.device)
                self._pack_items.append(dict(w=self.arena.views['final_conv.weight'], packed=self._pk[('final_conv', 'ky')],
                                             transpose=0, i_off=0, i_cnt=f, cin=f, dtype=act_dtype,
                                             wlayout=_lib.LV_W_KY_STACKED))
        # (measured at 1 x 270x480: 1.78 ms with the row chain vs 1.93 ms with 33 per-layer launches of the tile kernel)
        self.row_min_pixels = int(os.environ.get('LARVANET_B200_ROW_MIN_PIXELS', str(96 * 1024)))
        self._chain_ws = {}
        self._packed_version = None
        self._infer = _ShapeCache(int(os.environ.get('LARVANET_B200_SHAPE_CACHE', '4')))

    def use_row_path(self, n, h, w):
        if not self._row_ok or self.simt:
            return False
        mode = os.environ.get('LARVANET_B200_ROW', 'auto')
        if mode in ('0', '1'):
            return mode == '1'
        return n * h * w >= self.row_min_pixels

    def repack(self, force=False):
        ver = self.arena.version()
        if force or ver != self._packed_version:
            ops.pack_weights(self._pack_items)
            self._packed_version = ver

    def mark_weights_changed(self):
        self._packed_version = None

    def _conv(self, src, prefix, cout, **kw):
        ops.conv3x3([src], self._pk[prefix], cout, bias=self.arena.views[prefix + '.bias'], simt=self.simt, **kw)

    def _build(self, n, h, w):
        b = _Bufs()
        dev, dt, f = self.device, self.act_dtype, self.f
        b.x = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
        b.x0 = ops.act_empty(n, h, w, f, dt, dev)
        b.t = torch.empty_like(b.x0)
        b.pp = [torch.empty_like(b.x0), torch.empty_like(b.x0)]
        b.up = [ops.act_empty(n, h << (s + 1), w << (s + 1), f, dt, dev) for s in range(self.nup)]
        b.out = torch.empty((n, 3, h * self.scale, w * self.scale), dtype=torch.float32, device=dev)
        return b

    def _run(self, b):
        v = self.arena.views
        f = self.f
        ops.head_bicubic(b.x, v['first_conv.weight'], v['first_conv.bias'], b.x0, None,
                         pre_w=v['mean_shift.weight'], pre_b=v['mean_shift.bias'])
        a = b.x0
        skip = b.pp[self.nb & 1]
        if self.use_row_path(*ops.act_dims(b.x0)[:3]):
            KY = _lib.LV_W_KY_STACKED
            mk = lambda src, p, **kw: ops.make_conv_args([src], self._pk[(p, 'ky')], f, bias=v[p + '.bias'], wlayout=KY, **kw)
            chain = []
            for j in range(self.nb):
                p = f'res_blocks.{j}.body'
                chain.append(mk(a, p + '.0', out=b.t, relu=True))
                dst = b.pp[j & 1]
                chain.append(mk(b.t, p + '.2', out=dst, res1=a, res_scale=self.res_weight))
                a = dst
            chain.append(mk(a, 'after_res_conv', out=skip, res1=b.x0))
            n_, h_, w_, _ = ops.act_dims(b.x0)
            if (n_, h_, w_) not in self._chain_ws:
                self._chain_ws[(n_, h_, w_)] = ops.chain_workspace(n_, h_, w_, self.device)
            ops.conv3x3_chain(chain, self._chain_ws[(n_, h_, w_)])
        else:
            for j in range(self.nb):
                p = f'res_blocks.{j}.body'
                self._conv(a, p + '.0', f, out=b.t, relu=True)
                dst = b.pp[j & 1]
                self._conv(b.t, p + '.2', f, out=dst, res1=a, res_scale=self.res_weight)
                a = dst
            self._conv(a, 'after_res_conv', f, out=skip, res1=b.x0)
        a = skip
        for s in range(self.nup):
            self._conv(a, f'upsample.body.{2 * s}', 4 * f, out=b.up[s], epilogue=_lib.LV_EPI_PS2_NHWC)
            a = b.up[s]
        if ('final_conv', 'ky') in self._pk and self.use_row_path(*ops.act_dims(a)[:3]) and ops.act_dims(a)[2] >= 129:
            ops.conv3x3([a], self._pk[('final_conv', 'ky')], 3, bias=v['final_conv.bias'], wlayout=_lib.LV_W_KY_STACKED,
                        epilogue=_lib.LV_EPI_RGB_NCHW, out_hr=b.out, post_w=v['mean_inverse_shift.weight'],
                        post_b=v['mean_inverse_shift.bias'])
        else:
            self._conv(a, 'final_conv', 3, epilogue=_lib.LV_EPI_RGB_NCHW, out_hr=b.out,
                       post_w=v['mean_inverse_shift.weight'], post_b=v['mean_inverse_shift.bias'])

    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != 3:
            raise LarvaNetB200Error(f'expected NCHW input with 3 channels, got {tuple(x.shape)}')
        n, _, h, w = (int(t) for t in x.shape)
        self.repack()
        key = (n, h, w)
        ent = self._infer.get(key, lambda: self._build(n, h, w))
        b = ent['bufs']
        b.x.copy_(x.to(dtype=torch.float32), non_blocking=True)
        if n * h * w == 0:
            return b.out
        _run_cached(self, ent, lambda: self._run(b))
        return b.out
