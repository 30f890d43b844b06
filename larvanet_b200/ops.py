"""Tensor-level wrappers over the C-ABI: validate torch tensors, pass raw device pointers + the current stream.

Activation tensors use the library's planar-8 layout and have torch shape [n, h, c/8, w, 8] (see `act_empty`).

PyTorch is only the owner of device memory and streams here; every function ends in exactly one `lv_*` call.
No function in this file has a fallback path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (LV_BF16, LV_EPI_NHWC, LV_EPI_PS2_NHWC, LV_EPI_PS4_NCHW, LV_EPI_RGB_NCHW, LV_F32, ConvArgs,
                   PackItem, WgradItem, check)

_TORCH2LV = {torch.float32: LV_F32, torch.bfloat16: LV_BF16}


def dtype_id(dt):
    try:
        return _TORCH2LV[dt]
    except KeyError:
        raise _lib.LarvaNetB200Error(f'unsupported activation dtype {dt}; use torch.bfloat16 or torch.float32') from None


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream():
    # the raw-handle query is ~10x cheaper than building a torch.cuda.Stream object (this runs on every launch)
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, name='tensor'):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.LarvaNetB200Error(f'{name} must be a CUDA tensor (no CPU fallback)')
    if not t.is_contiguous():
        raise _lib.LarvaNetB200Error(f'{name} must be contiguous')
    if dtype is not None and t.dtype != dtype:
        raise _lib.LarvaNetB200Error(f'{name} must be {dtype}, got {t.dtype}')
    return t.data_ptr()


def act_empty(n, h, w, c, dtype, device):
    """Allocate an activation tensor in the library's planar-8 layout [n][h][c/8][w][8]."""
    if c % 8 != 0:
        raise _lib.LarvaNetB200Error(f'activation tensors need channels % 8 == 0 (got {c})')
    return torch.empty((n, h, c // 8, w, 8), dtype=dtype, device=device)


def act_dims(t):
    """(n, h, w, c) of a planar-8 activation tensor."""
    if t.dim() != 5 or t.shape[4] != 8:
        raise _lib.LarvaNetB200Error(f'expected a planar-8 activation tensor [n,h,c/8,w,8], got shape {tuple(t.shape)}')
    n, h, ch, w, _ = (int(v) for v in t.shape)
    return n, h, w, ch * 8


def device_check(device_index=None):
    lib = _lib.load()
    dev = torch.cuda.current_device() if device_index is None else device_index
    sm = C.c_int(0)
    check(lib.lv_device_check(dev, C.byref(sm)), 'lv_device_check')
    return sm.value


def packed_weight_bytes(cout, cin_total, dt):
    return int(_lib.load().lv_packed_weight_bytes(cout, cin_total, dtype_id(dt)))


def build_pack_arrays(items):
    """Marshal pack items once: list of (ctypes PackItem array, count) chunks of <= 128 (the C-ABI's per-call limit).
    items: list of dicts(w=fp32 OIHW tensor, packed=uint8/any tensor, transpose, i_off, i_cnt, cin, dtype, wlayout).
    The tensors must stay alive (and keep their storage) for as long as the arrays are used."""
    out = []
    for start in range(0, len(items), 128):
        chunk = items[start:start + 128]
        arr = (PackItem * len(chunk))()
        for k, it in enumerate(chunk):
            w = it['w']
            O, I = int(w.shape[0]), int(w.shape[1])
            assert w.shape[2] == 3 and w.shape[3] == 3
            arr[k].w = _ptr(w, torch.float32, 'weight')
            arr[k].packed = _ptr(it['packed'], None, 'packed')
            arr[k].O, arr[k].I = O, I
            arr[k].transpose = int(it.get('transpose', 0))
            arr[k].i_off = int(it.get('i_off', 0))
            arr[k].i_cnt = int(it.get('i_cnt', I))
            arr[k].cin = int(it['cin'])
            arr[k].dtype = dtype_id(it['dtype'])
            arr[k].wlayout = int(it.get('wlayout', 0))
        out.append((arr, len(chunk)))
    return out


def pack_weights_prebuilt(arrays):
    lib = _lib.load()
    for arr, cnt in arrays:
        check(lib.lv_pack_conv3x3_weights(arr, cnt, _stream()), 'lv_pack_conv3x3_weights')


def pack_weights(items):
    pack_weights_prebuilt(build_pack_arrays(items))


def make_conv_args(srcs, weights, cout, *, bias=None, relu=False, mask=None, res1=None, res2=None, out=None,
                   epilogue=LV_EPI_NHWC, out_hr=None, base_hr=None, truth_hr=None, loss_sum=None, grad_sign=None,
                   post_w=None, post_b=None, res_scale=1.0, wlayout=0, out_u8=None):
    """Build an lv_conv_args from torch tensors.  srcs: list of planar-8 activation tensors (same shape/dtype)."""
    x0 = srcs[0]
    dt = x0.dtype
    n, h, w, cin = act_dims(x0)
    a = ConvArgs()
    a.n, a.h, a.w, a.cin, a.num_src, a.cout = n, h, w, cin, len(srcs), int(cout)
    a.dtype = dtype_id(dt)
    a.relu = int(bool(relu))
    a.epilogue = int(epilogue)
    a.res_scale = float(res_scale)
    a.wlayout = int(wlayout)
    for i, s in enumerate(srcs):
        if tuple(s.shape) != tuple(x0.shape):
            raise _lib.LarvaNetB200Error('all conv sources must have the same shape')
        a.src[i] = _ptr(s, dt, f'src[{i}]')
    a.weights = _ptr(weights, None, 'weights')
    a.bias = _ptr(bias, torch.float32, 'bias')
    a.mask = _ptr(mask, dt, 'mask')
    a.res1 = _ptr(res1, dt, 'res1')
    a.res2 = _ptr(res2, dt, 'res2')
    a.out = _ptr(out, dt, 'out')
    a.out_hr = _ptr(out_hr, torch.float32, 'out_hr')
    a.base_hr = _ptr(base_hr, torch.float32, 'base_hr')
    a.truth_hr = _ptr(truth_hr, torch.float32, 'truth_hr')
    a.loss_sum = _ptr(loss_sum, torch.float64, 'loss_sum')
    a.grad_sign = _ptr(grad_sign, dt, 'grad_sign')
    a.post_w = _ptr(post_w, torch.float32, 'post_w')
    a.post_b = _ptr(post_b, torch.float32, 'post_b')
    a.out_u8 = _ptr(out_u8, torch.uint8, 'out_u8')
    return a


# bench.py's instrumented pass: when a list is installed here every conv launch is bracketed by CUDA events on the
# launching stream and (event0, event1, flops, tag) is appended
CONV_TIMERS = None


def conv3x3_launch(args, max_ctas=0, simt=False):
    lib = _lib.load()
    if CONV_TIMERS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.lv_conv3x3(C.byref(args), int(max_ctas), _stream()), 'lv_conv3x3')
        e1.record()
        flops = 2.0 * 9 * args.cin * args.num_src * args.cout * args.n * args.h * args.w
        CONV_TIMERS.append((e0, e1, flops, (args.cin * args.num_src, args.cout, args.epilogue)))
        return
    if simt:
        check(lib.lv_conv3x3_simt(C.byref(args), _stream()), 'lv_conv3x3_simt')
    else:
        check(lib.lv_conv3x3(C.byref(args), int(max_ctas), _stream()), 'lv_conv3x3')


def conv3x3(srcs, weights, cout, max_ctas=0, simt=False, **kw):
    conv3x3_launch(make_conv_args(srcs, weights, cout, **kw), max_ctas, simt)


def chain_workspace(n, h, w, device):
    """Zeroed data-flow flag workspace for `conv3x3_chain` on activations of n x h x w (one per stream)."""
    nbytes = int(_lib.load().lv_conv_chain_workspace_bytes(int(n), int(h), int(w)))
    return torch.zeros(max(nbytes // 4, 1), dtype=torch.int32, device=device)


def chain_eligible(args):
    """True when `args` (ConvArgs) can be a layer of `conv3x3_chain`."""
    if args.dtype != _lib.LV_BF16 or args.num_src != 1 or args.cin != args.cout:
        return False
    if args.wlayout == _lib.LV_W_KY_STACKED:      # row-marching chain: 48 or 64 channels, planar or PixelShuffle(4) epilogue
        return args.cin in (48, 64) and (args.epilogue == _lib.LV_EPI_NHWC or
                                         (args.epilogue == _lib.LV_EPI_PS4_NCHW and args.cin == 48))
    return args.cin == 48


def conv3x3_chain(arg_list, workspace, max_ctas=0):
    """Run the convs described by `arg_list` (ConvArgs from make_conv_args, in order) as ONE persistent launch."""
    lib = _lib.load()
    n = len(arg_list)
    if n == 0:
        return
    pos = 0
    while pos < n:   # the C-ABI takes at most LV_CHAIN_MAX_LAYERS layers per launch
        cnt = min(_lib.LV_CHAIN_MAX_LAYERS, n - pos)
        arr = (_lib.ConvArgs * cnt)(*arg_list[pos:pos + cnt])
        if CONV_TIMERS is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(lib.lv_conv3x3_chain(arr, cnt, workspace.data_ptr(), workspace.numel() * 4, int(max_ctas), _stream()),
              'lv_conv3x3_chain')
        if CONV_TIMERS is not None:
            e1.record()
            a0 = arr[0]
            CONV_TIMERS.append((e0, e1, 2.0 * 9 * a0.cin * a0.cout * a0.n * a0.h * a0.w * cnt, ('chain', cnt, a0.wlayout)))
        pos += cnt


def head_bicubic(x, w, b, fea, base_hr=None, pre_w=None, pre_b=None):
    """x fp32 NCHW [n,3,h,w] -> fea NHWC [n,h,w,cout] (+ base_hr fp32 NCHW [n,3,4h,4w])."""
    n, c, h, wd = (int(v) for v in x.shape)
    assert c == 3
    cout = int(w.shape[0])
    check(_lib.load().lv_head_bicubic_fwd(
        _ptr(x, torch.float32, 'x'), _ptr(w, torch.float32, 'head weight'), _ptr(b, torch.float32, 'head bias'),
        _ptr(pre_w, torch.float32, 'pre_w'), _ptr(pre_b, torch.float32, 'pre_b'), _ptr(fea, None, 'fea'),
        _ptr(base_hr, torch.float32, 'base_hr'), n, h, wd, cout, dtype_id(fea.dtype), _stream()), 'lv_head_bicubic_fwd')


def bicubic_x4(x, out):
    n, c, h, w = (int(v) for v in x.shape)
    check(_lib.load().lv_bicubic_x4(_ptr(x, torch.float32, 'x'), _ptr(out, torch.float32, 'out'), n, c, h, w, _stream()),
          'lv_bicubic_x4')


_HEAD_WS = {}


def head_wgrad(x, dy, dw, db, scale, overwrite=False, workspace=None):
    """Head conv weight / bias gradient (deterministic two-pass reduction).  `overwrite`: store instead of accumulate."""
    n, c, h, w = (int(v) for v in x.shape)
    cout = act_dims(dy)[3]
    if workspace is None:      # per-device scratch for callers that do not manage one (tests, module-level calls)
        key = (x.device.index, cout)
        if key not in _HEAD_WS:
            _HEAD_WS[key] = torch.empty(int(_lib.load().lv_head_wgrad_workspace_bytes(cout)), dtype=torch.uint8, device=x.device)
        workspace = _HEAD_WS[key]
    check(_lib.load().lv_head_wgrad(_ptr(x, torch.float32, 'x'), _ptr(dy, None, 'dy'), _ptr(dw, torch.float32, 'dw'),
                                    _ptr(db, torch.float32, 'db'), n, h, w, cout, dtype_id(dy.dtype), float(scale),
                                    workspace.data_ptr(), workspace.numel(), int(bool(overwrite)), _stream()), 'lv_head_wgrad')


class WgradBatch:
    """A prepared batch of weight-gradient items: host ctypes array + its device copy + split-K workspace."""

    def __init__(self, items, splits, device):
        lib = _lib.load()
        self.count = len(items)
        self.splits = int(splits)
        self.host = (WgradItem * self.count)()
        self._keep = []
        for k, it in enumerate(items):
            x, dy, dw, db = it['x'], it['dy'], it['dw'], it.get('db')
            n, h, w, cin = act_dims(x)
            hi = self.host[k]
            hi.n, hi.h, hi.w, hi.cin, hi.cout = n, h, w, cin, act_dims(dy)[3]
            hi.cin_total = int(it.get('cin_total', cin))
            hi.cin_off = int(it.get('cin_off', 0))
            hi.dtype = dtype_id(x.dtype)
            hi.x, hi.dy = _ptr(x, None, 'x'), _ptr(dy, x.dtype, 'dy')
            hi.dw, hi.db = _ptr(dw, torch.float32, 'dw'), _ptr(db, torch.float32, 'db')
            hi.scale = float(it.get('scale', 1.0))
            hi.overwrite = int(bool(it.get('overwrite', False)))
            self._keep.append((x, dy, dw, db))
        raw = bytes(self.host)
        self.dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        ws = int(lib.lv_wgrad_workspace_bytes(self.host, self.count, self.splits))
        self.workspace = torch.empty(max(ws, 16), dtype=torch.uint8, device=device)

    def set_scale(self, scale):
        """Rewrite every item's scale (host + device copies)."""
        for k in range(self.count):
            self.host[k].scale = float(scale)
        self.dev.copy_(torch.frombuffer(bytearray(bytes(self.host)), dtype=torch.uint8), non_blocking=False)

    def launch(self, simt=False):
        lib = _lib.load()
        if simt:
            check(lib.lv_conv3x3_wgrad_simt(self.host, self.dev.data_ptr(), self.count, self.splits, _stream()),
                  'lv_conv3x3_wgrad_simt')
        else:
            check(lib.lv_conv3x3_wgrad(self.host, self.dev.data_ptr(), self.count, self.splits,
                                       self.workspace.data_ptr(), _stream()), 'lv_conv3x3_wgrad')


def nchw_to_nhwc(src, dst):
    n, c, h, w = (int(v) for v in src.shape)
    check(_lib.load().lv_nchw_to_nhwc(_ptr(src, torch.float32, 'src'), _ptr(dst, None, 'dst'), n, c, h, w,
                                      dtype_id(dst.dtype), _stream()), 'lv_nchw_to_nhwc')


def nhwc_to_nchw(src, dst):
    n, h, w, c = act_dims(src)
    check(_lib.load().lv_nhwc_to_nchw(_ptr(src, None, 'src'), _ptr(dst, torch.float32, 'dst'), n, c, h, w,
                                      dtype_id(src.dtype), _stream()), 'lv_nhwc_to_nchw')


def image_to_uint8(src, dst=None):
    """clip(round(src), 0, 255) as uint8 (validate._image_to_uint8) on the device; src: fp32 CUDA tensor."""
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.uint8, device=src.device)
    check(_lib.load().lv_image_to_uint8(_ptr(src, torch.float32, 'src'), _ptr(dst, torch.uint8, 'dst'), src.numel(),
                                        _stream()), 'lv_image_to_uint8')
    return dst


def psnr_sqsum(out_chw, truth_chw, sq_sum):
    """sq_sum (float64[1]) += sum (u8(truth) - u8(out))^2 over out's [c,h,w]; truth may be larger (cropped)."""
    c, h, w = (int(v) for v in out_chw.shape)
    tc, th, tw = (int(v) for v in truth_chw.shape)
    if tc != c:
        raise _lib.LarvaNetB200Error(f"psnr: channel mismatch {tc} vs {c}")
    check(_lib.load().lv_psnr_sqsum(_ptr(out_chw, torch.float32, 'out'), _ptr(truth_chw, torch.float32, 'truth'),
                                    _ptr(sq_sum, torch.float64, 'sq_sum'), c, h, w, th, tw, _stream()), 'lv_psnr_sqsum')


def crop_augment(items_dev, count, out_lr, out_hr, patch, scale):
    """Device-side crop + rot90 + flip: `items_dev` = uint8 CUDA tensor holding `count` lv_patch_item records."""
    check(_lib.load().lv_crop_augment(items_dev.data_ptr(), int(count), _ptr(out_lr, torch.float32, 'out_lr'),
                                      _ptr(out_hr, torch.float32, 'out_hr'), int(patch), int(scale), _stream()), 'lv_crop_augment')


def l1_loss_grad(out_hr, truth_hr, loss_sum, grad_sign=None):
    n, c, h4, w4 = (int(v) for v in out_hr.shape)
    dt = dtype_id(grad_sign.dtype) if grad_sign is not None else LV_F32
    check(_lib.load().lv_l1_loss_grad(_ptr(out_hr, torch.float32, 'out'), _ptr(truth_hr, torch.float32, 'truth'),
                                      _ptr(loss_sum, torch.float64, 'loss_sum'), _ptr(grad_sign, None, 'grad_sign'),
                                      n, c, h4 // 4, w4 // 4, dt, _stream()), 'lv_l1_loss_grad')


def build_fused_convs(convs):
    """convs: list of dicts(w_off, cin_total, fwd=packed tensor, bwd=[packed tensor per 48-channel slice]) sorted by
    w_off -> (ctypes FusedConv array, count) for adamw_pack_step.  The tensors must stay alive."""
    arr = (_lib.FusedConv * len(convs))()
    for k, c in enumerate(convs):
        arr[k].w_off = int(c['w_off'])
        arr[k].cout = 48
        arr[k].cin_total = int(c['cin_total'])
        arr[k].fwd = _ptr(c['fwd'], None, 'fwd')
        for s, t in enumerate(c['bwd']):
            arr[k].bwd[s] = _ptr(t, None, 'bwd')
    return arr, len(convs)


def adamw_pack_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, fused, grad_scale=1.0):
    """AdamW over the whole arena + re-pack of the listed convs' forward / backward-data operands, one launch."""
    arr, cnt = fused
    check(_lib.load().lv_adamw_pack_step(_ptr(param, torch.float32, 'param'), _ptr(grad, torch.float32, 'grad'),
                                         _ptr(exp_avg, torch.float32, 'exp_avg'), _ptr(exp_avg_sq, torch.float32, 'exp_avg_sq'),
                                         param.numel(), float(lr), float(beta1), float(beta2), float(eps),
                                         float(weight_decay), int(step), float(grad_scale), arr, cnt, _stream()),
          'lv_adamw_pack_step')


def dp_adamw_pack_step(param, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, fused, peers, grad_scale=1.0):
    """Data-parallel optimizer step in one launch: gradient all-reduce over peer memory + AdamW + operand re-pack.
    `peers` = larvanet_b200.dist.PeerArena (peer-mapped gradient arenas, flag blocks, loss slots, local control words)."""
    arr, cnt = fused
    check(_lib.load().lv_dp_adamw_pack_step(
        _ptr(param, torch.float32, 'param'), _ptr(exp_avg, torch.float32, 'exp_avg'), _ptr(exp_avg_sq, torch.float32, 'exp_avg_sq'),
        param.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale),
        arr, cnt, peers.grad_ptrs, peers.reduced_ptrs if peers.two_shot else None, peers.flag_ptrs, peers.loss_ptrs,
        peers.loss_out.data_ptr(), peers.ctl.data_ptr(), int(peers.slice) if peers.two_shot else 0, int(peers.world),
        int(peers.rank), _stream()), 'lv_dp_adamw_pack_step')


def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    check(_lib.load().lv_adamw_step(_ptr(param, torch.float32, 'param'), _ptr(grad, torch.float32, 'grad'),
                                    _ptr(exp_avg, torch.float32, 'exp_avg'), _ptr(exp_avg_sq, torch.float32, 'exp_avg_sq'),
                                    param.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                                    int(step), float(grad_scale), _stream()), 'lv_adamw_step')
