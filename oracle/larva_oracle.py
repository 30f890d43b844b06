"""CPU oracle for the LarvaNet / LarvaNetV2 / EDSR-baseline x4 SR forward+backward path.

TEST INFRASTRUCTURE ONLY.  This file is a plain-numpy restatement of the arithmetic the
reference delegates to torch.nn (conv3x3+bias, ReLU, residual adds, PixelShuffle, bicubic x4,
L1 multi-exit loss and its full backward).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product path
(`larvanet_b200/`, `models/`) never does and fails loudly when the CUDA library is missing.

Parity pinning: the reference repository holds NO tests, golden vectors or checkpoints for this
path, so the oracle is pinned against outputs of the reference's own modules
executed in the build container (`tests/golden/make_golden.py` imports
`/root/reference/models/{LarvaNet,LarvaNetV2,edsr}.py`, runs them in fp32 on the CPU and commits
inputs' seeds + outputs as `tests/golden/*.npz`).  `tests/test_oracle_golden.py` checks every
function here against those fixtures.

All tensors are NCHW numpy arrays on the 0..255 value scale, like the reference's Python boundary
(reference models/LarvaNet.py:163-171).  Arithmetic is float64 unless `dtype` says otherwise, so
the oracle is *more* exact than the fp32 reference; tests state the tolerance they use.

Parameter dictionaries use the reference's `state_dict()` keys verbatim.
"""
from __future__ import annotations

import numpy as np

NUM_FILTERS = 48  # reference models/LarvaNet.py:226,239,254 (hard-wired)


# --------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------

def _im2col3x3(x):
    """x [N,C,H,W] -> cols [N,H,W,C*9] ordered (c, ky, kx), zero padding=1."""
    n, c, h, w = x.shape
    xp = np.zeros((n, c, h + 2, w + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    cols = np.empty((n, h, w, c, 3, 3), dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            cols[:, :, :, :, ky, kx] = xp[:, :, ky:ky + h, kx:kx + w].transpose(0, 2, 3, 1)
    return cols.reshape(n, h, w, c * 9)


def conv2d(x, weight, bias=None):
    """nn.Conv2d(k, stride=1, padding=k//2) for k in {1,3}.

    Restates the Conv2d call sites: reference models/LarvaNet.py:210-212 (ResidualBlock),
    :227 (LarvaHead), :256-258 (LarvaLeg), models/LarvaNetV2.py:318 (merge_conv),
    models/edsr.py:131 (MeanShift 1x1), :182-193.
    """
    co, ci, kh, kw = weight.shape
    if kh == 1:
        y = np.einsum('nchw,oc->nohw', x, weight[:, :, 0, 0])
    else:
        assert kh == 3 and kw == 3
        cols = _im2col3x3(x)                               # [N,H,W,ci*9]
        y = cols @ weight.reshape(co, ci * 9).T            # [N,H,W,co]
        y = y.transpose(0, 3, 1, 2)
    if bias is not None:
        y = y + bias.reshape(1, co, 1, 1)
    return np.ascontiguousarray(y)


def conv2d_backward(x, weight, dy):
    """Gradients of conv2d wrt input, weight and bias (autograd of the call sites above)."""
    co, ci, kh, kw = weight.shape
    n, _, h, w = x.shape
    db = dy.sum(axis=(0, 2, 3))
    if kh == 1:
        dw = np.einsum('nohw,nchw->oc', dy, x)[:, :, None, None]
        dx = np.einsum('nohw,oc->nchw', dy, weight[:, :, 0, 0])
        return dx, dw, db
    cols = _im2col3x3(x).reshape(-1, ci * 9)
    dyf = dy.transpose(0, 2, 3, 1).reshape(-1, co)
    dw = (dyf.T @ cols).reshape(co, ci, 3, 3)
    # dgrad = conv of dy with the 180-degree rotated, in/out swapped filter
    wt = weight[:, :, ::-1, ::-1].transpose(1, 0, 2, 3)
    dx = conv2d(dy, np.ascontiguousarray(wt))
    return dx, dw, db


def relu(x):
    """nn.ReLU (reference models/LarvaNet.py:211,257)."""
    return np.maximum(x, 0)


def pixel_shuffle(x, r):
    """nn.PixelShuffle(r): out[n,c,r*h+i,r*w+j] = in[n,c*r*r+i*r+j,h,w]
    (reference models/LarvaNet.py:261,265; models/edsr.py:164)."""
    n, c, h, w = x.shape
    co = c // (r * r)
    y = x.reshape(n, co, r, r, h, w).transpose(0, 1, 4, 2, 5, 3)
    return np.ascontiguousarray(y.reshape(n, co, h * r, w * r))


def pixel_unshuffle(y, r):
    """Inverse of pixel_shuffle (its backward)."""
    n, co, hr, wr = y.shape
    h, w = hr // r, wr // r
    x = y.reshape(n, co, h, r, w, r).transpose(0, 1, 3, 5, 2, 4)
    return np.ascontiguousarray(x.reshape(n, co * r * r, h, w))


def _cubic_coeffs(t, a=-0.75):
    """Keys cubic convolution weights for taps floor-1..floor+2 (ATen UpSample.h
    get_cubic_upsample_coefficients; A=-0.75)."""
    def c1(x):  # |x| <= 1
        return ((a + 2) * x - (a + 3)) * x * x + 1
    def c2(x):  # 1 < |x| < 2
        return ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    return np.stack([c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)], axis=-1)


def bicubic_upsample(x, scale=4):
    """F.interpolate(x, scale_factor=4, mode='bicubic', align_corners=False)
    (reference models/LarvaNet.py:283-285).  Separable, src=(d+0.5)/scale-0.5,
    taps floor-1..floor+2 with indices clamped to [0,n-1]."""
    n, c, h, w = x.shape

    def axis_tables(size):
        d = np.arange(size * scale, dtype=np.float64)
        src = (d + 0.5) / scale - 0.5
        fl = np.floor(src)
        t = src - fl
        idx = np.clip(fl[:, None].astype(np.int64) + np.arange(-1, 3)[None, :], 0, size - 1)
        return idx, _cubic_coeffs(t)

    iy, wy = axis_tables(h)
    ix, wx = axis_tables(w)
    xd = x.astype(np.float64)
    # horizontal pass then vertical pass
    tmp = (xd[:, :, :, ix] * wx[None, None, None]).sum(-1)          # [N,C,H,W*scale]
    out = (tmp[:, :, iy, :] * wy[None, None, :, :, None]).sum(3)    # [N,C,H*scale,W*scale]
    return out.astype(x.dtype)


def l1_loss(out, truth):
    """nn.L1Loss() mean reduction (reference models/LarvaNet.py:85,108)."""
    return np.abs(out - truth).mean()


def l1_loss_grad(out, truth, scale=1.0):
    """d mean|out-truth| / d out = sign(out-truth)/numel (times upstream scale)."""
    return np.sign(out - truth) * (scale / out.size)


# --------------------------------------------------------------------------------------
# module-level restatements (forward with tapes, backward)
# --------------------------------------------------------------------------------------

def _p(params, key, dtype):
    return np.asarray(params[key], dtype=dtype)


def _resblock_fwd(params, prefix, x, dtype, res_weight=1.0):
    """ResidualBlock.forward: x + conv(relu(conv(x)))  (reference models/LarvaNet.py:217-220;
    models/edsr.py:150-153 multiplies the branch by `weight`)."""
    t = relu(conv2d(x, _p(params, prefix + '.body.0.weight', dtype), _p(params, prefix + '.body.0.bias', dtype)))
    r = conv2d(t, _p(params, prefix + '.body.2.weight', dtype), _p(params, prefix + '.body.2.bias', dtype))
    return x + r * res_weight, (x, t)


def _resblock_bwd(params, prefix, tape, dout, grads, dtype, res_weight=1.0):
    x, t = tape
    dr = dout * res_weight
    dt, dw2, db2 = conv2d_backward(t, _p(params, prefix + '.body.2.weight', dtype), dr)
    grads[prefix + '.body.2.weight'] = dw2
    grads[prefix + '.body.2.bias'] = db2
    dt = dt * (t > 0)
    dx, dw1, db1 = conv2d_backward(x, _p(params, prefix + '.body.0.weight', dtype), dt)
    grads[prefix + '.body.0.weight'] = dw1
    grads[prefix + '.body.0.bias'] = db1
    return dout + dx


def _body_fwd(params, i, nblocks, x, dtype):
    """LarvaBody.forward: x + res_blocks(x)  (reference models/LarvaNet.py:246-248)."""
    a = x
    tapes = []
    for j in range(nblocks):
        a, tp = _resblock_fwd(params, f'body_{i}.res_blocks.{j}', a, dtype)
        tapes.append(tp)
    return x + a, tapes


def _body_bwd(params, i, tapes, dout, grads, dtype):
    da = dout
    for j in reversed(range(len(tapes))):
        da = _resblock_bwd(params, f'body_{i}.res_blocks.{j}', tapes[j], da, grads, dtype)
    return dout + da


def _recon_fwd(params, prefix, fea, base, dtype):
    """LarvaLeg.forward / tail of LarvaTail.forward: PixelShuffle(4)(conv(relu(conv(fea)))) + base
    (reference models/LarvaNet.py:263-267, models/LarvaNetV2.py:331-334)."""
    u = relu(conv2d(fea, _p(params, prefix + '.recon_block.0.weight', dtype), _p(params, prefix + '.recon_block.0.bias', dtype)))
    v = conv2d(u, _p(params, prefix + '.recon_block.2.weight', dtype), _p(params, prefix + '.recon_block.2.bias', dtype))
    return pixel_shuffle(v, 4) + base, (fea, u)


def _recon_bwd(params, prefix, tape, dout, grads, dtype):
    fea, u = tape
    dv = pixel_unshuffle(dout, 4)
    du, dw2, db2 = conv2d_backward(u, _p(params, prefix + '.recon_block.2.weight', dtype), dv)
    grads[prefix + '.recon_block.2.weight'] = dw2
    grads[prefix + '.recon_block.2.bias'] = db2
    du = du * (u > 0)
    dfea, dw1, db1 = conv2d_backward(fea, _p(params, prefix + '.recon_block.0.weight', dtype), du)
    grads[prefix + '.recon_block.0.weight'] = dw1
    grads[prefix + '.recon_block.0.bias'] = db1
    return dfea


def head_forward(params, x, dtype=np.float64):
    """LarvaHead.forward (reference models/LarvaNet.py:231-233)."""
    return conv2d(np.asarray(x, dtype), _p(params, 'head.feature_extraction.weight', dtype),
                  _p(params, 'head.feature_extraction.bias', dtype))


def larvanet_forward(params, x, blocks, exit_leg=None, dtype=np.float64):
    """LarvaNetModule.forward (reference models/LarvaNet.py:287-293): head -> bodies -> base ->
    last leg.  `exit_leg=k` restates the early exit of models/LarvaLeg.py:290-299
    (k==0 -> bicubic only; otherwise run k bodies and use leg k-1)."""
    x = np.asarray(x, dtype)
    m = len(blocks)
    k = m if exit_leg is None else exit_leg
    base = bicubic_upsample(x, 4)
    if k == 0:
        return base
    fea = head_forward(params, x, dtype)
    for i in range(k):
        fea, _ = _body_fwd(params, i, blocks[i], fea, dtype)
    out, _ = _recon_fwd(params, f'body_{k - 1}.leg', fea, base, dtype)
    return out


def larvanet_v2_forward(params, x, blocks, dtype=np.float64):
    """LarvaNetV2 LarvaNetModule.forward (reference models/LarvaNetV2.py:355-365): legs unused,
    tail = cat(features) -> merge_conv -> recon -> PS4 -> + base (:328-334)."""
    x = np.asarray(x, dtype)
    fea = head_forward(params, x, dtype)
    feats = []
    for i in range(len(blocks)):
        fea, _ = _body_fwd(params, i, blocks[i], fea, dtype)
        feats.append(fea)
    base = bicubic_upsample(x, 4)
    cat = np.concatenate(feats, axis=1)
    mfea = conv2d(cat, _p(params, 'tail.merge_conv.weight', dtype), _p(params, 'tail.merge_conv.bias', dtype))
    out, _ = _recon_fwd(params, 'tail', mfea, base, dtype)
    return out


def larvanet_train_step(params, x, truth, blocks, v2=False, dtype=np.float64, sign_from=None, tapes_from=None):
    """Forward + backward core of train_step_larva.

    V1: reference models/LarvaNet.py:102-113 -- loss = sum_i L1(leg_i(body_i(..)), truth) / M.
    V2: reference models/LarvaNetV2.py:105-118 -- M leg losses + tail loss, divided by M+1.
    Returns (loss, grads dict keyed like state_dict, list of per-exit outputs).

    Two optional hooks exist for parity tests of a reduced-precision device path, because ReLU and the L1 sign are
    discontinuous (a forward error of 1e-2 flips a few masks/signs, which moves gradients by several percent in L2
    although every kernel is exact):
      `sign_from`  : list of per-exit output arrays used inside sign(out - truth) instead of the oracle's own exits;
      `tapes_from` : dict of the device's saved forward activations (NCHW): 'f0', ('a',i,j) block inputs for j>0,
                     ('t',i,j) post-ReLU, ('feat',i), ('u',i) and for V2 'mf','ut'.  The backward pass then runs on
                     exactly the state the device's backward kernels saw.
    With both None this is the plain oracle that is pinned to the reference's golden outputs.
    """
    x = np.asarray(x, dtype)
    truth = np.asarray(truth, dtype)
    m = len(blocks)
    denom = m + 1 if v2 else m
    fea0 = head_forward(params, x, dtype)
    base = bicubic_upsample(x, 4)
    feats, body_tapes, leg_tapes, outs = [], [], [], []
    fea = fea0
    loss = 0.0
    for i in range(m):
        fea, tp = _body_fwd(params, i, blocks[i], fea, dtype)
        feats.append(fea)
        body_tapes.append(tp)
        out, ltp = _recon_fwd(params, f'body_{i}.leg', fea, base, dtype)
        leg_tapes.append(ltp)
        outs.append(out)
        loss += l1_loss(out, truth)
    cat = mw = ttp = None
    if v2:
        cat = np.concatenate(feats, axis=1)
        mw = _p(params, 'tail.merge_conv.weight', dtype)
        mfea = conv2d(cat, mw, _p(params, 'tail.merge_conv.bias', dtype))
        tout, ttp = _recon_fwd(params, 'tail', mfea, base, dtype)
        outs.append(tout)
        loss += l1_loss(tout, truth)
    loss = loss / denom

    if tapes_from is not None:
        T = {k: np.asarray(v, dtype) for k, v in tapes_from.items()}
        feats = [T[('feat', i)] for i in range(m)]
        body_tapes = []
        for i in range(m):
            fin = T['f0'] if i == 0 else feats[i - 1]
            body_tapes.append([(fin if j == 0 else T[('a', i, j)], T[('t', i, j)]) for j in range(blocks[i])])
        leg_tapes = [(feats[i], T[('u', i)]) for i in range(m)]
        if v2:
            cat = np.concatenate(feats, axis=1)
            ttp = (T['mf'], T['ut'])
    souts = outs if sign_from is None else [np.asarray(o, dtype) for o in sign_from]

    grads = {}
    dfeats = [np.zeros_like(f) for f in feats]
    if v2:
        dmfea = _recon_bwd(params, 'tail', ttp, l1_loss_grad(souts[m], truth, 1.0 / denom), grads, dtype)
        dcat, dmw, dmb = conv2d_backward(cat, mw, dmfea)
        grads['tail.merge_conv.weight'] = dmw
        grads['tail.merge_conv.bias'] = dmb
        for i in range(m):
            dfeats[i] = dfeats[i] + dcat[:, NUM_FILTERS * i:NUM_FILTERS * (i + 1)]
    dnext = None
    for i in reversed(range(m)):
        d = dfeats[i] + _recon_bwd(params, f'body_{i}.leg', leg_tapes[i],
                                   l1_loss_grad(souts[i], truth, 1.0 / denom), grads, dtype)
        if dnext is not None:
            d = d + dnext
        dnext = _body_bwd(params, i, body_tapes[i], d, grads, dtype)
    _, dwh, dbh = conv2d_backward(x, _p(params, 'head.feature_extraction.weight', dtype), dnext)
    grads['head.feature_extraction.weight'] = dwh
    grads['head.feature_extraction.bias'] = dbh
    return float(loss), grads, outs


def edsr_forward(params, x, num_res_blocks=16, res_weight=1.0, scale=4, dtype=np.float64):
    """EDSRModule.forward (reference models/edsr.py:195-207).  `mean_shift` and
    `mean_inverse_shift` are general 1x1 convs whose (frozen, default-initialised) weights live in
    the state_dict -- MeanShift.__init__ never overwrites them (models/edsr.py:129-136)."""
    x = np.asarray(x, dtype)
    x = conv2d(x, _p(params, 'mean_shift.weight', dtype), _p(params, 'mean_shift.bias', dtype))
    x = conv2d(x, _p(params, 'first_conv.weight', dtype), _p(params, 'first_conv.bias', dtype))
    res = x
    for j in range(num_res_blocks):
        res, _ = _resblock_fwd(params, f'res_blocks.{j}', res, dtype, res_weight)
    res = conv2d(res, _p(params, 'after_res_conv.weight', dtype), _p(params, 'after_res_conv.bias', dtype))
    x = x + res
    nup = {2: 1, 4: 2, 8: 3}[scale]
    for s in range(nup):
        x = conv2d(x, _p(params, f'upsample.body.{2 * s}.weight', dtype), _p(params, f'upsample.body.{2 * s}.bias', dtype))
        x = pixel_shuffle(x, 2)
    x = conv2d(x, _p(params, 'final_conv.weight', dtype), _p(params, 'final_conv.bias', dtype))
    x = conv2d(x, _p(params, 'mean_inverse_shift.weight', dtype), _p(params, 'mean_inverse_shift.bias', dtype))
    return x


# --------------------------------------------------------------------------------------
# metric helpers (reference validate.py:17-27)
# --------------------------------------------------------------------------------------

def image_to_uint8(image):
    """validate._image_to_uint8 (reference validate.py:17-18)."""
    return np.clip(np.round(image), 0, 255).astype(np.uint8)


def image_psnr(output_image, truth_image):
    """validate._image_psnr (reference validate.py:23-27)."""
    diff = np.float32(truth_image) - np.float32(output_image)
    mse = np.mean(np.power(diff, 2))
    return 10.0 * np.log10(255.0 ** 2 / mse)


def adamw_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """One torch.optim.AdamW update (reference models/LarvaNet.py:86-88,114; torch defaults
    betas=(0.9,0.999), eps=1e-8, weight_decay=0.01, amsgrad=False)."""
    param = param * (1.0 - lr * weight_decay)
    exp_avg = beta1 * exp_avg + (1 - beta1) * grad
    exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) * grad * grad
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(exp_avg_sq) / np.sqrt(bc2) + eps
    param = param - (lr / bc1) * exp_avg / denom
    return param, exp_avg, exp_avg_sq
