"""Kernel-level parity tests through the C-ABI (ctypes), `-m gpu` only.

Every test feeds the CUDA kernel and the numpy oracle the SAME operands (for bf16: the oracle gets the bf16-rounded
values as float64), so the only differences are fp32-vs-fp64 accumulation and the final store rounding:
  bf16 store: |err| <= 2^-8 * |ref| + small abs;   fp32 mode: <= 1e-5 relative.
"""
import os

import numpy as np
import pytest
import torch

from larvanet_b200 import _lib, ops
from oracle import larva_oracle as O
from tests.gpu_util import act_empty, bf16_round, from_nhwc, rel_l2, to_nhwc

pytestmark = pytest.mark.gpu

DT = {'bf16': torch.bfloat16, 'fp32': torch.float32}


def _rand_conv(rs, cout, cin_total, dtype, wscale=0.05):
    w = (rs.standard_normal((cout, cin_total, 3, 3)) * wscale).astype(np.float32)
    b = (rs.standard_normal(cout) * 0.5).astype(np.float32)
    wq = bf16_round(w) if dtype == torch.bfloat16 else w.astype(np.float64)
    return w, b, wq


def _pack(w, dtype, cin, transpose=0, i_off=0, i_cnt=None, wlayout=0):
    wt = torch.from_numpy(w).cuda()
    O_, I_ = w.shape[:2]
    i_cnt = I_ if i_cnt is None else i_cnt
    p_cout, p_cin_total = (i_cnt, O_) if transpose else (O_, i_cnt)
    packed = torch.zeros(ops.packed_weight_bytes(p_cout, p_cin_total, dtype), dtype=torch.uint8, device='cuda')
    ops.pack_weights([dict(w=wt, packed=packed, transpose=transpose, i_off=i_off, i_cnt=i_cnt, cin=cin, dtype=dtype,
                           wlayout=wlayout)])
    return packed


def _act(rs, n, c, h, w, dtype, scale=1.0):
    x = (rs.standard_normal((n, c, h, w)) * scale).astype(np.float32)
    xq = bf16_round(x) if dtype == torch.bfloat16 else x.astype(np.float64)
    return to_nhwc(x, dtype), xq


def _tol(dtype):
    return (2.0 ** -8, 2e-3) if dtype == torch.bfloat16 else (1e-5, 1e-5)


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
@pytest.mark.parametrize('shape', [(1, 16, 8), (2, 19, 13), (1, 5, 3), (3, 48, 48), (1, 33, 70)])
@pytest.mark.parametrize('simt', [False, True])
def test_conv48_plain(prec, shape, simt):
    dtype = DT[prec]
    if prec == 'fp32' and simt:
        pytest.skip('fp32 is always the CUDA-core kernel')
    n, h, w = shape
    rs = np.random.RandomState(1000 * shape[0] + 31 * shape[1] + shape[2])
    wt, b, wq = _rand_conv(rs, 48, 48, dtype)
    x, xq = _act(rs, n, 48, h, w, dtype)
    packed = _pack(wt, dtype, 48)
    out = torch.empty_like(x)
    ops.conv3x3([x], packed, 48, bias=torch.from_numpy(b).cuda(), out=out, simt=simt)
    torch.cuda.synchronize()
    ref = O.conv2d(xq, wq, b.astype(np.float64))
    rtol, atol = _tol(dtype)
    np.testing.assert_allclose(from_nhwc(out), ref, rtol=rtol, atol=atol)


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
def test_conv48_epilogue_relu_mask_residuals(prec):
    dtype = DT[prec]
    n, h, w = 2, 21, 11
    rs = np.random.RandomState(7)
    wt, b, wq = _rand_conv(rs, 48, 48, dtype)
    x, xq = _act(rs, n, 48, h, w, dtype)
    r1, r1q = _act(rs, n, 48, h, w, dtype)
    r2, r2q = _act(rs, n, 48, h, w, dtype)
    mk, mkq = _act(rs, n, 48, h, w, dtype)
    packed = _pack(wt, dtype, 48)
    bt = torch.from_numpy(b).cuda()
    rtol, atol = _tol(dtype)
    conv = O.conv2d(xq, wq, b.astype(np.float64))
    # relu + two residuals (ResidualBlock conv2 with the LarvaBody skip folded in)
    out = torch.empty_like(x)
    ops.conv3x3([x], packed, 48, bias=bt, out=out, relu=True, res1=r1, res2=r2)
    np.testing.assert_allclose(from_nhwc(out), np.maximum(conv, 0) + r1q + r2q, rtol=rtol, atol=atol)
    # mask + residual, no bias (backward-data epilogue)
    out2 = torch.empty_like(x)
    ops.conv3x3([x], packed, 48, bias=None, out=out2, mask=mk, res1=r1)
    ref2 = O.conv2d(xq, wq) * (mkq > 0) + r1q
    np.testing.assert_allclose(from_nhwc(out2), ref2, rtol=rtol, atol=atol)
    # in-place accumulate: out aliases res1
    acc = r1.clone()
    ops.conv3x3([x], packed, 48, bias=None, out=acc, res1=acc)
    np.testing.assert_allclose(from_nhwc(acc), O.conv2d(xq, wq) + r1q, rtol=rtol, atol=atol)


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
def test_conv_pixelshuffle_base_loss(prec):
    dtype = DT[prec]
    n, h, w = 2, 18, 9
    rs = np.random.RandomState(11)
    wt, b, wq = _rand_conv(rs, 48, 48, dtype, wscale=0.3)
    x, xq = _act(rs, n, 48, h, w, dtype)
    base = (rs.uniform(0, 255, (n, 3, 4 * h, 4 * w))).astype(np.float32)
    truth = (rs.uniform(0, 255, (n, 3, 4 * h, 4 * w))).astype(np.float32)
    packed = _pack(wt, dtype, 48)
    out_hr = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device='cuda')
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    g = torch.empty_like(x)
    ops.conv3x3([x], packed, 48, bias=torch.from_numpy(b).cuda(), epilogue=_lib.LV_EPI_PS4_NCHW, out_hr=out_hr,
                base_hr=torch.from_numpy(base).cuda(), truth_hr=torch.from_numpy(truth).cuda(), loss_sum=loss, grad_sign=g)
    ref = O.pixel_shuffle(O.conv2d(xq, wq, b.astype(np.float64)), 4) + base
    got = out_hr.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=2e-3 if prec == 'bf16' else 2e-4)
    # loss / sign gradient must be consistent with the kernel's own fp32 output
    assert abs(loss.item() - np.abs(got - truth).sum()) <= 1e-6 * np.abs(got - truth).sum()
    gref = O.pixel_unshuffle(np.sign(got - truth.astype(np.float64)), 4)
    np.testing.assert_array_equal(from_nhwc(g), gref)


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
@pytest.mark.parametrize('nsrc', [2, 4])
def test_conv_multi_source_merge(prec, nsrc):
    dtype = DT[prec]
    n, h, w = 1, 17, 20
    rs = np.random.RandomState(13 + nsrc)
    wt, b, wq = _rand_conv(rs, 48, 48 * nsrc, dtype)
    srcs, srcq = zip(*[_act(rs, n, 48, h, w, dtype) for _ in range(nsrc)])
    packed = _pack(wt, dtype, 48)
    out = torch.empty_like(srcs[0])
    ops.conv3x3(list(srcs), packed, 48, bias=torch.from_numpy(b).cuda(), out=out)
    ref = O.conv2d(np.concatenate(srcq, axis=1), wq, b.astype(np.float64))
    rtol, atol = _tol(dtype)
    np.testing.assert_allclose(from_nhwc(out), ref, rtol=rtol, atol=atol * 2)


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
def test_dgrad_operand(prec):
    """conv with the transposed/rotated packed operand == oracle conv2d_backward wrt the input."""
    dtype = DT[prec]
    n, h, w = 2, 16, 12
    rs = np.random.RandomState(17)
    wt, _, wq = _rand_conv(rs, 48, 96, dtype)
    dy, dyq = _act(rs, n, 48, h, w, dtype)
    x_dummy = np.zeros((n, 96, h, w))
    dx_ref, _, _ = O.conv2d_backward(x_dummy, wq, dyq)
    rtol, atol = _tol(dtype)
    for s in range(2):
        packed = _pack(wt, dtype, 48, transpose=1, i_off=48 * s, i_cnt=48)
        out = torch.empty_like(dy)
        ops.conv3x3([dy], packed, 48, bias=None, out=out)
        np.testing.assert_allclose(from_nhwc(out), dx_ref[:, 48 * s:48 * (s + 1)], rtol=rtol, atol=atol)


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
def test_head_and_bicubic(prec):
    dtype = DT[prec]
    n, h, w = 2, 37, 21
    rs = np.random.RandomState(19)
    x = rs.uniform(0, 255, (n, 3, h, w)).astype(np.float32)
    wt = (rs.standard_normal((48, 3, 3, 3)) * 0.05).astype(np.float32)
    b = (rs.standard_normal(48) * 0.5).astype(np.float32)
    fea = act_empty(n, h, w, 48, dtype)
    base = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device='cuda')
    ops.head_bicubic(torch.from_numpy(x).cuda(), torch.from_numpy(wt).cuda(), torch.from_numpy(b).cuda(), fea, base)
    ref = O.conv2d(x.astype(np.float64), wt.astype(np.float64), b.astype(np.float64))
    rtol, atol = _tol(dtype)
    np.testing.assert_allclose(from_nhwc(fea), ref, rtol=rtol, atol=max(atol, 1e-3))
    bref = O.bicubic_upsample(x.astype(np.float64), 4)
    np.testing.assert_allclose(base.cpu().numpy(), bref, rtol=0, atol=2e-4)
    out2 = torch.empty_like(base)
    ops.bicubic_x4(torch.from_numpy(x).cuda(), out2)
    assert torch.equal(out2, base)


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
@pytest.mark.parametrize('simt', [False, True])
def test_wgrad(prec, simt):
    dtype = DT[prec]
    if prec == 'fp32' and simt:
        pytest.skip('fp32 is always the CUDA-core kernel')
    n, h, w = 2, 23, 19
    rs = np.random.RandomState(23)
    x, xq = _act(rs, n, 48, h, w, dtype)
    dy, dyq = _act(rs, n, 48, h, w, dtype)
    dw = torch.zeros((48, 48, 3, 3), dtype=torch.float32, device='cuda')
    db = torch.zeros(48, dtype=torch.float32, device='cuda')
    batch = ops.WgradBatch([dict(x=x, dy=dy, dw=dw, db=db, scale=0.5)], splits=5, device='cuda')
    batch.launch(simt=simt)
    batch.launch(simt=simt)  # accumulates: second launch doubles the result
    _, dw_ref, db_ref = O.conv2d_backward(xq, np.zeros((48, 48, 3, 3)), dyq)
    assert rel_l2(dw.cpu().numpy(), dw_ref) < 1e-5
    assert rel_l2(db.cpu().numpy(), db_ref) < 1e-5


@pytest.mark.parametrize('prec', ['bf16', 'fp32'])
def test_wgrad_merge_slices_and_head(prec):
    dtype = DT[prec]
    n, h, w = 1, 16, 24
    rs = np.random.RandomState(29)
    xs, xq = zip(*[_act(rs, n, 48, h, w, dtype) for _ in range(2)])
    dy, dyq = _act(rs, n, 48, h, w, dtype)
    dw = torch.zeros((48, 96, 3, 3), dtype=torch.float32, device='cuda')
    db = torch.zeros(48, dtype=torch.float32, device='cuda')
    items = [dict(x=xs[s], dy=dy, dw=dw, db=db if s == 0 else None, cin_total=96, cin_off=48 * s) for s in range(2)]
    ops.WgradBatch(items, splits=3, device='cuda').launch()
    _, dw_ref, db_ref = O.conv2d_backward(np.concatenate(xq, 1), np.zeros((48, 96, 3, 3)), dyq)
    assert rel_l2(dw.cpu().numpy(), dw_ref) < 1e-5
    assert rel_l2(db.cpu().numpy(), db_ref) < 1e-5
    # head weight gradient (fp32 image input)
    img = rs.uniform(0, 255, (n, 3, h, w)).astype(np.float32)
    dwh = torch.zeros((48, 3, 3, 3), dtype=torch.float32, device='cuda')
    dbh = torch.zeros(48, dtype=torch.float32, device='cuda')
    ops.head_wgrad(torch.from_numpy(img).cuda(), dy, dwh, dbh, 0.25)
    _, dwh_ref, dbh_ref = O.conv2d_backward(img.astype(np.float64), np.zeros((48, 3, 3, 3)), dyq)
    assert rel_l2(dwh.cpu().numpy(), 0.25 * dwh_ref) < 1e-5
    assert rel_l2(dbh.cpu().numpy(), 0.25 * dbh_ref) < 1e-5
    # accumulate semantics (second call adds), overwrite semantics (garbage in, gradient out), bitwise determinism of the
    # two-pass reduction, and a shape with more tiles than resident blocks (ragged edges)
    ops.head_wgrad(torch.from_numpy(img).cuda(), dy, dwh, dbh, 0.25)
    assert rel_l2(dwh.cpu().numpy(), 0.5 * dwh_ref) < 1e-5
    junk_w = torch.full((48, 3, 3, 3), 7.0, device='cuda')
    junk_b = torch.full((48,), -3.0, device='cuda')
    ops.head_wgrad(torch.from_numpy(img).cuda(), dy, junk_w, junk_b, 0.25, overwrite=True)
    assert rel_l2(junk_w.cpu().numpy(), 0.25 * dwh_ref) < 1e-5 and rel_l2(junk_b.cpu().numpy(), 0.25 * dbh_ref) < 1e-5
    if dtype == torch.bfloat16:
        n2, h2, w2 = 9, 77, 150
        dy2, dy2q = _act(rs, n2, 48, h2, w2, dtype)
        img2 = rs.uniform(0, 255, (n2, 3, h2, w2)).astype(np.float32)
        a_w, a_b = torch.empty((48, 3, 3, 3), device='cuda'), torch.empty(48, device='cuda')
        b_w, b_b = torch.empty_like(a_w), torch.empty_like(a_b)
        ops.head_wgrad(torch.from_numpy(img2).cuda(), dy2, a_w, a_b, 1e-3, overwrite=True)
        ops.head_wgrad(torch.from_numpy(img2).cuda(), dy2, b_w, b_b, 1e-3, overwrite=True)
        assert torch.equal(a_w, b_w) and torch.equal(a_b, b_b)
        _, ref_w, ref_b = O.conv2d_backward(img2.astype(np.float64), np.zeros((48, 3, 3, 3)), dy2q)
        assert rel_l2(a_w.cpu().numpy(), 1e-3 * ref_w) < 1e-5 and rel_l2(a_b.cpu().numpy(), 1e-3 * ref_b) < 1e-5


def test_layout_loss_adamw_helpers():
    rs = np.random.RandomState(31)
    x = rs.standard_normal((2, 48, 7, 5)).astype(np.float32)
    for dtype in (torch.float32, torch.bfloat16):
        a = act_empty(2, 7, 5, 48, dtype)
        ops.nchw_to_nhwc(torch.from_numpy(x).cuda(), a)
        back = torch.empty((2, 48, 7, 5), dtype=torch.float32, device='cuda')
        ops.nhwc_to_nchw(a, back)
        ref = bf16_round(x) if dtype == torch.bfloat16 else x
        np.testing.assert_array_equal(back.cpu().numpy(), ref.astype(np.float32))
    out = rs.uniform(0, 255, (2, 3, 16, 12)).astype(np.float32)
    truth = rs.uniform(0, 255, (2, 3, 16, 12)).astype(np.float32)
    truth[0, 0, 0, 0] = out[0, 0, 0, 0]  # sign(0) == 0
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    g = act_empty(2, 4, 3, 48, torch.bfloat16)
    ops.l1_loss_grad(torch.from_numpy(out).cuda(), torch.from_numpy(truth).cuda(), loss, g)
    assert abs(loss.item() - np.abs(out.astype(np.float64) - truth).sum()) < 1e-3
    np.testing.assert_array_equal(from_nhwc(g), O.pixel_unshuffle(np.sign(out.astype(np.float64) - truth), 4))
    # AdamW, 3 steps vs the oracle restatement of torch.optim.AdamW
    p0 = rs.standard_normal(1000).astype(np.float32)
    p = torch.from_numpy(p0.copy()).cuda()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    pr, mr, vr = p0.astype(np.float64), np.zeros(1000), np.zeros(1000)
    for step in range(1, 4):
        gnp = rs.standard_normal(1000).astype(np.float32)
        ops.adamw_step(p, torch.from_numpy(gnp).cuda(), m, v, 4e-4, 0.9, 0.999, 1e-8, 0.01, step)
        pr, mr, vr = O.adamw_step(pr, gnp.astype(np.float64), mr, vr, step, 4e-4)
    np.testing.assert_allclose(p.cpu().numpy(), pr, rtol=1e-5, atol=1e-6)


def test_errors_are_loud():
    x = torch.zeros((1, 8, 5, 8, 8), dtype=torch.bfloat16, device='cuda')   # 40 channels: not a multiple of 16
    with pytest.raises(_lib.LarvaNetB200Error):
        ops.conv3x3([x], torch.zeros(16, dtype=torch.uint8, device='cuda'), 48, out=act_empty(1, 8, 8, 48, torch.bfloat16))
    with pytest.raises(_lib.LarvaNetB200Error):   # NHWC-shaped tensors are rejected: the library's layout is planar-8
        ops.conv3x3([torch.zeros((1, 8, 8, 48), dtype=torch.bfloat16, device='cuda')], torch.zeros(16, dtype=torch.uint8, device='cuda'), 48,
                    out=act_empty(1, 8, 8, 48, torch.bfloat16))
    with pytest.raises(_lib.LarvaNetB200Error):
        ops.conv3x3([x.cpu()], torch.zeros(16, dtype=torch.uint8), 48, out=x.cpu())


@pytest.mark.parametrize('shape', [(1, 14, 8), (2, 19, 13), (1, 5, 3), (3, 48, 48), (1, 33, 70), (1, 180, 320)])
@pytest.mark.parametrize('simt', [False, True])
def test_conv48_ky_stacked_layout(shape, simt):
    """The row-marching tensor-core kernel (ky-stacked weights: 9 MMAs of N=144 per 128 px of a row, vertical taps summed
    in TMEM; csrc/conv_row.cu) against the oracle, all epilogue paths."""
    dtype = torch.bfloat16
    n, h, w = shape
    rs = np.random.RandomState(7000 + 31 * h + w)
    wt, b, wq = _rand_conv(rs, 48, 48, dtype)
    x, xq = _act(rs, n, 48, h, w, dtype)
    r1, r1q = _act(rs, n, 48, h, w, dtype)
    mk, mkq = _act(rs, n, 48, h, w, dtype)
    packed = _pack(wt, dtype, 48, wlayout=_lib.LV_W_KY_STACKED)
    bt = torch.from_numpy(b).cuda()
    rtol, atol = _tol(dtype)
    conv = O.conv2d(xq, wq, b.astype(np.float64)) if h * w <= 4096 else None
    out = torch.empty_like(x)
    ops.conv3x3([x], packed, 48, bias=bt, out=out, relu=True, res1=r1, simt=simt, wlayout=_lib.LV_W_KY_STACKED)
    if conv is not None:
        np.testing.assert_allclose(from_nhwc(out), np.maximum(conv, 0) + r1q, rtol=rtol, atol=atol)
    else:   # full 720p frame: cross-check against the tap-major tensor-core kernel on the same operands
        ref = torch.empty_like(x)
        ops.conv3x3([x], _pack(wt, dtype, 48), 48, bias=bt, out=ref, relu=True, res1=r1)
        assert (out.float() - ref.float()).abs().max().item() <= 0.0625
    if conv is not None:
        out2 = torch.empty_like(x)
        ops.conv3x3([x], packed, 48, bias=None, out=out2, mask=mk, res1=r1, simt=simt, wlayout=_lib.LV_W_KY_STACKED)
        np.testing.assert_allclose(from_nhwc(out2), O.conv2d(xq, wq) * (mkq > 0) + r1q, rtol=rtol, atol=atol)
        # PixelShuffle(4)+base+loss epilogue and the backward-data operand in the same layout
        base = rs.uniform(0, 255, (n, 3, 4 * h, 4 * w)).astype(np.float32)
        out_hr = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device='cuda')
        ops.conv3x3([x], packed, 48, bias=bt, epilogue=_lib.LV_EPI_PS4_NCHW, out_hr=out_hr,
                    base_hr=torch.from_numpy(base).cuda(), simt=simt, wlayout=_lib.LV_W_KY_STACKED)
        np.testing.assert_allclose(out_hr.cpu().numpy(), O.pixel_shuffle(conv, 4) + base, rtol=2e-6, atol=2e-3)
        pk_t = _pack(wt, dtype, 48, transpose=1, wlayout=_lib.LV_W_KY_STACKED)
        dx = torch.empty_like(x)
        ops.conv3x3([x], pk_t, 48, bias=None, out=dx, simt=simt, wlayout=_lib.LV_W_KY_STACKED)
        dx_ref, _, _ = O.conv2d_backward(np.zeros_like(xq), wq, xq)
        np.testing.assert_allclose(from_nhwc(dx), dx_ref, rtol=rtol, atol=atol)


def test_conv_ky_stacked_multi_source_is_rejected():
    """The row-marching kernel keeps ONE layer's weights resident: multi-source convs (V2 merge) stay on the tap-major
    kernel and asking for the ky-stacked layout fails loudly instead of computing something else."""
    dtype = torch.bfloat16
    rs = np.random.RandomState(77)
    wt, b, wq = _rand_conv(rs, 48, 192, dtype)
    srcs, _ = zip(*[_act(rs, 1, 48, 30, 20, dtype) for _ in range(4)])
    packed = _pack(wt, dtype, 48, wlayout=_lib.LV_W_KY_STACKED)
    if os.environ.get('LARVANET_B200_EXPERIMENTAL') == '1':
        pytest.skip('experimental build routes this shape to tools/experiments/conv_tc_ky.cu')
    with pytest.raises(_lib.LarvaNetB200Error):
        ops.conv3x3(list(srcs), packed, 48, bias=torch.from_numpy(b).cuda(), out=torch.empty_like(srcs[0]),
                    wlayout=_lib.LV_W_KY_STACKED)


@pytest.mark.parametrize('shape', [(1, 12, 16), (2, 19, 13), (1, 3, 130), (3, 40, 40), (1, 70, 200)])
def test_conv64_row_kernel(shape):
    """EDSR's 64 -> 64 convs on the row-marching kernel (8-block accumulator ring, N = 192): ReLU and residual epilogues
    against the oracle on bf16-rounded operands."""
    dtype = torch.bfloat16
    n, h, w = shape
    rs = np.random.RandomState(6400 + 31 * h + w)
    wt, b, wq = _rand_conv(rs, 64, 64, dtype, wscale=0.04)
    x, xq = _act(rs, n, 64, h, w, dtype)
    r1, r1q = _act(rs, n, 64, h, w, dtype)
    packed = _pack(wt, dtype, 64, wlayout=_lib.LV_W_KY_STACKED)
    bt = torch.from_numpy(b).cuda()
    rtol, atol = _tol(dtype)
    conv = O.conv2d(xq, wq, b.astype(np.float64))
    out = torch.empty_like(x)
    ops.conv3x3([x], packed, 64, bias=bt, out=out, relu=True, wlayout=_lib.LV_W_KY_STACKED)
    np.testing.assert_allclose(from_nhwc(out), np.maximum(conv, 0), rtol=rtol, atol=atol)
    out2 = torch.empty_like(x)
    ops.conv3x3([x], packed, 64, bias=bt, out=out2, res1=r1, wlayout=_lib.LV_W_KY_STACKED)
    np.testing.assert_allclose(from_nhwc(out2), conv + r1q, rtol=rtol, atol=atol)
    out3 = torch.empty_like(x)          # res_scale != 1 takes the generic epilogue
    ops.conv3x3([x], packed, 64, bias=bt, out=out3, res1=r1, res_scale=0.5, wlayout=_lib.LV_W_KY_STACKED)
    np.testing.assert_allclose(from_nhwc(out3), 0.5 * conv + r1q, rtol=rtol, atol=atol)


@pytest.mark.parametrize('shape', [(1, 9, 130), (2, 21, 200), (1, 40, 129), (1, 5, 60)])
def test_conv64_to_rgb_row_kernel(shape):
    """EDSR's last conv (64 -> 3, RGB epilogue with the mean-shift 3x3 matrix) on the row-marching kernel (cout padded
    to 16: N = 48 per MMA, 32-block accumulator ring) against the oracle and against the tile kernel; widths below 129
    pixels take the cp.async build."""
    dtype = torch.bfloat16
    n, h, w = shape
    rs = np.random.RandomState(6403 + 17 * h + w)
    wt, b, wq = _rand_conv(rs, 3, 64, dtype, wscale=0.04)
    x, xq = _act(rs, n, 64, h, w, dtype)
    post_w = rs.uniform(-1, 1, (3, 3)).astype(np.float32)
    post_b = rs.uniform(-1, 1, 3).astype(np.float32)
    bt = torch.from_numpy(b).cuda()
    pw, pb = torch.from_numpy(post_w).cuda(), torch.from_numpy(post_b).cuda()
    conv = O.conv2d(xq, wq, b.astype(np.float64))                       # [n, 3, h, w]
    want = np.einsum('dc,nchw->ndhw', post_w.astype(np.float64), conv) + post_b.astype(np.float64)[None, :, None, None]
    out = torch.full((n, 3, h, w), -7.0, dtype=torch.float32, device='cuda')
    ops.conv3x3([x], _pack(wt, dtype, 64, wlayout=_lib.LV_W_KY_STACKED), 3, bias=bt, epilogue=_lib.LV_EPI_RGB_NCHW,
                out_hr=out, post_w=pw, post_b=pb, wlayout=_lib.LV_W_KY_STACKED)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=2e-5, atol=2e-3)
    ref = torch.empty_like(out)
    ops.conv3x3([x], _pack(wt, dtype, 64), 3, bias=bt, epilogue=_lib.LV_EPI_RGB_NCHW, out_hr=ref, post_w=pw, post_b=pb)
    assert (out - ref).abs().max().item() <= 2e-3     # same products, different fp32 summation order


@pytest.mark.parametrize('shape', [(1, 270, 480), (8, 64, 64), (16, 48, 48), (1, 1, 1), (1, 2, 127), (1, 129, 128), (2, 11, 128)])
def test_row_kernel_equals_tap_major_kernel_at_size(shape):
    """Row-marching kernel vs the tap-major tensor-core kernel on identical operands at BASELINE frame / batch sizes and
    at the lane-boundary widths (127 / 128 px, 1-row images): same products, different fp32 summation order."""
    dtype = torch.bfloat16
    n, h, w = shape
    rs = np.random.RandomState(9100 + h + w)
    wt, b, _ = _rand_conv(rs, 48, 48, dtype)
    x, _ = _act(rs, n, 48, h, w, dtype)
    r1, _ = _act(rs, n, 48, h, w, dtype)
    bt = torch.from_numpy(b).cuda()
    a, ref = torch.empty_like(x), torch.empty_like(x)
    ops.conv3x3([x], _pack(wt, dtype, 48, wlayout=_lib.LV_W_KY_STACKED), 48, bias=bt, out=a, relu=True, res1=r1,
                wlayout=_lib.LV_W_KY_STACKED)
    ops.conv3x3([x], _pack(wt, dtype, 48), 48, bias=bt, out=ref, relu=True, res1=r1)
    d = (a.float() - ref.float()).abs()
    assert d.max().item() <= 0.0625 and (d > 0).float().mean().item() < 0.05     # rare 1-ulp bf16 flips only


def _chain_layers(rs, n, h, w, depth, wlayout=0):
    """A LarvaNet-like layer list on ping-pong buffers: resblocks (relu conv, conv + skips), a leg with PixelShuffle +
    loss epilogue, and masked input-gradient style layers.  Returns (make_args(bufs) -> [ConvArgs], make_bufs)."""
    dtype = torch.bfloat16
    convs = []
    for _ in range(8):
        wt, b, _ = _rand_conv(rs, 48, 48, dtype, wscale=0.04)
        convs.append((_pack(wt, dtype, 48, wlayout=wlayout), torch.from_numpy(b).cuda()))
    x0, _ = _act(rs, n, 48, h, w, dtype)
    base = torch.from_numpy(rs.uniform(0, 255, (n, 3, 4 * h, 4 * w)).astype(np.float32)).cuda()
    truth = torch.from_numpy(rs.uniform(0, 255, (n, 3, 4 * h, 4 * w)).astype(np.float32)).cuda()

    def make_bufs():
        z = lambda: torch.zeros_like(x0)
        return dict(x=x0.clone(), t=z(), a=z(), b=z(), u=z(), d=z(), g=z(),
                    hr=torch.zeros_like(base), loss=torch.zeros(1, dtype=torch.float64, device='cuda'))

    def make_args(B):
        L = []
        cv = lambda i, src, **kw: L.append(ops.make_conv_args([src], convs[i % 8][0], 48, bias=convs[i % 8][1],
                                                              wlayout=wlayout, **kw))
        cur, nxt = B['x'], B['a']
        for blk in range(depth):
            cv(2 * blk, cur, out=B['t'], relu=True)                                     # t is rewritten every block (WAR)
            cv(2 * blk + 1, B['t'], out=nxt, res1=cur, res2=B['x'] if blk % 2 else None)
            cur, nxt = nxt, (B['b'] if nxt is B['a'] else B['a'])
        cv(5, cur, out=B['u'], relu=True)
        cv(6, B['u'], epilogue=_lib.LV_EPI_PS4_NCHW, out_hr=B['hr'], base_hr=base, truth_hr=truth, loss_sum=B['loss'],
           grad_sign=B['g'])
        cv(7, B['g'], out=B['d'], mask=B['u'])                                          # dgrad-style: ReLU mask of a saved act
        cv(3, B['d'], out=B['t'], res1=B['g'])
        return L
    return make_args, make_bufs


@pytest.mark.parametrize('wlayout', [0, 1], ids=['tile_chain', 'row_chain'])
@pytest.mark.parametrize('shape,ctas', [((2, 37, 45), 0), ((16, 48, 48), 0), ((1, 180, 320), 0), ((3, 20, 9), 5),
                                        ((1, 16, 8), 0), ((2, 135, 240), 0)])
def test_conv_chain_matches_sequential(shape, ctas, wlayout):
    """lv_conv3x3_chain (one persistent data-flow launch; tap-major weights -> 16x8-tile kernel, ky-stacked weights ->
    row-marching kernel) == the same layers launched one by one, bit for bit; repeated launches reuse the self-cleaning
    flag workspace."""
    n, h, w = shape
    rs = np.random.RandomState(5)
    make_args, make_bufs = _chain_layers(rs, n, h, w, depth=5, wlayout=wlayout)
    ref = make_bufs()
    for a in make_args(ref):
        ops.conv3x3_launch(a)
    torch.cuda.synchronize()
    ws = ops.chain_workspace(n, h, w, 'cuda')
    for rep in range(3):
        got = make_bufs()
        ops.conv3x3_chain(make_args(got), ws, max_ctas=ctas)
        torch.cuda.synchronize()
        assert int(ws.abs().sum().item()) == 0, 'flag workspace not re-zeroed'
        for k in ref:
            if k == 'loss':
                assert abs(got[k].item() - ref[k].item()) <= 1e-7 * abs(ref[k].item())   # fp32 partial sums regroup
            else:
                assert torch.equal(got[k], ref[k]), f'buffer {k} differs (rep {rep})'


def test_conv_chain_rejects_bad_layers():
    rs = np.random.RandomState(6)
    make_args, make_bufs = _chain_layers(rs, 1, 16, 8, depth=1)
    B = make_bufs()
    L = make_args(B)
    ws = ops.chain_workspace(1, 16, 8, 'cuda')
    bad = ops.make_conv_args([B['x']], L[0].weights and _pack(_rand_conv(rs, 48, 48, torch.bfloat16)[0], torch.bfloat16, 48),
                             48, out=B['t'])
    bad.h = 15
    with pytest.raises(_lib.LarvaNetB200Error):
        ops.conv3x3_chain([L[0], bad], ws)
    with pytest.raises(_lib.LarvaNetB200Error):
        ops.conv3x3_chain(L, torch.zeros(1, dtype=torch.int32, device='cuda'))


def test_device_side_crop_rot90_flip_matches_torch():
    """lv_crop_augment (dataloaders/synthetic_loader_tensor.py) == crop -> torch.rot90(k, dims=(1,2)) -> flip of the last axis,
    the reference's augmentation (dataloaders/div2k_train_loader_tensor.py:72-93), for every k and both flips."""
    import importlib
    ld = importlib.import_module('dataloaders.synthetic_loader_tensor').create_loader()
    ld.parse_args(['--synthetic_images=3', '--synthetic_height=40', '--synthetic_width=56'])
    ld.prepare(scales=[4])
    p = 12
    draws = [(i % 3, (5 * i) % 29, (7 * i) % 45, 1 + i % 4, (i // 4) % 2) for i in range(8)]
    x, t = ld.get_patch_batch(len(draws), 4, p, draws=draws)
    assert x.shape == (8, 3, p, p) and t.shape == (8, 3, 4 * p, 4 * p) and x.is_cuda
    for b, (idx, y, xx, rot, flip) in enumerate(draws):
        lr, hr = ld.dev_lr[4][idx], ld.dev_hr[4][idx]
        rl = torch.rot90(lr[:, y:y + p, xx:xx + p], k=rot, dims=(1, 2))
        rh = torch.rot90(hr[:, 4 * y:4 * (y + p), 4 * xx:4 * (xx + p)], k=rot, dims=(1, 2))
        if flip:
            rl, rh = torch.flip(rl, (2,)), torch.flip(rh, (2,))
        assert torch.equal(x[b], rl) and torch.equal(t[b], rh), (b, rot, flip)
    # random draws stay inside the images and reproduce with the seed
    a1 = ld.get_patch_batch(5, 4, 16)[0].clone()
    ld2 = importlib.import_module('dataloaders.synthetic_loader_tensor').create_loader()
    ld2.parse_args(['--synthetic_images=3', '--synthetic_height=40', '--synthetic_width=56'])
    ld2.prepare(scales=[4])
    ld2.get_patch_batch(len(draws), 4, p, draws=draws)
    assert torch.equal(a1, ld2.get_patch_batch(5, 4, 16)[0])


def test_uint8_and_psnr_helpers():
    """lv_image_to_uint8 == validate._image_to_uint8 (round-half-even, clip), lv_psnr_sqsum == _fit_truth + _image_psnr."""
    rs = np.random.RandomState(21)
    img = rs.uniform(-20, 280, (3, 37, 53)).astype(np.float32)
    img.flat[:8] = [0.5, 1.5, 2.5, 254.5, 255.5, -0.5, 127.49999, 300.0]      # ties and clip edges
    got = ops.image_to_uint8(torch.from_numpy(img).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got, O.image_to_uint8(img))
    odd = rs.uniform(0, 255, 1001).astype(np.float32)                            # tail that is not a multiple of 4
    np.testing.assert_array_equal(ops.image_to_uint8(torch.from_numpy(odd).cuda()).cpu().numpy(), O.image_to_uint8(odd))
    truth = rs.uniform(-5, 260, (3, 40, 60)).astype(np.float32)                  # larger than the output: cropped
    sq = torch.zeros(1, dtype=torch.float64, device='cuda')
    ops.psnr_sqsum(torch.from_numpy(img).cuda(), torch.from_numpy(truth).cuda(), sq)
    o8, t8 = O.image_to_uint8(img), O.image_to_uint8(truth)[:, :37, :53]
    ref_sq = float(((t8.astype(np.int64) - o8.astype(np.int64)) ** 2).sum())
    assert sq.item() == ref_sq
    psnr = 10.0 * np.log10(255.0 ** 2 / (sq.item() / img.size))
    assert abs(psnr - O.image_psnr(o8, t8)) < 1e-4


@pytest.mark.skipif(os.environ.get('LARVANET_B200_EXPERIMENTAL') != '1',
                    reason='tools/experiments kernels are only built with LARVANET_B200_EXPERIMENTAL=1')
def test_cluster_resident_strip_chain_is_bit_exact():
    """The opt-in cluster-resident chain (conv_strip.cu, LARVANET_B200_STRIP=1: activations in shared memory, halo columns
    through DSMEM) must give the per-layer launches' results bit for bit.  Runs in a subprocess: the switch is read once."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LARVANET_B200_STRIP='1')
    for shape in (('2', '37', '45'), ('3', '48', '48')):
        out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'strip_debug2.py'), *shape], env=env, cwd=root,
                             capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr[-400:]
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith('prefix')]
        assert len(lines) >= 10 and all(ln.rstrip().endswith('[]') for ln in lines), out.stdout[-600:]
