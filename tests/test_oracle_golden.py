"""Pin the numpy oracle against outputs of the unmodified reference modules (tests/golden/*.npz,
made by tests/golden/make_golden.py).  fp64 reference outputs pin the algorithm (<=1e-9 relative);
fp32 reference outputs are the reference's own numerics (<=2e-5 relative on the 0..255 scale)."""
import os

import numpy as np
import pytest

from larvanet_b200 import synth
from oracle import larva_oracle as O

LARVA_CASES = ['larvanet_m2_b21', 'larvanet_m3_b111', 'larvanetv2_m2_b11', 'larvanetv2_m4_b1111']
EDSR_CASES = ['edsr_f64_b2', 'edsr_f16_b3']


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + '.npz'), allow_pickle=False)


def _grad_summary(g):
    g = np.asarray(g, dtype=np.float64).ravel()
    samp = np.zeros(32)
    s = g[::97][:32]
    samp[:s.size] = s
    return np.concatenate([[g.sum(), np.abs(g).sum(), np.sqrt((g * g).sum())], g[:8], samp])


@pytest.mark.parametrize('name', LARVA_CASES)
def test_larva_forward_and_step(golden_dir, name):
    g = _load(golden_dir, name)
    v2 = bool(g['v2'])
    blocks = [int(b) for b in g['blocks']]
    params = synth.make_larva_params(blocks, v2=v2, seed=int(g['seed']), bias_std=0.02)
    assert list(params.keys()) == [str(k) for k in g['state_dict_keys']]
    lr, hr = synth.make_images(int(g['n']), int(g['h']), int(g['w']), seed=int(g['seed']) + 100)

    fwd = O.larvanet_v2_forward if v2 else O.larvanet_forward
    out = fwd(params, lr, blocks)
    np.testing.assert_allclose(out, g['out_f64'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out, g['out_f32'], rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(O.bicubic_upsample(lr.astype(np.float64)), g['base_f32'], rtol=0, atol=2e-4)
    np.testing.assert_allclose(O.head_forward(params, lr), g['head_f32'], rtol=1e-5, atol=1e-4)

    loss, grads, outs = O.larvanet_train_step(params, lr, hr, blocks, v2=v2)
    assert abs(loss - float(g['loss_f64'])) <= 1e-10 * abs(loss)
    assert abs(loss - float(g['loss_f32'])) <= 1e-5 * abs(loss)
    np.testing.assert_allclose(np.stack(outs), g['exits_f32'], rtol=2e-5, atol=2e-3)
    summ = np.stack([_grad_summary(grads[k]) for k in params.keys()])
    ref = g['grad_summary_f64']
    scale = np.abs(ref[:, 2:3]) + 1e-30     # per-tensor L2 norm
    assert np.max(np.abs(summ - ref) / scale) < 1e-8
    ref32 = g['grad_summary_f32']
    # sign() makes the L1 gradient discontinuous: fp32 vs fp64 may flip a few signs -> loose
    assert np.max(np.abs(summ[:, 2] - ref32[:, 2]) / np.abs(ref32[:, 2])) < 5e-3


@pytest.mark.parametrize('name', ['larvanet_m2_b21', 'larvanet_m3_b111'])
def test_early_exit(golden_dir, name):
    g = _load(golden_dir, name)
    blocks = [int(b) for b in g['blocks']]
    params = synth.make_larva_params(blocks, seed=int(g['seed']), bias_std=0.02)
    lr, _ = synth.make_images(int(g['n']), int(g['h']), int(g['w']), seed=int(g['seed']) + 100)
    for k in range(len(blocks) + 1):
        out = O.larvanet_forward(params, lr, blocks, exit_leg=k)
        np.testing.assert_allclose(out, g[f'exit_leg{k}_f32'], rtol=2e-5, atol=2e-3)


@pytest.mark.parametrize('name', EDSR_CASES)
def test_edsr_forward(golden_dir, name):
    g = _load(golden_dir, name)
    params = synth.make_edsr_params(int(g['features']), int(g['res_blocks']), 4, seed=int(g['seed']))
    assert list(params.keys()) == [str(k) for k in g['state_dict_keys']]
    lr, _ = synth.make_images(int(g['n']), int(g['h']), int(g['w']), seed=int(g['seed']) + 100)
    out = O.edsr_forward(params, lr, num_res_blocks=int(g['res_blocks']))
    np.testing.assert_allclose(out, g['out_f64'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out, g['out_f32'], rtol=1e-4, atol=1e-2)


def test_known_answers():
    """Invariants derivable from the reference code."""
    # (i) all-zero conv weights => output == bicubic base exactly (LarvaLeg.forward :263-267)
    blocks = [1, 1]
    params = {k: np.zeros_like(v) for k, v in synth.make_larva_params(blocks).items()}
    lr, _ = synth.make_images(1, 6, 7, seed=3)
    out = O.larvanet_forward(params, lr, blocks)
    np.testing.assert_array_equal(out, O.bicubic_upsample(lr.astype(np.float64)))
    # (ii) PixelShuffle index identity out[n,c,4h+i,4w+j] = in[n,16c+4i+j,h,w]
    x = np.arange(2 * 48 * 3 * 5, dtype=np.float64).reshape(2, 48, 3, 5)
    y = O.pixel_shuffle(x, 4)
    for (n, c, h, w, i, j) in [(0, 0, 0, 0, 0, 0), (1, 2, 2, 4, 3, 1), (0, 1, 1, 3, 2, 2)]:
        assert y[n, c, 4 * h + i, 4 * w + j] == x[n, 16 * c + 4 * i + j, h, w]
    np.testing.assert_array_equal(O.pixel_unshuffle(y, 4), x)
    # bicubic phase weights
    w0 = O._cubic_coeffs(np.array(0.625))
    np.testing.assert_allclose(w0, [-0.06591797, 0.42626953, 0.74951172, -0.10986328], atol=1e-8)
    # constant image stays constant (weights sum to 1, clamped borders)
    c = np.full((1, 3, 5, 4), 77.0)
    np.testing.assert_allclose(O.bicubic_upsample(c), 77.0, atol=1e-12)
    # PSNR helper: identical images -> inf, known mse
    a = np.zeros((3, 4, 4), np.uint8)
    b = np.full((3, 4, 4), 5, np.uint8)
    assert abs(O.image_psnr(a, b) - 10 * np.log10(255.0 ** 2 / 25.0)) < 1e-5  # float32 metric


def test_conv_backward_matches_finite_difference():
    rs = np.random.RandomState(5)
    x = rs.standard_normal((1, 3, 5, 4))
    w = rs.standard_normal((2, 3, 3, 3))
    dy = rs.standard_normal((1, 2, 5, 4))
    dx, dw, db = O.conv2d_backward(x, w, dy)
    eps = 1e-6
    for idx in [(0, 1, 2, 3), (0, 0, 0, 0), (0, 2, 4, 1)]:
        xp = x.copy(); xp[idx] += eps
        num = ((O.conv2d(xp, w) - O.conv2d(x, w)) * dy).sum() / eps
        assert abs(num - dx[idx]) < 1e-4
    for idx in [(1, 2, 0, 2), (0, 0, 1, 1)]:
        wp = w.copy(); wp[idx] += eps
        num = ((O.conv2d(x, wp) - O.conv2d(x, w)) * dy).sum() / eps
        assert abs(num - dw[idx]) < 1e-4
    np.testing.assert_allclose(db, dy.sum((0, 2, 3)))


@pytest.mark.parametrize('name', ['larvanet_m2_b21', 'larvanetv2_m2_b11'])
def test_torch_cpu_port_matches_golden(golden_dir, name):
    """The timed CPU baseline (oracle/torch_port.py) computes the same function as the reference modules."""
    import torch
    from oracle import torch_port
    g = _load(golden_dir, name)
    v2 = bool(g['v2'])
    blocks = [int(b) for b in g['blocks']]
    params = synth.make_larva_params(blocks, v2=v2, seed=int(g['seed']), bias_std=0.02)
    lr, hr = synth.make_images(int(g['n']), int(g['h']), int(g['w']), seed=int(g['seed']) + 100)
    tr = torch_port.CpuTrainer(params, blocks, v2=v2, threads=2)
    out = tr.infer(torch.from_numpy(lr)).numpy()
    np.testing.assert_allclose(out, g['out_f32'], rtol=1e-5, atol=1e-3)
    loss = torch_port.loss_fn(tr.p, torch.from_numpy(lr), torch.from_numpy(hr), blocks, v2).item()
    assert abs(loss - float(g['loss_f32'])) <= 1e-5 * loss
