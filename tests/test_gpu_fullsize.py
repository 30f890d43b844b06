"""Parity at the BASELINE.json configurations' OWN sizes (`-m gpu`): the CUDA path against the reference's modules run
in fp32 on the host of the GPU box.

The comparator is the UNMODIFIED reference (`oracle/_ref/`, staged by oracle/make_ref.py and shipped with the snapshot;
oracle/ref_loader.py drives it).  If that directory is missing the golden-pinned restatement oracle/torch_port.py is
used instead and the test says so.  Sizes: cfg1 = 1 x 320x180 LarvaNet inference, cfg2 = one training step of 16 x 48x48,
cfg3 = EDSR-baseline 480x270 -> 1080p, cfg4 = one LarvaNetV2 training step of 16 x 64x64 (the per-GPU shard of global
batch 128 on 8 GPUs), cfg5 = 2 of the 480x270 frames; plus a 50-step loss trajectory of the bf16 path against the
reference optimiser on identical batches.

Tolerances (north star): bf16 -- max |err| <= 2.0 on the 0..255 HR image, |dPSNR| <= 0.01 dB; fp32 mode -- 1e-4
relative.  Gradients: see tests/test_gpu_network.py (2e-2 on identical forward state is checked there on small shapes;
here the end-to-end figures against the reference's autograd are asserted with the discontinuity budget and printed).
"""
import importlib

import numpy as np
import pytest
import torch

from larvanet_b200 import synth
from oracle import larva_oracle as O
from oracle import ref_loader, torch_port
from tests.gpu_util import load_params, rel_l2

pytestmark = pytest.mark.gpu

BLOCKS = [4, 4, 4, 4]


def _make(kind, blocks, precision, training=False, extra=()):
    m = importlib.import_module('models.' + kind).create_model()
    m.parse_args([f'--num_modules={len(blocks)}', '--num_blocks=' + ','.join(map(str, blocks)),
                  f'--precision={precision}', *extra])
    m.prepare(is_training=training, scales=[4])
    return m


def _ref_kind():
    return 'reference modules (oracle/_ref)' if ref_loader.load() is not None else 'port (oracle/torch_port.py)'


def _ref_forward(params, lr, blocks, v2=False):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    x = torch.from_numpy(lr)
    if ref_loader.load() is not None:
        with torch.no_grad():
            return ref_loader.make_module(blocks, v2, params)(x).numpy()
    p = {k: torch.from_numpy(v) for k, v in params.items()}
    with torch.no_grad():
        return torch_port.forward(p, x, blocks, v2).numpy()


def _ref_step(params, lr, hr, blocks, v2=False):
    """(loss, {name: grad}) of the reference's multi-exit step, fp32 on the host."""
    x, t = torch.from_numpy(lr), torch.from_numpy(hr)
    if ref_loader.load() is not None:
        tr = ref_loader.RefTrainer(params, blocks, v2)
        loss = tr.loss_and_backward(x, t)
        return float(loss.item()), tr.grads()
    tr = torch_port.CpuTrainer(params, blocks, v2)
    loss = torch_port.loss_fn(tr.p, x, t, blocks, v2)
    loss.backward()
    return float(loss.item()), {k: v.grad.numpy() for k, v in tr.p.items()}


def _check_image(out, ref, hr, precision):
    assert out.shape == ref.shape and np.isfinite(out).all()
    err = float(np.max(np.abs(out - ref)))
    if precision == 'fp32':
        assert err <= 1e-4 * 255, err
    else:
        assert err <= 2.0, err
        p_ref = O.image_psnr(O.image_to_uint8(ref), O.image_to_uint8(hr))
        p_out = O.image_psnr(O.image_to_uint8(out), O.image_to_uint8(hr))
        assert abs(p_ref - p_out) <= 0.01, (p_ref, p_out)
    return err


@pytest.mark.parametrize('precision', ['bf16', 'fp32'])
def test_cfg1_720p_inference_matches_reference(precision):
    """BASELINE configs[0]: LarvaNet x4, batch 1, 320x180 -> 1280x720 (reference models/LarvaNet.py:287-293)."""
    params = synth.make_larva_params(BLOCKS, seed=0, bias_std=0.02)
    lr, hr = synth.make_smooth_images(1, 180, 320, seed=1)
    m = _make('LarvaNet', BLOCKS, precision)
    load_params(m.get_model(), params)
    out = m.upscale(list(lr), 4)
    err = _check_image(out, _ref_forward(params, lr, BLOCKS), hr, precision)
    print(f'cfg1 {precision}: max|err| {err:.4f} on 0..255 vs {_ref_kind()}')


@pytest.mark.parametrize('v2,n,p', [(False, 16, 48), (True, 16, 64)], ids=['cfg2_larvanet_16x48', 'cfg4_v2_16x64'])
@pytest.mark.parametrize('precision', ['bf16', 'fp32'])
def test_full_training_step_matches_reference_autograd(v2, n, p, precision):
    """BASELINE configs[1] / [3] (per-GPU shard): loss and every parameter gradient of ONE full-size step against the
    reference's autograd (models/LarvaNet.py:102-113, models/LarvaNetV2.py:105-120)."""
    params = synth.make_larva_params(BLOCKS, v2=v2, seed=0, bias_std=0.02)
    lr, hr = synth.make_smooth_images(n, p, p, seed=2)
    m = _make('LarvaNetV2' if v2 else 'LarvaNet', BLOCKS, precision, training=True)
    load_params(m.get_model(), params)
    eng = m._engine()
    loss = eng.train_step(torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda()).item()
    ref_loss, ref_grads = _ref_step(params, lr, hr, BLOCKS, v2)
    ltol = 1e-5 if precision == 'fp32' else 2e-3
    assert abs(loss - ref_loss) <= ltol * abs(ref_loss), (loss, ref_loss)
    worst_r, worst_c = 0.0, 1.0
    num = den = 0.0
    for name, prm in m.get_model().named_parameters():
        got = prm.grad.cpu().numpy().astype(np.float64)
        ref = ref_grads[name].astype(np.float64)
        r = rel_l2(got, ref)
        c = float(np.sum(got * ref) / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-300))
        worst_r, worst_c = max(worst_r, r), min(worst_c, c)
        num += float(np.sum((got - ref) ** 2))
        den += float(np.sum(ref ** 2))
        # fp32 mode has no activation rounding, hence no ReLU-mask / L1-sign flips: tight.  bf16: measured on B200 at
        # these sizes (band-limited images): whole arena 3.5e-3, worst tensor 9e-3, worst cosine 0.99996 -- the 2.5-5 %
        # seen in test_gpu_network come from white-noise inputs on 12x10-pixel images, where a flipped ReLU mask / L1
        # sign is a large fraction of the few hundred terms of a gradient
        assert r <= (2e-3 if precision == 'fp32' else 2e-2), (name, r)      # rel-L2 <= 2e-2 for bf16
        assert c >= (0.99999 if precision == 'fp32' else 0.9995), (name, c)
    total = (num / den) ** 0.5
    assert total <= (1e-3 if precision == 'fp32' else 1e-2), total
    print(f'{"cfg4" if v2 else "cfg2"} {precision}: loss {loss:.6f} vs {ref_loss:.6f}; grads vs {_ref_kind()}: '
          f'whole-arena rel-L2 {total:.3e}, worst tensor rel-L2 {worst_r:.3e}, worst cos {worst_c:.6f}')


@pytest.mark.parametrize('precision', ['bf16', 'fp32'])
def test_cfg3_edsr_1080p_matches_reference(precision):
    """BASELINE configs[2]: EDSR-baseline x4 (16 resblocks, 64 ch), 480x270 -> 1920x1080 (reference models/edsr.py:
    195-207).  Untrained default-init weights amplify the signal, so the bf16 bound is relative to the output scale
    (2/255 of it), as in the small-shape golden test."""
    params = synth.make_edsr_params(64, 16, 4, seed=0)
    lr, _ = synth.make_smooth_images(1, 270, 480, seed=3)
    m = importlib.import_module('models.edsr').create_model()
    m.parse_args(['--edsr_conv_features=64', '--edsr_res_blocks=16', f'--precision={precision}'])
    m.prepare(is_training=False, scales=[4])
    load_params(m.get_model(), params)
    out = m.upscale(list(lr), 4)
    if ref_loader.load() is not None:
        with torch.no_grad():
            ref = ref_loader.make_edsr(64, 16, params)(torch.from_numpy(lr)).numpy()
    else:
        ref = O.edsr_forward(params, lr[:, :, :32, :32], 64, 16).astype(np.float32)   # numpy oracle: a crop only
        out = m.upscale(list(lr[:, :, :32, :32]), 4)
    assert out.shape == ref.shape and np.isfinite(out).all()
    scale = float(np.abs(ref).max())
    err = float(np.max(np.abs(out - ref)))
    if precision == 'fp32':
        assert err <= 1e-4 * max(255.0, scale), (err, scale)
    else:
        assert err <= max(2.0, 2.0 / 255.0 * scale), (err, scale)
        r_row = rel_l2(out, ref)
        assert r_row <= 2e-2, r_row                      # 39 bf16 layers with amplifying default-init weights
        # the row-marching body chain (taken at this size) must be as accurate as the 16x8-tile kernels
        import os
        os.environ['LARVANET_B200_ROW'] = '0'
        try:
            m2 = importlib.import_module('models.edsr').create_model()
            m2.parse_args(['--edsr_conv_features=64', '--edsr_res_blocks=16', '--precision=bf16'])
            m2.prepare(is_training=False, scales=[4])
            load_params(m2.get_model(), params)
            r_tile = rel_l2(m2.upscale(list(lr), 4), ref)
        finally:
            del os.environ['LARVANET_B200_ROW']
        print(f'cfg3 bf16 rel-L2 vs reference: row path {r_row:.3e}, tile path {r_tile:.3e}')
        assert r_row <= 1.5 * r_tile + 1e-4, (r_row, r_tile)
    print(f'cfg3 {precision}: max|err| {err:.4f} at output scale {scale:.1f} vs {_ref_kind()}')


def test_cfg5_frames_match_reference():
    """BASELINE configs[4] frame size (480x270), two frames as one batch through the plugin."""
    params = synth.make_larva_params(BLOCKS, seed=0, bias_std=0.02)
    lr, hr = synth.make_smooth_images(2, 270, 480, seed=4)
    m = _make('LarvaNet', BLOCKS, 'bf16')
    load_params(m.get_model(), params)
    out = m.upscale(list(lr), 4)
    err = _check_image(out, _ref_forward(params, lr, BLOCKS), hr, 'bf16')
    print(f'cfg5 bf16: max|err| {err:.4f} vs {_ref_kind()}')


def test_loss_trajectory_50_steps_tracks_reference_optimizer():
    """50 optimiser steps of the bf16 product path (fused fwd+bwd, FusedAdamW + re-pack) against the reference's own
    step (modules + torch.optim.AdamW, fp32, host) on IDENTICAL batches: the per-step gradient differences caused by
    bf16 ReLU-mask / L1-sign flips must behave as unbiased noise, i.e. the two loss curves stay together and the
    signed deviation averages out."""
    blocks = BLOCKS
    params = synth.make_larva_params(blocks, seed=0, bias_std=0.0)
    pool = [synth.make_smooth_images(16, 48, 48, seed=50 + i) for i in range(5)]
    m = _make('LarvaNet', blocks, 'bf16', training=True, extra=('--lr=4e-4',))
    load_params(m.get_model(), params)
    eng = m._engine()
    if ref_loader.load() is not None:
        ref = ref_loader.RefTrainer(params, blocks, lr=4e-4)
    else:
        ref = torch_port.CpuTrainer(params, blocks, lr=4e-4)
    dev_pool = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in pool]
    host_pool = [(torch.from_numpy(a), torch.from_numpy(b)) for a, b in pool]
    ours, theirs = [], []
    for s in range(50):
        x, t = dev_pool[s % 5]
        loss = eng.train_step(x, t)
        m.optim.step()
        ours.append(loss.item())
        theirs.append(ref.step(*host_pool[s % 5]))
    ours, theirs = np.array(ours), np.array(theirs)
    rel = (ours - theirs) / theirs
    print(f'trajectory vs {_ref_kind()}: loss {theirs[0]:.4f} -> {theirs[-1]:.4f} (ref), {ours[0]:.4f} -> {ours[-1]:.4f} '
          f'(bf16); max |rel dev| {np.abs(rel).max():.3e}, mean signed rel dev {rel.mean():.3e}')
    assert theirs[-1] < theirs[0] * 0.9            # the run actually trains
    assert np.abs(rel).max() <= 2e-2               # curves stay together at every step
    assert abs(rel.mean()) <= 5e-3                 # and the deviation has no systematic sign
    # the weights after 50 steps agree as well as AdamW's sign-like updates allow
    sd = m.get_model().state_dict()
    refp = ref.params() if hasattr(ref, 'params') else {k: v.detach().numpy() for k, v in ref.p.items()}
    num = sum(float(np.sum((sd[k].cpu().numpy().astype(np.float64) - refp[k]) ** 2)) for k in refp)
    den = sum(float(np.sum((refp[k].astype(np.float64) - params[k]) ** 2)) for k in refp)
    assert (num / den) ** 0.5 <= 0.5, (num / den) ** 0.5   # distance between the two runs << distance travelled
