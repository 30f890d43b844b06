"""The reference's entry points, run as the reference documents them (README.md:28-63) but on the synthetic loaders:
train_larva.py / train_larvaV2.py (a few steps, checkpoint written), validate.py and get_sr.py on that checkpoint,
runtime.py.  `-m gpu` only: they drive the CUDA path end to end through the plugin API."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    env = dict(os.environ, PYTHONPATH=REPO)
    r = subprocess.run([sys.executable] + args, cwd=REPO, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + '\n' + r.stderr[-4000:]
    return r.stdout


@pytest.mark.parametrize('script,model', [('train_larva.py', 'LarvaNet'), ('train_larvaV2.py', 'LarvaNetV2')])
def test_train_validate_get_sr_runtime(tmp_path, script, model):
    import cv2 as cv
    train_dir = str(tmp_path / 'train')
    net = ['--model=' + model, '--num_modules=2', '--num_blocks=1,1']
    out = _run([script, '--batch_size=4', '--input_patch_size=24', '--max_steps=6', '--log_freq=2', '--sleep_ratio=0',
                '--train_path=' + train_dir, '--val_volume=20000', '--synthetic_images=2', '--synthetic_height=32',
                '--synthetic_width=40'] + net)
    assert 'begin training' in out and 'finished' in out and 'psnr=' in out
    losses = [float(l.split('loss ')[1].split(' ')[0]) for l in out.splitlines() if l.startswith('step ') and 'loss ' in l]
    assert losses and all(np.isfinite(losses))
    # the GPU-resident loader (device-side crop / rot90 / flip) feeds the same loop
    out2 = _run([script, '--dataloader=synthetic_loader_tensor', '--batch_size=4', '--input_patch_size=24', '--max_steps=4',
                 '--log_freq=2', '--sleep_ratio=0', '--train_path=' + str(tmp_path / 'train_t'), '--val_volume=1e12',
                 '--synthetic_images=2', '--synthetic_height=32', '--synthetic_width=40'] + net)
    assert 'finished' in out2 and 'loss ' in out2
    ckpts = sorted(glob.glob(os.path.join(train_dir, 'model_step*_vol*G.pth')))
    assert ckpts, out
    out = _run(['validate.py', '--restore_path=' + ckpts[-1], '--save_path=' + str(tmp_path / 'val'),
                '--synthetic_images=2'] + net)
    assert 'restored the model' in out and 'x4, psnr=' in out
    assert glob.glob(str(tmp_path / 'val' / 'x4' / '*.png'))
    out = _run(['validate.py', '--restore_path=' + ckpts[-1], '--chop_forward', '--chop_overlap_size=20'] + net)
    assert 'x4, psnr=' in out
    # exact band-sharded inference gives the full-frame PSNR to the last digit
    full = [l for l in _run(['validate.py', '--restore_path=' + ckpts[-1]] + net).splitlines() if l.startswith('x4, psnr=')]
    band = [l for l in _run(['validate.py', '--restore_path=' + ckpts[-1], '--exact_bands=3'] + net).splitlines()
            if l.startswith('x4, psnr=')]
    assert full and band and full[-1].split(',')[1] == band[-1].split(',')[1], (full, band)
    lr_dir = tmp_path / 'lr'
    lr_dir.mkdir()
    rs = np.random.RandomState(0)
    for i in range(2):
        cv.imwrite(str(lr_dir / f'img{i}.png'), rs.randint(0, 256, (20 + i, 28, 3), dtype=np.uint8))
    out = _run(['get_sr.py', '--restore_path=' + ckpts[-1], '--input_path=' + str(lr_dir),
                '--output_path=' + str(tmp_path / 'sr')] + net)
    assert 'finished' in out
    sr = cv.imread(str(tmp_path / 'sr' / 'img1.png'))
    assert sr.shape == (84, 112, 3)
    out = _run(['runtime.py'] + net + ['--synthetic_images=2'])
    assert 'runtime=' in out and 'finished' in out
