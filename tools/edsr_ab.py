"""EDSR-baseline 1080p inference, body chain on the row-marching kernel vs per-layer tile launches (LARVANET_B200_ROW=1/0)."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import synth
for mode in ('0', '1'):
    os.environ['LARVANET_B200_ROW'] = mode
    for n in (1, 4):
        m = importlib.import_module('models.edsr').create_model()
        m.parse_args(['--edsr_conv_features=64', '--edsr_res_blocks=16'])
        m.prepare(is_training=False, scales=[4])
        m.get_model().load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_edsr_params(64, 16, 4, seed=0).items()})
        eng = m.get_model().engine()
        x = torch.from_numpy(synth.make_images(n, 270, 480, seed=7)[0]).cuda()
        for _ in range(4):
            eng.forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.forward(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * 1983321 * 270 * 480 * n
        print(f'ROW={mode} batch {n}: {ms:.3f} ms per batch, {n * 16 * 270 * 480 / ms / 1e3:.0f} Mpix/s, {fl / ms / 1e9:.0f} TFLOP/s', flush=True)
        del m, eng, x
        torch.cuda.empty_cache()
