"""Per-launch CUDA-event times of one eager EDSR-baseline 1080p forward (which layers cost what)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import ops, synth
m = importlib.import_module('models.edsr').create_model()
m.parse_args(['--edsr_conv_features=64', '--edsr_res_blocks=16'])
m.prepare(is_training=False, scales=[4])
m.get_model().load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_edsr_params(64, 16, 4, seed=0).items()})
eng = m.get_model().engine()
eng.use_graphs = False
x = torch.from_numpy(synth.make_images(1, 270, 480, seed=7)[0]).cuda()
for _ in range(3):
    eng.forward(x)
torch.cuda.synchronize()
ops.CONV_TIMERS = []
torch.cuda._sleep(int(4e7))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.forward(x)
e1.record()
torch.cuda.synchronize()
recs, ops.CONV_TIMERS = ops.CONV_TIMERS, None
tot = 0.0
for a, b, fl, tag in recs:
    us = a.elapsed_time(b) * 1e3
    tot += us
    print(f'{str(tag):>24} {us:9.1f} us  {fl / us / 1e6:8.1f} TFLOP/s')
print(f'conv launches {tot:.1f} us of {e0.elapsed_time(e1) * 1e3:.1f} us total')
