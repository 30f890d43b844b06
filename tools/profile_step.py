"""Workload for `ncu`: one eager LarvaNet training step at BASELINE configs[1] (16 x 48x48), one 720p inference frame, one
EDSR 1080p frame and a 4-layer row-marching chain on 8 x 270x480 -- every kernel of the path launches at least once after
warm-up.  Run plain first, then under ncu (see tools/gpu_profile.sh)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from larvanet_b200 import _lib, ops, synth  # noqa: E402
from tools.row_vs_tile import build  # noqa: E402

BLOCKS = [4, 4, 4, 4]
os.environ['LARVANET_B200_GRAPHS'] = '0'      # eager launches: ncu sees every kernel by name


def main():
    dev = torch.device('cuda', 0)
    m = importlib.import_module('models.LarvaNet').create_model()
    m.parse_args(['--num_modules=4', '--num_blocks=4,4,4,4'])
    m.prepare(is_training=True, scales=[4])
    m.get_model().load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_larva_params(BLOCKS, seed=0).items()})
    eng = m._engine()
    lr, hr = synth.make_images(16, 48, 48, seed=1)
    x, t = torch.from_numpy(lr).to(dev), torch.from_numpy(hr).to(dev)
    frame = torch.from_numpy(synth.make_images(1, 180, 320, seed=2)[0]).to(dev)
    e = importlib.import_module('models.edsr').create_model()
    e.parse_args(['--edsr_conv_features=64', '--edsr_res_blocks=16'])
    e.prepare(is_training=False, scales=[4])
    e.get_model().load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_edsr_params(64, 16, 4, seed=0).items()})
    ee = e.get_model().engine()
    f1080 = torch.from_numpy(synth.make_images(1, 270, 480, seed=3)[0]).to(dev)
    args, keep = build(8, 270, 480, 4, _lib.LV_W_KY_STACKED, dev)
    ws = ops.chain_workspace(8, 270, 480, dev)
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    for _ in range(reps):
        eng.train_step(x, t)
        m.optim.step()
        eng.forward(frame)
        ee.forward(f1080)
        ops.conv3x3_chain(args, ws)
    torch.cuda.synchronize()
    print('profile workload done,', _lib.launch_count(), 'launches')


if __name__ == '__main__':
    main()
