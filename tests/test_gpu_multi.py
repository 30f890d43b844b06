"""Multi-GPU checks (`-m gpu`, skipped on boxes with fewer than 2 GPUs; run them with `gpurun --gpus 2`):
data-parallel gradients == single-process gradients on the global batch, and `torchrun train_larva.py` shards the
batch instead of training silent replicas (reference step: models/LarvaNet.py:102-114)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
needs2 = pytest.mark.skipif(NGPU < 2, reason='needs >= 2 GPUs (peer kernels of different ranks must run concurrently)')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(nproc, args, timeout=600, env_extra=None):
    env = dict(os.environ, PYTHONPATH=REPO)
    env.update(env_extra or {})
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port())] + args
    r = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + '\n' + r.stderr[-5000:]
    return r.stdout


@needs2
@pytest.mark.parametrize('symm,two_shot', [('1', '0'), ('1', '1'), ('0', '0')],
                         ids=['peer_memory_one_shot', 'peer_memory_two_shot', 'nccl_allreduce'])
def test_dp_gradients_equal_single_process(symm, two_shot):
    """tools/dp_check.py: gradients of 2 ranks x half the batch == one process on the whole batch (also uneven shards), and
    four optimizer steps through the fused exchange + AdamW + re-pack kernel (one-shot and two-shot forms) track a single
    process and leave bit-identical replicas."""
    out = _torchrun(2, ['tools/dp_check.py'], env_extra={'LARVANET_B200_SYMM_ALLREDUCE': symm,
                                                          'LARVANET_B200_DP_TWO_SHOT': two_shot})
    assert out.count('rel grad diff') == 3, out
    assert 'replicas bit-identical: True' in out, out
    if symm == '1':
        assert 'fused=True' in out, out


@needs2
@pytest.mark.parametrize('script,model', [('train_larva.py', 'LarvaNet'), ('train_larvaV2.py', 'LarvaNetV2')])
def test_torchrun_train_shards_the_batch(tmp_path, script, model):
    out = _torchrun(2, [script, '--model=' + model, '--num_modules=2', '--num_blocks=1,1', '--batch_size=6',
                        '--input_patch_size=24', '--max_steps=6', '--log_freq=2', '--sleep_ratio=0',
                        '--train_path=' + str(tmp_path), '--val_volume=20000', '--synthetic_images=2',
                        '--synthetic_height=32', '--synthetic_width=40'])
    assert 'data parallel: 2 ranks x 3 patches (global batch 6)' in out
    assert out.count('begin training') == 1 and 'finished' in out      # rank 0 reports, the others stay quiet
    import glob
    assert glob.glob(os.path.join(str(tmp_path), 'model_step*_vol*G.pth'))
