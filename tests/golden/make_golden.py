"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case the script builds the reference nn.Module (`/root/reference/models/LarvaNet.py`,
`LarvaNetV2.py`, `LarvaLeg.py`, `edsr.py`), loads weights produced by `larvanet_b200.synth`
(frozen numpy RandomState, so only the seed is stored), runs it on the CPU in fp32 (the reference's
own numerics) and in fp64 (pins the algorithm to ~1e-12), and stores inputs' seeds, outputs, the
multi-exit loss, and compact summaries of every parameter gradient.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'


def _import_reference():
    # import the reference's packages under their own names, isolated from this repo's `models/`
    for name in list(sys.modules):
        if name == 'models' or name.startswith('models.') or name in ('validate', 'dataloaders', 'utils'):
            del sys.modules[name]
    sys.path.insert(0, REF)
    import importlib
    mods = {n: importlib.import_module('models.' + n) for n in ('LarvaNet', 'LarvaNetV2', 'LarvaLeg', 'LarvaLegV2', 'edsr')}
    sys.path.remove(REF)
    for name in list(sys.modules):
        if name == 'models' or name.startswith('models.') or name in ('validate', 'dataloaders', 'utils') \
                or name.startswith('dataloaders.') or name.startswith('utils.'):
            del sys.modules[name]
    return mods


def grad_summary(g):
    g = np.asarray(g, dtype=np.float64).ravel()
    samp = np.zeros(32)
    s = g[::97][:32]
    samp[:s.size] = s
    return np.concatenate([[g.sum(), np.abs(g).sum(), np.sqrt((g * g).sum())], g[:8], samp])


def larva_case(mods, name, v2, blocks, n, h, w, seed, out_dir):
    sys.path.insert(0, REPO)
    from larvanet_b200 import synth
    params = synth.make_larva_params(blocks, v2=v2, seed=seed, bias_std=0.02)
    lr, hr = synth.make_images(n, h, w, seed=seed + 100)
    args = types.SimpleNamespace(num_modules=len(blocks), num_blocks=','.join(map(str, blocks)), interpolate='bicubic')
    mod = mods['LarvaNetV2' if v2 else 'LarvaNet'].LarvaNetModule(args)
    sd = {k: torch.from_numpy(v) for k, v in params.items()}
    assert list(mod.state_dict().keys()) == list(sd.keys()), 'state_dict key order mismatch'
    mod.load_state_dict(sd)
    store = dict(v2=v2, blocks=np.array(blocks), n=n, h=h, w=w, seed=seed,
                 state_dict_keys=np.array(list(sd.keys())))
    for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        m = mod.to(dt)
        x = torch.from_numpy(lr).to(dt)
        t = torch.from_numpy(hr).to(dt)
        with torch.no_grad():
            store[f'out_{tag}'] = m(x).numpy()
            if tag == 'f32':
                store[f'base_{tag}'] = m.base(x).numpy()
                store[f'head_{tag}'] = m.head(x).numpy()
        # the training step core, restated from reference models/LarvaNet.py:102-113 and
        # models/LarvaNetV2.py:105-118 using the reference module's own sub-modules
        loss_fn = torch.nn.L1Loss()
        m.zero_grad()
        fea = m.head(x)
        base = m.base(x)
        loss = 0
        feats = []
        exits = []
        for i in range(len(blocks)):
            fea = getattr(m, f'body_{i}')(fea)
            feats.append(fea)
            out = getattr(m, f'body_{i}').leg(fea, base)
            exits.append(out.detach().numpy().copy())
            loss = loss + loss_fn(out, t)
        if v2:
            out = m.tail(feats, base)
            exits.append(out.detach().numpy().copy())
            loss = loss + loss_fn(out, t)
            loss = loss / (len(blocks) + 1)
        else:
            loss = loss / len(blocks)
        loss.backward()
        store[f'loss_{tag}'] = float(loss.item())
        if tag == 'f32':
            store[f'exits_{tag}'] = np.stack(exits)
            store[f'feat_last_{tag}'] = feats[-1].detach().numpy()
        store[f'grad_summary_{tag}'] = np.stack([grad_summary(p.grad.numpy()) for _, p in m.named_parameters()])
    # early exits: reference models/LarvaLeg.py:290-299 and models/LarvaLegV2.py:357-368 (--leg=k)
    for k in range(len(blocks) + 1):
        largs = types.SimpleNamespace(num_modules=len(blocks), num_blocks=args.num_blocks,
                                      interpolate='bicubic', leg=k)
        lm = mods['LarvaLegV2' if v2 else 'LarvaLeg'].LarvaNetModule(largs)
        lm.load_state_dict(sd)
        with torch.no_grad():
            store[f'exit_leg{k}_f32'] = lm(torch.from_numpy(lr)).numpy()
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), **store)
    print('wrote', name, 'loss', store['loss_f32'])


def edsr_case(mods, name, features, res_blocks, n, h, w, seed, out_dir):
    from larvanet_b200 import synth
    params = synth.make_edsr_params(features, res_blocks, 4, seed=seed)
    lr, _ = synth.make_images(n, h, w, seed=seed + 100)
    args = types.SimpleNamespace(edsr_conv_features=features, edsr_res_blocks=res_blocks, edsr_res_weight=1.0)
    mod = mods['edsr'].EDSRModule(args, scale=4)
    sd = {k: torch.from_numpy(v) for k, v in params.items()}
    assert list(mod.state_dict().keys()) == list(sd.keys()), 'state_dict key order mismatch'
    mod.load_state_dict(sd)
    store = dict(features=features, res_blocks=res_blocks, n=n, h=h, w=w, seed=seed,
                 state_dict_keys=np.array(list(sd.keys())))
    for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        with torch.no_grad():
            store[f'out_{tag}'] = mod.to(dt)(torch.from_numpy(lr).to(dt)).numpy()
    np.savez_compressed(os.path.join(out_dir, name + '.npz'), **store)
    print('wrote', name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=HERE)
    a = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(4)
    mods = _import_reference()
    sys.path.insert(0, REPO)
    larva_case(mods, 'larvanet_m2_b21', False, [2, 1], 2, 12, 10, 11, a.out)
    larva_case(mods, 'larvanet_m3_b111', False, [1, 1, 1], 1, 12, 16, 12, a.out)
    larva_case(mods, 'larvanetv2_m2_b11', True, [1, 1], 2, 16, 8, 13, a.out)
    larva_case(mods, 'larvanetv2_m4_b1111', True, [1, 1, 1, 1], 1, 8, 16, 14, a.out)
    edsr_case(mods, 'edsr_f64_b2', 64, 2, 1, 8, 16, 15, a.out)
    edsr_case(mods, 'edsr_f16_b3', 16, 3, 2, 6, 5, 16, a.out)


if __name__ == '__main__':
    main()
