"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header declares, the plugin
surface mirrors the reference's (flags, state_dict layout, error behaviour), and computing without a GPU fails loudly."""
import importlib
import os
import re

import numpy as np
import pytest
import torch

from larvanet_b200 import _lib, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, 'include', 'larvanet_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(lv_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()   # binds every symbol or raises
    assert lib.lv_abi_version() == _lib.ABI_VERSION
    assert lib.lv_packed_weight_bytes(48, 48, _lib.LV_BF16) == 48 * 48 * 9 * 2
    assert lib.lv_packed_weight_bytes(3, 64, _lib.LV_F32) == 16 * 64 * 9 * 4
    assert lib.lv_launch_count() == 0


def test_ctypes_struct_layout_matches_header_sizes():
    import ctypes as C
    # 12 x 4-byte scalars, then 18 pointers (4 sources + 14 others)
    assert C.sizeof(_lib.ConvArgs) == 12 * 4 + 18 * 8
    assert C.sizeof(_lib.WgradItem) == 8 * 4 + 4 * 8 + 8
    assert C.sizeof(_lib.PackItem) == 2 * 8 + 8 * 4


@pytest.mark.parametrize('modname,v2', [('LarvaNet', False), ('larvanet', False), ('LarvaNetV2', True), ('larvanetv2', True)])
def test_plugin_surface(modname, v2):
    mod = importlib.import_module('models.' + modname)
    m = mod.create_model()
    args, rest = m.parse_args(['--num_modules=3', '--num_blocks=2,1,1', '--lr=1e-3', '--unknown_flag=7'])
    assert rest == ['--unknown_flag=7'] and args.num_modules == 3 and args.lr == 1e-3
    with pytest.raises(ValueError):
        m.prepare(is_training=False, scales=[5])
    with pytest.raises(ValueError):
        m.prepare(is_training=False, scales=[2, 4])
    m.prepare(is_training=True, scales=[4], global_step=7)
    assert m.global_step == 7 and m.get_next_train_scale() == 4 and abs(m.get_lr() - 1e-3) < 1e-12
    sd = m.get_model().state_dict()
    shapes = synth.larva_param_shapes([2, 1, 1], v2=v2)
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    # reference init statistics: kaiming-normal(fan_in) * 0.1, zero bias (models/LarvaNet.py:22-39)
    w = sd['body_0.res_blocks.0.body.0.weight'].numpy()
    assert abs(w.std() - 0.1 * np.sqrt(2.0 / 432)) < 0.1 * 0.1 * np.sqrt(2.0 / 432)
    assert float(sd['body_0.res_blocks.0.body.0.bias'].abs().max()) == 0.0
    for name in ('parse_args', 'prepare', 'save', 'restore', 'get_model', 'get_next_train_scale', 'train_step_larva',
                 'validate_for_train', 'upscale', 'test', 'fwd_runtime', 'get_lr'):
        assert callable(getattr(m, name))
    mod2 = mod.create_model()
    mod2.parse_args(['--num_modules=2', '--num_blocks=1,1,1'])
    with pytest.raises(GeneratorExit):
        mod2.prepare(is_training=False, scales=[4])


def test_checkpoint_round_trip_and_v1_to_v2_warm_start(tmp_path):
    v1 = importlib.import_module('models.LarvaNet').create_model()
    v1.parse_args(['--num_modules=2', '--num_blocks=1,1'])
    v1.prepare(is_training=False, scales=[4])
    v1.total_volume = 3e9
    v1.global_step = 12
    v1.save(str(tmp_path))
    ckpt = os.path.join(str(tmp_path), 'model_step12_vol3G.pth')
    assert os.path.exists(ckpt) and os.path.getsize(ckpt) < 2 * sum(p.numel() * 4 for p in v1.get_model().parameters())
    again = importlib.import_module('models.LarvaNet').create_model()
    again.parse_args(['--num_modules=2', '--num_blocks=1,1'])
    again.prepare(is_training=False, scales=[4])
    again.restore(ckpt)
    for k, v in v1.get_model().state_dict().items():
        assert torch.equal(v, again.get_model().state_dict()[k])
    v2 = importlib.import_module('models.LarvaNetV2').create_model()
    v2.parse_args(['--num_modules=2', '--num_blocks=1,1'])
    v2.prepare(is_training=False, scales=[4])
    tail_before = v2.get_model().state_dict()['tail.merge_conv.weight'].clone()
    v2.restore(ckpt)   # reference models/LarvaNetV2.py:196-206: keys the model lacks are ignored
    sd2 = v2.get_model().state_dict()
    assert torch.equal(sd2['head.feature_extraction.weight'], v1.get_model().state_dict()['head.feature_extraction.weight'])
    assert torch.equal(sd2['tail.merge_conv.weight'], tail_before)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_compute_without_gpu_fails_loudly():
    m = importlib.import_module('models.LarvaNet').create_model()
    m.parse_args(['--num_modules=1', '--num_blocks=1'])
    m.prepare(is_training=False, scales=[4])
    with pytest.raises(_lib.LarvaNetB200Error):
        m.upscale([np.zeros((3, 8, 8), np.float32)], 4)
    with pytest.raises(_lib.LarvaNetB200Error):
        m.get_model().head(torch.zeros(1, 3, 8, 8))
    from larvanet_b200 import ops
    with pytest.raises(_lib.LarvaNetB200Error):
        ops.bicubic_x4(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 16, 16))


def test_synthetic_loader_contract():
    ld = importlib.import_module('dataloaders.synthetic_loader').create_loader()
    _, rest = ld.parse_args(['--synthetic_images=2', '--synthetic_height=40', '--synthetic_width=56', '--x=1'])
    assert rest == ['--x=1']
    ld.prepare(scales=[4])
    assert ld.get_num_images() == 2 and ld.is_threaded is False
    inp, tru = ld.get_patch_batch(batch_size=3, scale=4, input_patch_size=16)
    assert len(inp) == 3 and inp[0].shape == (3, 16, 16) and tru[0].shape == (3, 64, 64)
    lr, hr, name = ld.get_image_pair(image_index=1, scale=4)
    assert lr.shape == (3, 40, 56) and hr.shape == (3, 160, 224) and isinstance(name, str)
    assert 0 <= lr.min() and hr.max() <= 255


def test_psnr_helpers_match_oracle():
    import validate
    from oracle import larva_oracle as O
    rs = np.random.RandomState(0)
    a = rs.uniform(-5, 260, (3, 9, 7))
    b = rs.uniform(0, 255, (3, 12, 9))
    u8 = validate._image_to_uint8(a)
    np.testing.assert_array_equal(u8, O.image_to_uint8(a))
    t = validate._fit_truth_image_size(output_image=u8, truth_image=validate._image_to_uint8(b))
    assert t.shape == u8.shape
    assert abs(validate._image_psnr(u8, t) - O.image_psnr(u8, t)) < 1e-9


def test_chop_forward_split_and_combine_are_inverse_for_identity_model():
    from utils import image_utils

    class Nearest:
        def upscale(self, input_list, scale):
            x = np.asarray(input_list)
            return np.repeat(np.repeat(x, scale, axis=2), scale, axis=3)

    img = np.random.RandomState(1).uniform(0, 255, (3, 22, 30))
    out = image_utils.upscale_with_chop_forward(Nearest(), img, scale=4, overlap_size=10)
    np.testing.assert_array_equal(out, np.repeat(np.repeat(img, 4, axis=1), 4, axis=2))


def test_edsr_plugin_surface():
    m = importlib.import_module('models.edsr').create_model()
    args, rest = m.parse_args(['--edsr_res_blocks=3', '--edsr_conv_features=32', '--zzz'])
    assert rest == ['--zzz'] and args.edsr_res_blocks == 3
    m.prepare(is_training=False, scales=[4])
    sd = m.get_model().state_dict()
    shapes = synth.edsr_param_shapes(32, 3, 4)
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == shapes[k] for k in shapes)
    assert not m.get_model().mean_shift.weight.requires_grad       # frozen like the reference (models/edsr.py:135-136)
    with pytest.raises(NotImplementedError):
        m.train_step([], 4, [])


def test_workspace_size_queries_need_no_gpu():
    """Pure host arithmetic of the C-ABI: chain flag workspace and the stream-K style weight-gradient schedule."""
    lib = _lib.load()
    # one workspace serves both chain kernels: max(16x8 tile flags, row-job flags at 1 row per job) + exit counter
    def ws(n, h, w):
        tiles = n * -(-h // 16) * -(-w // 8)
        strips = max(1, -(-(n * (w + 1) - 1) // 128))
        return (max(tiles, strips * max(h, 1)) + 1) * 4
    assert lib.lv_conv_chain_workspace_bytes(16, 48, 48) == ws(16, 48, 48) == (7 * 48 + 1) * 4
    assert lib.lv_conv_chain_workspace_bytes(1, 180, 320) == ws(1, 180, 320) == (3 * 180 + 1) * 4
    assert lib.lv_conv_chain_workspace_bytes(8, 270, 480) == ws(8, 270, 480)
    assert lib.lv_conv_chain_workspace_bytes(0, 48, 48) == ws(0, 48, 48)
    slot = 448 * 128 * 4                      # accumulator columns x TMEM lanes x fp32

    def items(shapes):
        arr = (_lib.WgradItem * len(shapes))()
        for it, (n, h, w) in zip(arr, shapes):
            it.n, it.h, it.w, it.cin, it.cout, it.cin_total, it.cin_off, it.dtype = n, h, w, 48, 48, 48, 0, _lib.LV_BF16
            it.x, it.dy, it.dw = 0x1000, 0x2000, 0x3000
        return arr
    # 40 layers x 288 tiles on <= 148 CTAs: every CTA's range touches at most two layers -> two slots per CTA
    b = lib.lv_wgrad_workspace_bytes(items([(16, 48, 48)] * 40), 40, 4)
    assert b == 148 * 2 * slot
    # fewer jobs than SMs: one CTA per tile, one slot each
    assert lib.lv_wgrad_workspace_bytes(items([(1, 16, 8)] * 3), 3, 50) == 3 * slot
    # the requested splits bound the grid; layers without tiles get no slots
    assert lib.lv_wgrad_workspace_bytes(items([(2, 32, 16), (0, 32, 16)]), 2, 2) == 4 * slot
    # fp32 items take the CUDA-core path: no workspace
    f32 = items([(1, 16, 8)])
    f32[0].dtype = _lib.LV_F32
    assert lib.lv_wgrad_workspace_bytes(f32, 1, 4) == 0


def test_device_scalar_reads_lazily():
    import torch
    from larvanet_b200.engine import DeviceScalar
    acc = torch.tensor([12.0], dtype=torch.float64)
    s = DeviceScalar(acc, 0.25)
    assert s.item() == 3.0 and float(s) == 3.0
    acc[0] = 20.0                                  # read at call time, not at construction
    assert s.item() == 5.0 and float(s.tensor()) == 5.0


# defaults of every wrapper flag, transcribed from the reference (models/LarvaNet.py:47-66, models/LarvaNetV2.py:47-66,
# models/LarvaLeg.py:47-64, models/LarvaLegV2.py:47-67); when the reference modules are importable (oracle/_ref or
# /root/reference) the table itself is checked against them below
REFERENCE_DEFAULTS = {
    'LarvaNet': dict(num_modules=2, num_blocks=16, interpolate='bicubic', val_volume=30e9, lr=4e-4, lr_decay=0.5,
                     lr_step=20000, threshold=0.001, min_lr=1e-8, patience=3, cooldown=6),
    'LarvaNetV2': dict(num_modules=2, num_blocks=16, interpolate='bicubic', val_volume=3e9, lr=1e-4, lr_decay=0.5,
                       threshold=0.001, min_lr=1e-7, patience=3),
    'LarvaLeg': dict(num_modules=2, num_blocks=16, leg=4, interpolate='bicubic', val_volume=3e9, lr=1e-4, lr_decay=0.5,
                     threshold=0.001, min_lr=1e-7, patience=3),
    'LarvaLegV2': dict(num_modules=2, num_blocks=16, leg=4, interpolate='bicubic', val_volume=3e9, lr=1e-4,
                       lr_decay=0.5, threshold=0.001, min_lr=1e-7, patience=3),
}


@pytest.mark.parametrize('modname', sorted(REFERENCE_DEFAULTS))
def test_plugin_flag_defaults_match_reference(modname):
    m = importlib.import_module('models.' + modname).create_model()
    args, _ = m.parse_args([])
    ours = {k: v for k, v in vars(args).items() if k != 'precision'}     # --precision is this repo's extension
    assert ours == REFERENCE_DEFAULTS[modname]
    from oracle import ref_loader
    mods = ref_loader.load()
    if mods is not None:                                                 # the table above == the reference's own parser
        ref_args, _ = mods[modname].create_model().parse_args([])
        assert vars(ref_args) == REFERENCE_DEFAULTS[modname]
    # scheduler: cooldown only where the reference has the flag (models/LarvaNet.py:90-92 vs models/LarvaNetV2.py:86-88)
    m.parse_args(['--num_modules=1', '--num_blocks=1'] + (['--leg=1'] if 'Leg' in modname else []))
    m.prepare(is_training=True, scales=[4])
    assert m.scheduler.cooldown == (6 if modname == 'LarvaNet' else 0)
    assert m.scheduler.patience == 3 and m.scheduler.factor == 0.5 and m.scheduler.mode == 'max'
    assert m.scheduler.min_lrs == [REFERENCE_DEFAULTS[modname]['min_lr']]


@pytest.mark.parametrize('modname,v2', [('LarvaLeg', False), ('LarvaLegV2', True)])
def test_leg_plugins_share_the_state_dict_and_validate_leg(modname, v2):
    m = importlib.import_module('models.' + modname).create_model()
    args, rest = m.parse_args(['--num_modules=3', '--num_blocks=1,1,1', '--leg=2', '--x'])
    assert rest == ['--x'] and args.leg == 2
    m.prepare(is_training=False, scales=[4])
    assert m.get_model().leg == 2
    assert list(m.get_model().state_dict().keys()) == list(synth.larva_param_shapes([1, 1, 1], v2=v2).keys())
    bad = importlib.import_module('models.' + modname).create_model()
    bad.parse_args(['--num_modules=2', '--num_blocks=1,1', '--leg=3'])
    with pytest.raises(ValueError):
        bad.prepare(is_training=False, scales=[4])


@pytest.mark.parametrize('defer,call_hook', [(False, False), (True, True), (True, False)])
def test_prefetcher_control_flow_without_a_gpu(monkeypatch, defer, call_hook):
    """DevicePrefetcher's slot / refill book-keeping with the CUDA stream and event objects replaced by recorders (the
    copies themselves are plain CPU copies here): every batch comes out once, in order, nothing is dropped at the end,
    a slot is only refilled after the consumer's "done with it" event was recorded, and with defer=True the refill
    happens in run_deferred() -- or at the next __next__ when nobody calls it."""
    import contextlib
    from larvanet_b200 import prefetch

    log = []

    class FakeEvent:
        def record(self, stream=None):
            log.append(('record', id(self)))

    class FakeStream:
        def __init__(self, device=None):
            pass

        def wait_event(self, ev):
            log.append(('wait', id(ev)))

    monkeypatch.setattr(torch.cuda, 'Stream', FakeStream)
    monkeypatch.setattr(torch.cuda, 'Event', FakeEvent)
    monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda device=None: FakeStream())
    host = [(torch.full((2, 3), float(i)), torch.full((2, 5), float(-i))) for i in range(7)]
    issued = []
    real_issue = prefetch.DevicePrefetcher._issue

    def counting_issue(self):
        issued.append(len(got))
        real_issue(self)

    monkeypatch.setattr(prefetch.DevicePrefetcher, '_issue', counting_issue)
    got = []
    feeder = prefetch.DevicePrefetcher(iter(host), 'cpu', depth=2, defer=defer)
    assert issued == [0, 0]                      # two batches in flight before the first hand-out
    for x, t in feeder:
        got.append((x.clone(), t.clone()))
        if defer and call_hook:
            before = len(issued)
            prefetch.run_deferred()
            assert len(issued) - before == (1 if len(got) >= 2 else 0)    # the refill owed by this hand-out, now
            prefetch.run_deferred()
            assert len(issued) - before == (1 if len(got) >= 2 else 0)    # and only once
    assert len(got) == len(host)
    for (x, t), (hx, ht) in zip(got, host):
        assert torch.equal(x, hx) and torch.equal(t, ht)
    # a refill of a slot is always preceded by the consumer's event for that slot having been recorded
    recorded = set()
    for kind, ev in log:
        if kind == 'record':
            recorded.add(ev)
        else:
            assert ev in recorded


def test_band_ranges_cover_the_frame_with_the_receptive_halo():
    """tiling.band_ranges / receptive_halo (exact band-sharded inference): bands partition the rows, read ranges add the
    halo clipped at the frame, the halo equals the number of 3x3 conv rings on the longest path + the bicubic margin."""
    from larvanet_b200 import tiling
    assert tiling.receptive_halo([4, 4, 4, 4]) == 1 + 2 * 16 + 2 + 2           # head + 32 body convs + 2 leg convs + bicubic
    assert tiling.receptive_halo([4, 4, 4, 4], v2=True) == 1 + 32 + 3 + 2        # V2 tail: merge conv + 2
    assert tiling.receptive_halo([4, 4, 4, 4], exit_leg=1) == 1 + 8 + 2 + 2
    assert tiling.receptive_halo([4, 4], exit_leg=0) == 2                        # bicubic base only
    for height, bands, halo in [(270, 4, 37), (10, 3, 37), (7, 16, 2), (1, 1, 5)]:
        r = tiling.band_ranges(height, bands, halo)
        assert len(r) == min(bands, height)
        assert r[0][0] == 0 and r[-1][1] == height
        for (y0, y1, lo, hi), nxt in zip(r, r[1:] + [None]):
            assert y0 < y1 and lo == max(0, y0 - halo) and hi == min(height, y1 + halo)
            if nxt is not None:
                assert nxt[0] == y1
    assert tiling.bands_for_rank(7, 1, 3) == [1, 4]


def test_shape_cache_is_lru_bounded():
    """engine._ShapeCache (unbounded per-shape buffers + graphs): at most `capacity` entries, least
    recently used evicted first, hit counts drive the capture-on-second-use policy."""
    from larvanet_b200.engine import _ShapeCache
    built = []
    c = _ShapeCache(2)
    mk = lambda k: (lambda: built.append(k) or object())
    e1 = c.get('a', mk('a'))
    assert e1['hits'] == 1 and e1['graph'] is None
    assert c.get('a', mk('a')) is e1 and e1['hits'] == 2 and built == ['a']
    c.get('b', mk('b'))
    c.get('a', mk('a'))            # 'a' is now the most recent
    c.get('c', mk('c'))            # evicts 'b'
    assert len(c) == 2 and built == ['a', 'b', 'c']
    c.get('b', mk('b'))            # rebuilt: it was evicted
    assert built == ['a', 'b', 'c', 'b'] and len(c) == 2
    c.clear()
    assert len(c) == 0


def test_fused_adamw_is_a_plain_adamw_to_torch_without_the_profiler_wrapper():
    """FusedAdamW keeps the torch.optim.AdamW surface (param_groups, schedulers, state_dict) and declines the per-class
    profiler wrapper torch puts around Optimizer.step (host time per step); without an attached engine step() fails
    loudly -- there is no unfused fallback."""
    from larvanet_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(3))
    opt = FusedAdamW([p], lr=4e-4)
    assert isinstance(opt, torch.optim.AdamW) and opt.param_groups[0]['lr'] == 4e-4
    assert getattr(FusedAdamW.step, 'hooked', False) and not hasattr(FusedAdamW.step, '__wrapped__')
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='max', factor=0.5, patience=0, threshold=0.0)
    sched.step(1.0)
    sched.step(0.5)
    sched.step(0.4)
    assert opt.param_groups[0]['lr'] < 4e-4                 # the scheduler drives our param group
    with pytest.raises(_lib.LarvaNetB200Error):
        opt.step()
    assert 'param_groups' in opt.state_dict()
