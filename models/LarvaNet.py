"""LarvaNet model plugin -- drop-in for reference models/LarvaNet.py with the arithmetic on B200 kernels.

Same plugin surface (`create_model()`, `LarvaNet.parse_args/prepare/train_step_larva/upscale/test/fwd_runtime/
save/restore/...`), same flags, same sub-module names and therefore the same `state_dict()` keys
(`head.feature_extraction`, `body_{i}.res_blocks.{j}.body.{0,2}`, `body_{i}.leg.recon_block.{0,2}`), so reference
checkpoints load unchanged.  What differs is below that surface: parameters live in one flat fp32 arena and every
forward/backward runs through `larvanet_b200.engine.LarvaEngine` (hand-written sm_100a CUDA behind a C-ABI).
There is no PyTorch/CPU fallback: computing without a B200 raises `LarvaNetB200Error`.
"""
import argparse
import copy
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from models.base import BaseModel
from larvanet_b200.prefetch import run_deferred as _run_deferred_refills

NUM_FILTERS = 48  # reference models/LarvaNet.py:226,239,254


def create_model():
    return LarvaNet()


def initialize_weights(net_l, scale=1):
    """Kaiming-normal (fan_in) scaled by `scale`, zero bias -- reference models/LarvaNet.py:22-39."""
    nets = net_l if isinstance(net_l, list) else [net_l]
    for net in nets:
        for m in net.modules():
            if isinstance(m, Conv3x3):
                fan_in = m.weight.shape[1] * 9
                with torch.no_grad():
                    m.weight.normal_(0.0, math.sqrt(2.0 / fan_in))
                    m.weight.mul_(scale)
                    m.bias.zero_()


def _engine_of(mod):
    root = getattr(mod, '_lv_root', None)
    root = root() if callable(root) else None
    if root is None:
        from larvanet_b200._lib import LarvaNetB200Error
        raise LarvaNetB200Error(
            f'{type(mod).__name__} can only compute as part of a LarvaNetModule (its kernels use the network\'s packed '
            'weight arena); there is no stand-alone PyTorch path')
    return root.engine()


class Conv3x3(nn.Module):
    """Parameter holder with nn.Conv2d's names/shapes (weight [O,I,3,3], bias [O]).  It owns no arithmetic: the
    enclosing block runs the fused kernels."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, 3, 3))
        self.bias = nn.Parameter(torch.empty(out_channels))
        bound = 1.0 / math.sqrt(in_channels * 9)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            self.bias.uniform_(-bound, bound)

    def extra_repr(self):
        return f'{self.in_channels}, {self.out_channels}, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1)'

    def forward(self, x):
        from larvanet_b200._lib import LarvaNetB200Error
        raise LarvaNetB200Error('Conv3x3 is a parameter holder; call the enclosing ResidualBlock/LarvaLeg/LarvaHead')


class ResidualBlock(nn.Module):
    """x + conv(relu(conv(x))) -- reference models/LarvaNet.py:205-220."""

    def __init__(self, num_channels):
        super().__init__()
        self.body = nn.Sequential(Conv3x3(num_channels, num_channels), nn.ReLU(inplace=True),
                                  Conv3x3(num_channels, num_channels))
        initialize_weights(self.body, 0.1)

    def forward(self, x):
        i, j = self._lv_index
        return _engine_of(self).run_resblock(i, j, x)


class LarvaHead(nn.Module):
    """Conv 3->48 -- reference models/LarvaNet.py:223-233."""

    def __init__(self):
        super().__init__()
        self.feature_extraction = Conv3x3(3, NUM_FILTERS)
        initialize_weights(self.feature_extraction, 0.1)

    def forward(self, x):
        return _engine_of(self).run_head(x)


class LarvaLeg(nn.Module):
    """PixelShuffle(4)(conv(relu(conv(fea)))) + base -- reference models/LarvaNet.py:251-267."""

    def __init__(self):
        super().__init__()
        self.recon_block = nn.Sequential(Conv3x3(NUM_FILTERS, NUM_FILTERS), nn.ReLU(inplace=True),
                                         Conv3x3(NUM_FILTERS, NUM_FILTERS))
        initialize_weights(self.recon_block, 0.1)
        self.upsample = nn.PixelShuffle(4)

    def forward(self, fea, base):
        return _engine_of(self).run_leg(self._lv_index, fea, base)


class LarvaBody(nn.Module):
    """x + res_blocks(x), plus its own early-exit leg -- reference models/LarvaNet.py:236-248."""

    def __init__(self, num_blocks):
        super().__init__()
        self.res_blocks = nn.Sequential(*[ResidualBlock(NUM_FILTERS) for _ in range(num_blocks)])
        self.leg = LarvaLeg()

    def forward(self, x):
        return _engine_of(self).run_body(self._lv_index, x)


class LarvaNetModule(nn.Module):
    """head -> body_0..body_{M-1} -> last leg (+ bicubic base) -- reference models/LarvaNet.py:270-293."""

    V2 = False
    EARLY_EXIT = False

    def __init__(self, args):
        super().__init__()
        self.len = args.num_modules
        self.interpolate = args.interpolate
        self.head = LarvaHead()
        blocks = [int(v) for v in str(args.num_blocks).split(',')]
        if len(blocks) != self.len:
            raise GeneratorExit('Argument num_blocks should have the same number of elements as num_modules.')
        self.blocks = blocks
        for i, nb in enumerate(blocks):
            setattr(self, f'body_{i}', LarvaBody(num_blocks=nb))
        self._build_extra()
        self._wire()
        self._engine = None
        self.precision = getattr(args, 'precision', 'bf16')
        # early-exit plugins only (reference models/LarvaLeg.py:275, models/LarvaLegV2.py:342): None = full network
        self.leg = getattr(args, 'leg', None) if self.EARLY_EXIT else None
        if self.leg is not None and not 0 <= self.leg <= self.len:
            raise ValueError(f'--leg={self.leg} is outside 0..num_modules={self.len}')

    def _build_extra(self):
        pass

    def _wire(self):
        import weakref
        ref = weakref.ref(self)
        for m in self.modules():
            object.__setattr__(m, '_lv_root', ref)
        for i in range(self.len):
            body = getattr(self, f'body_{i}')
            object.__setattr__(body, '_lv_index', i)
            object.__setattr__(body.leg, '_lv_index', i)
            for j, rb in enumerate(body.res_blocks):
                object.__setattr__(rb, '_lv_index', (i, j))

    def engine(self):
        """The LarvaEngine bound to this module (created on first use; needs a B200)."""
        if self._engine is None:
            from larvanet_b200._lib import LarvaNetB200Error
            from larvanet_b200.engine import LarvaEngine
            if self.interpolate != 'bicubic':
                raise LarvaNetB200Error(f'--interpolate={self.interpolate}: only bicubic has a kernel')
            if not torch.cuda.is_available():
                raise LarvaNetB200Error('LarvaNet on larvanet_b200 needs a CUDA device (sm_100a); there is no CPU path')
            dt = {'bf16': torch.bfloat16, 'fp32': torch.float32}[self.precision]
            self._engine = LarvaEngine(self, self.blocks, v2=self.V2, act_dtype=dt)
        return self._engine

    def _apply(self, fn, *a, **k):
        # .to()/.cuda()/.float() must not re-allocate parameters once they are views of the engine arena
        if self._engine is not None:
            probe = fn(torch.empty(0, device=self._engine.device))
            if probe.device == self._engine.device and probe.dtype == torch.float32:
                return self
            from larvanet_b200._lib import LarvaNetB200Error
            raise LarvaNetB200Error('cannot move/cast a LarvaNetModule after its engine has been created')
        return super()._apply(fn, *a, **k)

    def base(self, x):
        return self.engine().run_base(x)

    def forward(self, x):
        # fused path; result is a fresh tensor like the reference's (the engine's buffer is reused per call)
        # (early-exit plugins: leg == 0 returns the bicubic base, else bodies 0..leg-1 and that body's leg --
        # reference models/LarvaLeg.py:289-299, models/LarvaLegV2.py:357-368)
        return self.engine().forward(x, exit_leg=self.leg).clone()


class LarvaNet(BaseModel):
    MODULE = LarvaNetModule

    def __init__(self):
        super().__init__()
        self.volume_per_step = 0

    # Flag defaults that differ between the reference's plugins (models/LarvaNet.py:47-66, models/LarvaNetV2.py:47-66,
    # models/LarvaLeg.py:47-64, models/LarvaLegV2.py:47-67): subclasses override these class attributes.
    DEFAULTS = dict(val_volume=30e9, lr=4e-4, min_lr=1e-8)
    HAS_COOLDOWN = True      # --lr_step / --cooldown exist only in models/LarvaNet.py (:58,:63); elsewhere cooldown = 0
    HAS_LEG = False          # --leg exists only in the LarvaLeg plugins (models/LarvaLeg.py:52)

    def parse_args(self, args):
        parser = argparse.ArgumentParser()
        d = self.DEFAULTS
        parser.add_argument('--num_modules', type=int, default=2, help='Number of bodies (early exits).')
        parser.add_argument('--num_blocks', type=str, default=16, help='Residual blocks per body, comma separated.')
        if self.HAS_LEG:
            parser.add_argument('--leg', type=int, default=4, help='The early exit leg number, starts at 1.')
        parser.add_argument('--interpolate', type=str, default='bicubic', help='Interpolation of the base image.')
        parser.add_argument('--val_volume', type=float, default=d['val_volume'], help='Training volume between validations.')
        parser.add_argument('--lr', type=float, default=d['lr'], help='Initial learning rate.')
        parser.add_argument('--lr_decay', type=float, default=0.5, help='Learning rate decay factor.')
        if self.HAS_COOLDOWN:
            parser.add_argument('--lr_step', type=int, default=20000, help='Learning rate decay step.')
        parser.add_argument('--threshold', type=float, default=0.001, help='Plateau threshold (absolute, dB).')
        parser.add_argument('--min_lr', type=float, default=d['min_lr'], help='Minimum learning rate.')
        parser.add_argument('--patience', type=int, default=3, help='Plateau patience.')
        if self.HAS_COOLDOWN:
            parser.add_argument('--cooldown', type=int, default=6, help='Plateau cooldown.')
        # larvanet_b200 extension (absent in the reference)
        parser.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'fp32'],
                            help='bf16: tcgen05 tensor-core path; fp32: CUDA-core validation mode.')
        self.args, remaining_args = parser.parse_known_args(args=args)
        return copy.deepcopy(self.args), remaining_args

    def prepare(self, is_training, scales, global_step=0):
        self.global_step = global_step
        self.total_volume = 0.0
        self.temp_volume = 0
        self.scale_list = scales
        for scale in self.scale_list:
            if scale not in (2, 3, 4):
                raise ValueError('Unsupported scale is provided.')
        if len(self.scale_list) != 1:
            raise ValueError('Only one scale should be provided.')
        self.scale = self.scale_list[0]

        self.model = self.MODULE(args=self.args)
        self.is_training = is_training
        if is_training:
            from larvanet_b200.optim import FusedAdamW
            self.loss_fn = nn.L1Loss()  # kept for API parity; the loss itself is fused into the exit-conv epilogue
            self.optim = FusedAdamW([p for p in self.model.parameters() if p.requires_grad], lr=self.args.lr)
            self.scheduler = optim.lr_scheduler.ReduceLROnPlateau(
                self.optim, mode='max', factor=self.args.lr_decay, patience=self.args.patience,
                cooldown=getattr(self.args, 'cooldown', 0), threshold=self.args.threshold, threshold_mode='abs', min_lr=self.args.min_lr)
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.model = self.model.to(self.device)

    def _engine(self):
        eng = self.model.engine()
        if getattr(self, 'is_training', False) and self.optim._engine is None:
            self.optim.attach(eng)
        return eng

    # reference models/LarvaNet.py:98-139
    def train_step_larva(self, args, val_dataloader, input_tensor, truth_tensor, summary=None):
        self.global_step += 1
        self.temp_volume += self.volume_per_step
        eng = self._engine()
        # forward of every exit + L1 losses + full backward in one fused pass; gradients land in param.grad
        loss = eng.train_step(input_tensor, truth_tensor)
        self.optim.step()
        _run_deferred_refills()     # host-side input prefetch work, now that the GPU is busy (larvanet_b200/prefetch.py)

        if self.global_step == 1:
            self.validate_for_train(args, val_dataloader)
        if self.temp_volume >= self.args.val_volume:
            self.total_volume += self.temp_volume
            self.temp_volume = 0
            self.validate_for_train(args, val_dataloader)
            if getattr(self, 'dp_rank', 0) == 0:     # data parallel: every rank holds the same weights, one writes
                self.save(base_path=args.train_path)
                print(f'saved a model checkpoint at volume {self.total_volume/1e9:.0f}G')
            if summary is not None:
                summary.add_scalar('loss', float(loss), self.global_step)
                summary.add_scalar('lr', self.get_lr(), self.global_step)
                out = self.model(input_tensor)
                for name, t in (('input', input_tensor), ('output', out), ('truth', truth_tensor)):
                    u8 = t.clamp(0, 255).byte()
                    for i in range(min(4, len(u8))):
                        summary.add_image(f'{name}/{i}', u8[i], self.global_step)
        return loss.item()

    # reference models/LarvaNet.py:141-161
    def validate_for_train(self, args, dataloader):
        import validate
        print('begin validation')
        # Scored on the device (lv_psnr_sqsum == _image_to_uint8 + _fit_truth_image_size + _image_psnr): the truth
        # images are uploaded once and kept, each validation reads back 8 bytes per image instead of the HR frame.
        from larvanet_b200 import ops
        if not hasattr(self, '_val_truth'):
            self._val_truth = {}
        sq_sums, numels = [], []
        for image_index in range(dataloader.get_num_images()):
            input_image, truth_image, _ = dataloader.get_image_pair(image_index=image_index, scale=4)
            out = self.model(self._as_input([input_image]))[0].contiguous()
            key = (id(dataloader), image_index)
            if key not in self._val_truth:
                self._val_truth[key] = torch.as_tensor(np.asarray(truth_image), dtype=torch.float32,
                                                       device=self.device).contiguous()
            sq = torch.zeros(1, dtype=torch.float64, device=self.device)
            ops.psnr_sqsum(out, self._val_truth[key], sq)
            sq_sums.append(sq)
            numels.append(out.numel())
        mses = torch.cat(sq_sums).cpu().numpy() / np.asarray(numels, dtype=np.float64)
        psnr_list = 10.0 * np.log10(255.0 ** 2 / mses)
        average_psnr = np.mean(psnr_list)
        print(f'step {self.global_step}, volume {self.total_volume/1e9:.0f}G,'
              f' psnr={average_psnr:.8f}, lr = {self.get_lr():.8f}')
        self.scheduler.step(average_psnr)

    def _as_input(self, input_list):
        if torch.is_tensor(input_list):
            return input_list.to(device=self.device, dtype=torch.float32)
        return torch.as_tensor(np.asarray(input_list), dtype=torch.float32, device=self.device)

    def upscale(self, input_list, scale):
        return self.model(self._as_input(input_list)).detach().cpu().numpy()

    def upscale_uint8(self, input_list, scale):
        """`validate._image_to_uint8(self.upscale(...))` with the round/clip done on the device: the device->host copy
        is 1 byte per sample instead of 4 (what get_sr.py / validate.py need before they write or score a PNG).  The uint8
        frame comes straight out of the exit conv's PixelShuffle epilogue (`lv_conv_args.out_u8`): no fp32 frame is written."""
        m = self.model
        return m.engine().forward(self._as_input(input_list), exit_leg=m.leg, uint8=True).cpu().numpy()

    def test(self, input_list):
        return self.model(self._as_input(input_list))

    def fwd_runtime(self, input_tensor):
        return self.model(input_tensor)

    def save(self, base_path):
        save_path = os.path.join(base_path, 'model_step%d_vol%.0fG.pth' % (self.global_step, self.total_volume / 1e9))
        torch.save({k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}, save_path)

    def restore(self, ckpt_path, target=None):
        self.model.load_state_dict(torch.load(ckpt_path, map_location=self.device))

    def get_model(self):
        return self.model

    def get_next_train_scale(self):
        return self.scale_list[np.random.randint(len(self.scale_list))]

    def get_lr(self):
        return self.optim.param_groups[0]['lr']
