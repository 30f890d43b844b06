"""ctypes binding of liblarvanet_b200.so (the C-ABI declared in include/larvanet_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a `LarvaNetB200Error` is raised.
`load()` never builds on a machine with a GPU silently either -- `larvanet_b200.build.build()` is invoked only when
the .so is absent and nvcc is present (the build container), otherwise the error tells the user what to run.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LARVANET_B200_LIB') or os.path.join(HERE, 'liblarvanet_b200.so')   # env: A/B builds (developer)

LV_F32, LV_BF16 = 0, 1
LV_EPI_NHWC, LV_EPI_PS4_NCHW, LV_EPI_PS2_NHWC, LV_EPI_RGB_NCHW = 0, 1, 2, 3
LV_MAX_SRC = 4
LV_W_TAP_MAJOR, LV_W_KY_STACKED = 0, 1
LV_CHAIN_MAX_LAYERS = 96
ABI_VERSION = 2


class LarvaNetB200Error(RuntimeError):
    pass


class ConvArgs(C.Structure):
    _fields_ = [
        ('n', C.c_int32), ('h', C.c_int32), ('w', C.c_int32),
        ('cin', C.c_int32), ('num_src', C.c_int32), ('cout', C.c_int32),
        ('dtype', C.c_int32), ('relu', C.c_int32), ('epilogue', C.c_int32), ('wlayout', C.c_int32),
        ('res_scale', C.c_float), ('reserved1', C.c_float),
        ('src', C.c_void_p * LV_MAX_SRC),
        ('weights', C.c_void_p), ('bias', C.c_void_p), ('mask', C.c_void_p),
        ('res1', C.c_void_p), ('res2', C.c_void_p), ('out', C.c_void_p),
        ('out_hr', C.c_void_p), ('base_hr', C.c_void_p), ('truth_hr', C.c_void_p),
        ('loss_sum', C.c_void_p), ('grad_sign', C.c_void_p),
        ('post_w', C.c_void_p), ('post_b', C.c_void_p), ('out_u8', C.c_void_p),
    ]


class WgradItem(C.Structure):
    _fields_ = [
        ('n', C.c_int32), ('h', C.c_int32), ('w', C.c_int32),
        ('cin', C.c_int32), ('cout', C.c_int32), ('cin_total', C.c_int32), ('cin_off', C.c_int32),
        ('dtype', C.c_int32),
        ('x', C.c_void_p), ('dy', C.c_void_p), ('dw', C.c_void_p), ('db', C.c_void_p),
        ('scale', C.c_float), ('overwrite', C.c_int32),
    ]


class FusedConv(C.Structure):
    _fields_ = [('w_off', C.c_int64), ('cout', C.c_int32), ('cin_total', C.c_int32), ('fwd', C.c_void_p),
                ('bwd', C.c_void_p * LV_MAX_SRC)]


class PatchItem(C.Structure):
    _fields_ = [('lr', C.c_void_p), ('hr', C.c_void_p), ('h', C.c_int32), ('w', C.c_int32), ('y', C.c_int32), ('x', C.c_int32),
                ('rot', C.c_int32), ('flip', C.c_int32)]


class PackItem(C.Structure):
    _fields_ = [
        ('w', C.c_void_p), ('packed', C.c_void_p),
        ('O', C.c_int32), ('I', C.c_int32), ('transpose', C.c_int32),
        ('i_off', C.c_int32), ('i_cnt', C.c_int32), ('cin', C.c_int32),
        ('dtype', C.c_int32), ('wlayout', C.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol include/larvanet_b200.h declares
SIGNATURES = {
    'lv_last_error': (C.c_char_p, []),
    'lv_abi_version': (C.c_int, []),
    'lv_device_check': (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    'lv_packed_weight_bytes': (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    'lv_pack_conv3x3_weights': (C.c_int, [C.POINTER(PackItem), C.c_int, C.c_void_p]),
    'lv_conv3x3': (C.c_int, [C.POINTER(ConvArgs), C.c_int, C.c_void_p]),
    'lv_conv3x3_simt': (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    'lv_conv_chain_workspace_bytes': (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    'lv_conv3x3_chain': (C.c_int, [C.POINTER(ConvArgs), C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    'lv_head_bicubic_fwd': (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p]),
    'lv_bicubic_x4': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'lv_head_wgrad_workspace_bytes': (C.c_int64, [C.c_int]),
    'lv_head_wgrad': (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_float, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    'lv_wgrad_workspace_bytes': (C.c_int64, [C.POINTER(WgradItem), C.c_int, C.c_int]),
    'lv_conv3x3_wgrad': (C.c_int, [C.POINTER(WgradItem), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'lv_conv3x3_wgrad_simt': (C.c_int, [C.POINTER(WgradItem), C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'lv_nchw_to_nhwc': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'lv_nhwc_to_nchw': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'lv_image_to_uint8': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'lv_psnr_sqsum': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'lv_l1_loss_grad': (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p]),
    'lv_adamw_pack_step': (C.c_int, [C.c_void_p] * 4 + [C.c_int64] + [C.c_float] * 5 + [C.c_int, C.c_float,
                                                                                      C.POINTER(FusedConv), C.c_int, C.c_void_p]),
    'lv_adamw_step': (C.c_int, [C.c_void_p] * 4 + [C.c_int64] + [C.c_float] * 5 + [C.c_int, C.c_float, C.c_void_p]),
    'lv_dp_adamw_pack_step': (C.c_int, [C.c_void_p] * 3 + [C.c_int64] + [C.c_float] * 5 + [C.c_int, C.c_float,
                                                                                         C.POINTER(FusedConv), C.c_int] +
                              [C.POINTER(C.c_void_p)] * 4 + [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    'lv_crop_augment': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'lv_launch_count': (C.c_int64, []),
}

_lib = None


def load():
    """Load (once) and return the ctypes library.  Raises LarvaNetB200Error when it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise LarvaNetB200Error(
                f'{LIB_PATH} is missing and could not be built ({e}). Run `python -m larvanet_b200.build` '
                'on a machine with the CUDA 12.9 toolkit; there is no CPU or PyTorch fallback.') from e
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise LarvaNetB200Error(f'cannot load {LIB_PATH}: {e}') from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise LarvaNetB200Error(f'{LIB_PATH} does not export {name}; rebuild it') from e
        fn.restype = res
        fn.argtypes = args
    if lib.lv_abi_version() != ABI_VERSION:
        raise LarvaNetB200Error(f'ABI mismatch: library {lib.lv_abi_version()} vs binding {ABI_VERSION}; rebuild')
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().lv_last_error().decode('utf-8', 'replace')
        raise LarvaNetB200Error(f'{what or "larvanet_b200 call"} failed (code {rc}): {msg}')


def launch_count():
    return int(load().lv_launch_count())
