"""ctypes binding of libfaststyle_b200.so (the C-ABI in include/faststyle_b200.h).

There is deliberately no fallback: if the shared library is missing or cannot be
loaded, importing any compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfaststyle_b200.so")

FS_V_NCONV = 10
FS_T_NCONV = 16
ENG_TRANSFORM, ENG_TRANSFORM_BWD, ENG_VGG, ENG_VGG_BWD, ENG_DECONV = 1, 2, 4, 8, 16


class FsError(RuntimeError):
    pass


class LossConfigC(C.Structure):
    _fields_ = [("n_content", C.c_int), ("content_layer", C.c_int * FS_V_NCONV),
                ("content_w", C.c_float * FS_V_NCONV),
                ("n_style", C.c_int), ("style_layer", C.c_int * FS_V_NCONV),
                ("style_w", C.c_float * FS_V_NCONV), ("beta", C.c_float)]


_P, _I, _LL, _F, _U, _SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint, C.c_size_t
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every int-returning function is status-checked
SIGNATURES = {
    "fs_last_error": (C.c_char_p, []),
    "fs_version": (_I, []),
    "fs_launch_count": (_LL, []),
    "fs_transform_param_count": (_LL, []),
    "fs_transform_param_slot": (_I, [_I, _I, C.POINTER(_LL), C.POINTER(_LL)]),
    "fs_vgg_flat_floats": (_LL, []),
    "fs_vgg_packed_floats": (_LL, []),
    "fs_vgg_pack": (_I, [_P, _P, _P]),
    "fs_engine_create": (_I, [_I, _I, _I, _I, _U, _U, _PP]),
    "fs_engine_destroy": (_I, [_P]),
    "fs_engine_set_tensor_path": (_I, [_P, _I]),
    "fs_set_tc_pair": (_I, [_I]),
    "fs_set_tc_epilogue_warps": (_I, [_I]),
    "fs_engine_keep_activations": (_I, [_P, _I]),
    "fs_engine_set_frozen_weights": (_I, [_P, _I]),
    "fs_engine_profile": (_I, [_P, _I]),
    "fs_engine_profile_read": (_I, [_P, _I, _P, _P, _P]),
    "fs_engine_profile_bytes": (_I, [_P, _I, _P]),
    "fs_engine_profile_records": (_I, [_P, _I, _P, _P, _P, _P]),
    "fs_engine_workspace_bytes": (_SZ, [_P]),
    "fs_engine_bind": (_I, [_P, _P, _SZ]),
    "fs_engine_output_dims": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "fs_engine_vgg_activation": (_I, [_P, _I, _PP, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "fs_engine_transform_activation": (_I, [_P, _I, _I, _PP, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "fs_transform_forward": (_I, [_P, _P, _P, _P, _P]),
    "fs_vgg_forward": (_I, [_P, _P, _P, _I, _P]),
    "fs_vgg_grams": (_I, [_P, _P, _P, _I, C.POINTER(_I), _PP, _P]),
    "fs_vgg_set_content_targets": (_I, [_P, _P, _P, C.POINTER(LossConfigC), _P]),
    "fs_perceptual_loss": (_I, [_P, _P, _P, C.POINTER(LossConfigC), _PP, _P, _P, _P]),
    "fs_train_fwd_bwd": (_I, [_P, _P, _P, _P, C.POINTER(LossConfigC), _PP, _P, _P, _P, _P]),
    "fs_adam_step": (_I, [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _P, _P]),
    "fs_enable_peer_access": (_I, [_I]),
    "fs_peer_buffer_create": (_I, [_SZ, _PP, _P]),
    "fs_peer_buffer_open": (_I, [_P, _PP]),
    "fs_peer_buffer_close": (_I, [_P]),
    "fs_peer_buffer_free": (_I, [_P]),
    "fs_dp_allreduce_adam": (_I, [_PP, _I, _I, _LL, _LL, _LL, _I, _LL, _U, _P, _P, _P, _F, _F, _F, _F, _P, _P, _P, _P]),
    "fs_conv2d_forward": (_I, [_P, _P, _P, _P] + [_I] * 10 + [_P]),
    "fs_conv2d_dgrad": (_I, [_P, _P, _P, _P] + [_I] * 9 + [_P]),
    "fs_conv2d_wgrad_scratch_floats": (_LL, [_I, _I, _I, _I]),
    "fs_conv2d_wgrad": (_I, [_P, _P, _P, _P, _LL] + [_I] * 9 + [_P]),
    "fs_upconv2d_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "fs_instnorm_forward": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "fs_maxpool2x2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "fs_gram_scratch_floats": (_LL, [_I, _I]),
    "fs_gram_forward": (_I, [_P, _P, _P, _LL, _I, _I, _I, _I, _P]),
    "fs_loss_sqdiff": (_I, [_P, _P, _LL, C.c_double, _P, _P, _P]),
    "fs_loss_style": (_I, [_P, _P, _I, _I, C.c_double, _P, _P, _P]),
    "fs_loss_tv": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "fs_resize_bicubic_tf1_u8": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "fs_frame_u8_to_f32": (_I, [_P, _P, _LL, _P]),
    "fs_frame_f32_to_u8": (_I, [_P, _P, _LL, _I, _P]),
    "fs_conv3x3_tc_scratch_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "fs_conv3x3_tc_forward": (_I, [_P, _P, _P, _P, _P, _SZ] + [_I] * 7 + [_P]),
    "fs_conv3x3_tc_dgrad": (_I, [_P, _P, _P, _P, _SZ] + [_I] * 6 + [_P]),
    "fs_wgrad3x3_tc_scratch_bytes": (_SZ, [_I, _I, _I]),
    "fs_gram_tc_scratch_bytes": (_SZ, [_I, _I, _I, _I]),
    "fs_gram_tc_forward": (_I, [_P, _P, _P, _SZ, _I, _I, _I, _I, _P]),
    "fs_wgrad3x3_tc": (_I, [_P, _P, _P, _P, _SZ, _I, _I, _I, _I, _P]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library; raises FsError if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FsError(
            "libfaststyle_b200.so is not built (%s). Run `python -m faststyle_b200.build` "
            "(needs nvcc). There is no CPU fallback." % LIB_PATH)
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise FsError("cannot load %s: %s" % (LIB_PATH, e))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError => header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise FsError(load().fs_last_error().decode("utf-8", "replace") or "faststyle_b200 error %d" % status)


def call(name: str, *args):
    """Call a status-returning entry point and raise FsError on failure."""
    check(getattr(load(), name)(*args))
