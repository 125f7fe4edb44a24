"""ctypes binding of libdtp_sm100.so (include/dtp.h). Fails loudly when the CUDA library is missing: there is no
CPU or PyTorch fallback behind this module."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdtp_sm100.so")

_lib = None

vp, i32, i64, f32, cp = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_char_p

_OPS = {
    "dtp_op_linear": [vp, i32, i32, vp, i32, i32, i32, vp, i32, i32, vp, vp, i32, vp, i32, i32, f32, i32, i32, i32, vp],
    "dtp_op_conv3x3": [vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, vp, i32, vp, i32, i32, f32, i32, i32, i32, vp],
    "dtp_op_conv3x3_shortcut": [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, vp, i32, i32, vp],
    "dtp_op_conv3x3_s2": [vp, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, i32, vp],
    "dtp_op_upconv2x": [vp, i32, i32, i32, i32, vp, i32, vp, vp, vp, i32, vp],
    "dtp_op_bmm": [vp, i32, i64, i64, vp, i32, i64, i64, i32, i32, i32, i32, i32, i32, vp, i32, i64, i64, f32, i32, i32,
                   vp],
    "dtp_op_groupnorm": [vp, i32, vp, i32, i32, i32, i32, vp, vp, f32, i32, vp, vp],
    "dtp_op_layernorm": [vp, i32, i32, vp, vp, f32, vp, vp],
    "dtp_op_softmax": [vp, i64, i32, i32, vp],
    "dtp_op_attn_small": [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, i64, i64, i64, vp, f32, vp],
    "dtp_op_flash_attn": [vp, vp, vp, i32, i64, vp, i32, i64, i32, i32, i32, i32, vp],
    "dtp_op_upsample2x": [vp, i32, i32, i32, i32, vp, vp],
    "dtp_op_im2col_s2": [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp],
    "dtp_op_ddim_step": [vp, vp, vp, i32, i32, f32, f32, f32, f32, vp],
    "dtp_op_pack_unet_input": [vp, vp, vp, i32, i32, vp, vp],
    "dtp_op_nchw_to_nhwc_pad": [vp, i32, i32, i32, i32, f32, vp, vp],
    "dtp_op_canvas_preprocess": [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "dtp_op_composite": [vp, vp, i32, i32, vp, vp, vp],
}

_PIPE = {
    "dtp_create": [vp, C.POINTER(vp)],
    "dtp_destroy": [vp],
    "dtp_set_tensor": [vp, cp, vp, C.POINTER(i64), i32, i32],
    "dtp_finalize_weights": [vp],
    "dtp_encode_patches": [vp, vp, vp, vp, vp],
    "dtp_set_condition": [vp, vp, vp, vp],
    "dtp_set_schedule": [vp, i32, C.POINTER(f32), C.POINTER(f32), C.POINTER(f32), f32, f32, i32],
    "dtp_infer": [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp],
    "dtp_stamp": [vp, i32, i32, vp, vp, i32, vp, vp, i32, vp, vp, vp],
    "dtp_vae_encode": [vp, i32, i32, vp, vp, vp, vp],
    "dtp_vae_decode": [vp, i32, i32, vp, vp, vp],
    "dtp_unet_forward": [vp, i32, i32, vp, vp, vp, i32, vp, vp],
    "dtp_get_counter": [vp, cp],
    "dtp_set_option": [vp, cp, i32],
    "dtp_profile_dump": [vp, cp],
}


class DtpError(RuntimeError):
    pass


def lib():
    """Load the shared library (building nothing: run `python -m diffusiontexturepainting_b200.build` first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DtpError(
            f"{LIB_PATH} is missing: build it with `python -m diffusiontexturepainting_b200.build` "
            "(there is no CPU / PyTorch fallback for the stamp path)")
    L = C.CDLL(LIB_PATH)
    for name, args in list(_OPS.items()) + list(_PIPE.items()):
        fn = getattr(L, name, None)
        if fn is None:
            continue
        fn.argtypes = args
        fn.restype = i32 if name not in ("dtp_destroy",) else None
    if hasattr(L, "dtp_get_counter"):
        L.dtp_get_counter.restype = i64
    L.dtp_ops_last_error.restype = cp
    if hasattr(L, "dtp_last_error"):
        L.dtp_last_error.argtypes = [vp]
        L.dtp_last_error.restype = cp
    _lib = L
    return L


def exported_symbols():
    return list(_OPS) + list(_PIPE) + ["dtp_ops_last_error", "dtp_last_error", "dtp_ops_set_debug_buffer"]


def ptr(t):
    """Device (or host) pointer of a torch tensor / None as c_void_p."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check_op(rc, what="dtp op"):
    if rc != 0:
        raise DtpError(f"{what} failed ({rc}): {lib().dtp_ops_last_error().decode()}")
