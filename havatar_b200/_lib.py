"""ctypes binding of libhavatar_b200.so (C ABI: include/havatar_b200.h).

There is no CPU fallback: lib() raises if the library is missing or stale symbols are found, and
every op in this package goes through it.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhavatar_b200.so")

HAV_ABI_VERSION = 3
PREC_FP32, PREC_BF16, PREC_FP16, PREC_FP16X3 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp16": PREC_FP16, "fp16x3": PREC_FP16X3}

_fp = C.c_void_p  # device pointers travel as integers


class RenderArgs(C.Structure):
    """hav_render_args (include/havatar_b200.h) -- field order and types must match the header."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("precision", C.c_int32), ("batch", C.c_int32), ("rays", C.c_int32),
        ("num_coarse", C.c_int32), ("num_fine", C.c_int32), ("plane_c", C.c_int32),
        ("plane_h", C.c_int32), ("plane_w", C.c_int32),
        ("vol_d", C.c_int32), ("vol_h", C.c_int32), ("vol_w", C.c_int32), ("flags", C.c_int32),
        ("plane_scale", C.c_float * 3), ("plane_trans", C.c_float * 3),
        ("skin_scale", C.c_float * 3), ("skin_trans", C.c_float * 3),
        ("ray_batch", _fp), ("background", _fp), ("inv_head_T", _fp), ("planes", _fp), ("wvol", _fp),
        ("w0", _fp), ("b0", _fp), ("w1", _fp), ("b1", _fp), ("w_alpha", _fp), ("b_alpha", _fp),
        ("w_feat", _fp), ("b_feat", _fp), ("w_rgb", _fp), ("b_rgb", _fp),
        ("t_rand", _fp), ("noise_coarse", _fp), ("u_rand", _fp), ("noise_fine", _fp),
        ("rgb_coarse", _fp), ("depth_coarse", _fp), ("acc_coarse", _fp), ("weights_max", _fp),
        ("rgb_fine", _fp), ("depth_fine", _fp), ("acc_fine", _fp), ("z_fine", _fp),
        ("workspace", _fp), ("workspace_bytes", C.c_uint64),
        ("camera", _fp), ("pixel_index", _fp), ("img_h", C.c_int32), ("img_w", C.c_int32), ("pdf_inds", _fp), ("range_status", _fp),
    ]


class RenderBwdArgs(C.Structure):
    """hav_render_bwd_args (include/havatar_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("grad_scale", C.c_float), ("fwd", C.POINTER(RenderArgs)),
        ("g_rgb_coarse", _fp), ("g_depth_coarse", _fp), ("g_acc_coarse", _fp),
        ("g_rgb_fine", _fp), ("g_depth_fine", _fp), ("g_acc_fine", _fp),
        ("g_planes", _fp), ("g_wvol", _fp),
        ("g_w0", _fp), ("g_b0", _fp), ("g_w1", _fp), ("g_b1", _fp), ("g_w_alpha", _fp), ("g_b_alpha", _fp),
        ("g_w_feat", _fp), ("g_b_feat", _fp), ("g_w_rgb", _fp), ("g_b_rgb", _fp),
        ("workspace", _fp), ("workspace_bytes", C.c_uint64),
    ]


class StyleLayer(C.Structure):
    """hav_style_layer (include/havatar_b200.h)."""
    _fields_ = [("mod_w", C.c_void_p), ("mod_b", C.c_void_p), ("wsq", C.c_void_p), ("cin", C.c_int32), ("cout", C.c_int32),
                ("latent_index", C.c_int32), ("s_off", C.c_int32), ("d_off", C.c_int32), ("mod_scale", C.c_float),
                ("mod_lr_mul", C.c_float), ("conv_scale", C.c_float)]


class ConvArgs(C.Structure):
    """hav_conv_args (include/havatar_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("precision", C.c_int32), ("batch", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("in_h", C.c_int32), ("in_w", C.c_int32), ("ksize", C.c_int32), ("up", C.c_int32), ("down", C.c_int32),
        ("act", C.c_int32), ("noise_per_sample", C.c_int32), ("noise_weight", C.c_float),
        ("in_layout", C.c_int32), ("out_layout", C.c_int32),
        ("x", _fp), ("wpack", _fp), ("in_scale", _fp), ("out_scale", _fp), ("noise", _fp), ("bias", _fp), ("out", _fp),
        ("residual", _fp),
    ]


class ConvWgradArgs(C.Structure):
    """hav_conv_wgrad_args (include/havatar_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("batch", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("in_h", C.c_int32),
        ("in_w", C.c_int32), ("ksize", C.c_int32), ("up", C.c_int32), ("down", C.c_int32), ("accumulate", C.c_int32),
        ("wscale", C.c_float), ("g", _fp), ("x", _fp), ("in_scale", _fp), ("out_scale", _fp), ("dw", _fp),
    ]


# symbol -> (restype, argtypes); tests check that every one of these is exported
SIGNATURES = {
    "hav_abi_version": (C.c_int, []),
    "hav_error_string": (C.c_char_p, [C.c_int]),
    "hav_fused_bias_act": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                     C.c_float, C.c_float, _fp]),
    "hav_upfirdn2d_cl": (C.c_int, [_fp, _fp, _fp] + [C.c_int] * 12 + [_fp, C.c_float, C.c_int, _fp, C.c_int, _fp]),
    "hav_upfirdn2d": (C.c_int, [_fp, _fp, _fp] + [C.c_int] * 14 + [_fp]),
    "hav_render_workspace_bytes": (C.c_uint64, [C.POINTER(RenderArgs)]),
    "hav_render_forward": (C.c_int, [C.POINTER(RenderArgs), _fp]),
    "hav_sample_pdf": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp]),
    "hav_render_backward_workspace_bytes": (C.c_uint64, [C.POINTER(RenderBwdArgs)]),
    "hav_render_backward": (C.c_int, [C.POINTER(RenderBwdArgs), _fp]),
    "hav_conv_wpack_bytes": (C.c_uint64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "hav_conv_pack_weights": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _fp]),
    "hav_bias_act_backward_splits": (C.c_int, [C.c_int, C.c_int, C.c_int64]),
    "hav_bias_act_backward": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_float, _fp]),
    "hav_noise_bias_act": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_float, _fp]),
    "hav_noise_bias_act_backward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float, _fp]),
    "hav_modconv_demod": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _fp]),
    "hav_conv2d_forward": (C.c_int, [C.POINTER(ConvArgs), _fp]),
    "hav_conv_tap_squares": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    "hav_style_plan_run": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, _fp]),
    "hav_conv2d_wgrad": (C.c_int, [C.POINTER(ConvWgradArgs), _fp]),
    "hav_rowscale_dot": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int64, _fp]),
    "hav_adam_flat": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, _fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _fp]),
    "hav_make_render_cond": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, _fp]),
    "hav_get_rays": (C.c_int, [_fp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                               C.c_float, _fp]),
}

_lib = None


class HavError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises HavError when the CUDA library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HavError(
            "libhavatar_b200.so is not built (%s). Run `python -m havatar_b200.build` (needs nvcc); "
            "havatar_b200 has no CPU fallback." % LIB_PATH)
    h = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(h, name)
        except AttributeError as e:
            raise HavError("libhavatar_b200.so does not export %s -- rebuild it" % name) from e
        fn.restype, fn.argtypes = res, args
    if h.hav_abi_version() != HAV_ABI_VERSION:
        raise HavError("libhavatar_b200.so ABI %d != binding ABI %d -- rebuild it" % (h.hav_abi_version(), HAV_ABI_VERSION))
    _lib = h
    return h


def check(rc, what):
    """Turn a C-ABI return code into a RuntimeError (the reference raises RuntimeError via TORCH_CHECK,
    model/op/fused_bias_act.cpp:10-16)."""
    if rc != 0:
        msg = lib().hav_error_string(int(rc)).decode()
        raise HavError("%s failed: %s (code %d)" % (what, msg, rc))
