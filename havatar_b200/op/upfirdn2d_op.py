"""Module-level stand-in for the reference's compiled extension `upfirdn2d`
(reference model/op/upfirdn2d.cpp:17-31): same function name, argument order and meaning.

    upfirdn2d(input[major,in_h,in_w,minor], kernel[kh,kw], up_x, up_y, down_x, down_y,
              pad_x0, pad_x1, pad_y0, pad_y1) -> Tensor[major,out_h,out_w,minor]
"""
import ctypes as C

import torch

from .. import _lib


def upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    if not input.is_cuda or not kernel.is_cuda:
        raise RuntimeError("input and kernel must be CUDA tensors")
    if input.dtype != torch.float32 or kernel.dtype != torch.float32:
        raise RuntimeError("havatar_b200 upfirdn2d is float32 only")
    if input.dim() != 4 or kernel.dim() != 2:
        raise RuntimeError("input must be [major,in_h,in_w,minor] and kernel [kh,kw]")
    x, k = input.contiguous(), kernel.contiguous()
    major, in_h, in_w, minor = [int(v) for v in x.shape]
    kh, kw = int(k.shape[0]), int(k.shape[1])
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y      # upfirdn2d_kernel.cu:236-241
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
    if out_h < 1 or out_w < 1:
        raise RuntimeError("upfirdn2d: empty output (%d x %d)" % (out_h, out_w))
    out = torch.empty((major, out_h, out_w, minor), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = _lib.lib().hav_upfirdn2d(C.c_void_p(out.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(k.data_ptr()),
                                      major, in_h, in_w, minor, kh, kw, int(up_x), int(up_y), int(down_x), int(down_y),
                                      int(pad_x0), int(pad_x1), int(pad_y0), int(pad_y1), C.c_void_p(stream))
    _lib.check(rc, "hav_upfirdn2d")
    return out
