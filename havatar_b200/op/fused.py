"""Module-level stand-in for the reference's compiled extension `fused`
(reference model/op/fused_bias_act.cpp:18-32): same function name, argument order and meaning.

    fused_bias_act(input, bias, refer, act, grad, alpha, scale) -> Tensor

The callee allocates the output like the reference (torch::empty_like) and launches on torch's
current stream; an empty tensor means "no bias" / "no refer" (fused_bias_act_kernel.cu:79-80).
Non-CUDA tensors raise RuntimeError like the reference's CHECK_CUDA (fused_bias_act.cpp:10-16).
"""
import ctypes as C

import torch

from .. import _lib


def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    if not input.is_cuda:
        raise RuntimeError("input must be a CUDA tensor")
    if input.dtype != torch.float32:
        raise RuntimeError("havatar_b200 fused_bias_act is float32 only (the reference project never uses another dtype)")
    x = input.contiguous()
    b = bias.contiguous() if bias is not None and bias.numel() else None
    ref = refer.contiguous() if refer is not None and refer.numel() else None
    if b is not None and (not b.is_cuda or b.dtype != torch.float32):
        raise RuntimeError("bias must be a float32 CUDA tensor")
    if ref is not None and (not ref.is_cuda or ref.dtype != torch.float32 or ref.numel() != x.numel()):
        raise RuntimeError("refer must be a float32 CUDA tensor with input's numel")
    step_b = 1
    for i in range(2, x.dim()):            # fused_bias_act_kernel.cu:86-88
        step_b *= x.size(i)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = _lib.lib().hav_fused_bias_act(
            C.c_void_p(out.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(b.data_ptr()) if b is not None else None,
            C.c_void_p(ref.data_ptr()) if ref is not None else None, x.numel(), step_b,
            b.numel() if b is not None else 1, int(act), int(grad), float(alpha), float(scale), C.c_void_p(stream))
    _lib.check(rc, "hav_fused_bias_act")
    return out
