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


def fused_bias_act_backward(grad_output, out, alpha, scale):
    """(grad_input, grad_bias) of FusedLeakyReLU in one pass (hav_bias_act_backward): the gated gradient of
    fused_bias_act(..., act=3, grad=1) and its sum over every dim but the channel (reference model/op/fused_act.py:31-45, which
    runs the kernel and then a separate full-tensor reduction)."""
    g, ref = grad_output.contiguous(), out.contiguous()
    if not g.is_cuda or g.dtype != torch.float32 or ref.dtype != torch.float32 or ref.shape != g.shape or g.dim() < 2:
        raise RuntimeError("fused_bias_act_backward needs float32 CUDA tensors of one shape [B, C, ...]")
    B, Cc = int(g.shape[0]), int(g.shape[1])
    inner = 1
    for i in range(2, g.dim()):
        inner *= int(g.size(i))
    L = _lib.lib()
    grad_input = torch.empty_like(g)
    if g.numel() == 0:
        return grad_input, g.new_zeros(Cc)
    splits = int(L.hav_bias_act_backward_splits(B, Cc, inner))
    partials = torch.empty((splits, Cc), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        stream = torch.cuda.current_stream(g.device).cuda_stream
        _lib.check(L.hav_bias_act_backward(C.c_void_p(grad_input.data_ptr()), C.c_void_p(partials.data_ptr()), C.c_void_p(g.data_ptr()),
                                           C.c_void_p(ref.data_ptr()), B, Cc, inner, splits, float(alpha), float(scale), C.c_void_p(stream)),
                   "hav_bias_act_backward")
    return grad_input, (partials[0] if splits == 1 else partials.sum(0))
