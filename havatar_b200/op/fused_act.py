"""FusedLeakyReLU / fused_leaky_relu with first- and second-order autograd on the sm_100a kernel.
Mirrors reference model/op/fused_act.py:23-122 (same names, arguments and gradient definitions):
  y  = lrelu(x + b, slope) * scale
  dx = dy * (y > 0 ? 1 : slope) * scale          (gated by the sign of the saved OUTPUT, :31-33)
  db = sum of dx over every dim but the channel dim (:36-45)
  second order: the same gate applied to (ddx + ddb) (:49-56).
No CPU branch: CPU tensors raise (the reference's CPU branch is its oracle, fused_act.py:107-119).
"""
import torch
from torch import nn
from torch.autograd import Function

from . import fused


class _LReLUBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, has_bias, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        empty = grad_output.new_empty(0)
        if has_bias and grad_output.dim() >= 2:      # gradient and its per-channel sum in one pass over the tensor
            grad_input, grad_bias = fused.fused_bias_act_backward(grad_output, out, negative_slope, scale)
            return grad_input, grad_bias.detach()
        grad_input = fused.fused_bias_act(grad_output.contiguous(), empty, out, 3, 1, negative_slope, scale)
        return grad_input, empty

    @staticmethod
    def backward(ctx, gg_input, gg_bias):
        (out,) = ctx.saved_tensors
        if gg_bias is None:
            gg_bias = gg_input.new_empty(0)
        gg_out = fused.fused_bias_act(gg_input.contiguous(), gg_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gg_out, None, None, None, None


class _LReLU(Function):
    @staticmethod
    def forward(ctx, x, bias, negative_slope, scale):
        empty = x.new_empty(0)
        ctx.has_bias = bias is not None
        out = fused.fused_bias_act(x, bias if bias is not None else empty, empty, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        grad_input, grad_bias = _LReLUBackward.apply(grad_output, out, ctx.has_bias, ctx.negative_slope, ctx.scale)
        return grad_input, (grad_bias if ctx.has_bias else None), None, None


class _NoiseLReLU(Function):
    """lrelu(x + weight * noise + bias) * scale in one launch, and in backward one pass that yields grad_x, grad_bias and
    grad_weight (hav_noise_bias_act / hav_noise_bias_act_backward).  First order only (the generators' StyledConv tails; the
    discriminator, whose double backward R1 needs, has no noise injection)."""

    @staticmethod
    def forward(ctx, x, noise, weight, bias, negative_slope, scale):
        import ctypes as C

        from .. import _lib

        x, noise = x.contiguous(), noise.contiguous()
        B, Cc = int(x.shape[0]), int(x.shape[1])
        inner = x.numel() // max(B * Cc, 1)
        per_sample = int(noise.shape[0] == B and B > 1)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(_lib.lib().hav_noise_bias_act(C.c_void_p(out.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(bias.data_ptr()),
                                                     C.c_void_p(noise.data_ptr()), C.c_void_p(weight.data_ptr()), B, Cc, inner, per_sample,
                                                     float(negative_slope), float(scale), C.c_void_p(st)), "hav_noise_bias_act")
        ctx.save_for_backward(out, noise)
        ctx.cfg = (negative_slope, scale, per_sample)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        import ctypes as C

        from .. import _lib

        out, noise = ctx.saved_tensors
        negative_slope, scale, per_sample = ctx.cfg
        g = grad_output.contiguous()
        B, Cc = int(g.shape[0]), int(g.shape[1])
        inner = g.numel() // max(B * Cc, 1)
        L = _lib.lib()
        splits = int(L.hav_bias_act_backward_splits(B, Cc, inner))
        gx = torch.empty_like(g)
        parts = torch.empty((2, splits, Cc), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            st = torch.cuda.current_stream(g.device).cuda_stream
            _lib.check(L.hav_noise_bias_act_backward(C.c_void_p(gx.data_ptr()), C.c_void_p(parts[0].data_ptr()), C.c_void_p(parts[1].data_ptr()),
                                                     C.c_void_p(g.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(noise.data_ptr()), B, Cc,
                                                     inner, per_sample, splits, float(negative_slope), float(scale), C.c_void_p(st)),
                       "hav_noise_bias_act_backward")
        sums = parts.sum(1)                      # [2, C]: grad_bias | per-channel share of grad_weight
        return gx, None, sums[1].sum().reshape(1), sums[0], None, None


def noise_leaky_relu(x, noise, weight, bias, negative_slope=0.2, scale=2 ** 0.5):
    """StyledConv's tail (model/styleUnet.py:596-598): fused_leaky_relu(x + weight * noise, bias).  x [B,C,H,W] float32 CUDA, noise
    [B or 1,1,H,W], weight [1] (NoiseInjection.weight), bias [C]."""
    fits = (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and noise.dim() == 4 and noise.shape[1] == 1 and
            noise.shape[0] in (1, x.shape[0]) and tuple(noise.shape[2:]) == tuple(x.shape[2:]) and bias is not None and weight.numel() == 1
            and not noise.requires_grad)
    if not fits:
        return fused_leaky_relu(x + weight * noise, bias, negative_slope, scale)
    return _NoiseLReLU.apply(x, noise.detach().float(), weight, bias, negative_slope, scale)


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    if not input.is_cuda:
        raise RuntimeError("havatar_b200.op.fused_leaky_relu needs CUDA tensors (no CPU fallback)")
    return _LReLU.apply(input.contiguous(), bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    """state_dict key `bias` [channel] as in the reference (fused_act.py:90-104)."""

    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
