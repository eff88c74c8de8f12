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
