"""upfirdn2d with first- and second-order autograd on the sm_100a kernel.
Mirrors reference model/op/upfirdn2d.py:22-169: NCHW in/out, the image is handed to the kernel as
[N*C, H, W, 1]; the gradient is the same operator with up/down swapped, the flipped kernel and the
g_pad of upfirdn2d.py:116-121; the second-order gradient is the forward operator again (:65-88).
"""
from collections import abc

import torch
from torch.autograd import Function

from . import upfirdn2d_op


class _UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        go = grad_output.reshape(-1, out_size[0], out_size[1], 1)
        gi = upfirdn2d_op.upfirdn2d(go, grad_kernel, down[0], down[1], up[0], up[1], *g_pad)
        ctx.save_for_backward(kernel)
        ctx.cfg = (up, down, pad, in_size, out_size)
        return gi.view(in_size[0], in_size[1], in_size[2], in_size[3])

    @staticmethod
    def backward(ctx, gg_input):
        (kernel,) = ctx.saved_tensors
        up, down, pad, in_size, out_size = ctx.cfg
        ggi = gg_input.reshape(-1, in_size[2], in_size[3], 1)
        ggo = upfirdn2d_op.upfirdn2d(ggi, kernel, up[0], up[1], down[0], down[1], *pad)
        return (ggo.view(in_size[0], in_size[1], out_size[0], out_size[1]),) + (None,) * 8


class _UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kh, kw = kernel.shape
        _, channel, in_h, in_w = input.shape
        ctx.in_size = tuple(input.shape)
        out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) // down_y
        out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) // down_x
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        ctx.g_pad = (kw - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                     kh - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out = upfirdn2d_op.upfirdn2d(input.reshape(-1, in_h, in_w, 1), kernel, up_x, up_y, down_x, down_y,
                                     pad_x0, pad_x1, pad_y0, pad_y1)
        return out.view(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = None
        if ctx.needs_input_grad[0]:
            grad_input = _UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad,
                                                  ctx.g_pad, ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not isinstance(up, abc.Iterable):
        up = (up, up)
    if not isinstance(down, abc.Iterable):
        down = (down, down)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    if not input.is_cuda:
        raise RuntimeError("havatar_b200.op.upfirdn2d needs CUDA tensors (no CPU fallback)")
    if not (torch.is_grad_enabled() and input.requires_grad):
        # inference: straight to the kernel, no autograd bookkeeping (the flipped gradient kernel is not needed)
        _, channel, in_h, in_w = input.shape
        out = upfirdn2d_op.upfirdn2d(input.reshape(-1, in_h, in_w, 1), kernel, up[0], up[1], down[0], down[1], *pad)
        return out.view(-1, channel, out.shape[1], out.shape[2])
    return _UpFirDn2d.apply(input, kernel, tuple(up), tuple(down), tuple(pad))
