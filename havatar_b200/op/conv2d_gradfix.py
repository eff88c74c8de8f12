"""Drop-in for the reference's `model/op/conv2d_gradfix.py` (public surface :1-75: `enabled`, `weight_gradients_disabled`,
`no_weight_gradients()`, `conv2d`, `conv_transpose2d`) on the tcgen05 convolution kernels.

The reference routes every StyleUNet convolution through these two functions so that cuDNN's backward can be differentiated
again for the R1 penalty (utils/styleUnet_util.py:72-79) and so that weight gradients can be skipped inside it.  Here the
forward, data-gradient and weight-gradient kernels are closed under differentiation (havatar_b200/conv.py), so gradients of any
order come out of the same three kernels; `no_weight_gradients()` makes the recorded backward skip the weight gradient.

Supported forms = the ones model/styleUnet.py issues: kernel 1x1 or 3x3, dilation 1, and
    conv2d            stride 1 / padding k//2   or   stride 2 / padding 0                     (:108-118, :281-291)
    conv_transpose2d  stride 2 / padding 0 / output_padding 0                                 (:264-270)
with groups == 1 or the per-sample grouped form of the fused ModulatedConv2d branch (input [1, G*Cin, H, W], weight
[G*Cout, Cin, k, k], groups = G, :253-297), which runs as G launches on the shared-weight kernel.  Anything else raises:
there is no library fallback."""
import contextlib

import torch

from .. import conv as _conv

enabled = True
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    _conv.WEIGHT_GRADIENTS_DISABLED[0] = True
    try:
        yield
    finally:
        weight_gradients_disabled = old
        _conv.WEIGHT_GRADIENTS_DISABLED[0] = old


def _pair(v):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    if v[0] != v[1]:
        raise NotImplementedError("conv2d_gradfix on havatar_b200: square stride / padding only, got %r" % (v,))
    return int(v[0])


def _one(x, w, up, down):
    return _conv.conv2d_autograd(x.contiguous(), w.contiguous(), None, None, 1.0, up=up, down=down)


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    k, stride, padding = int(weight.shape[-1]), _pair(stride), _pair(padding)
    if _pair(dilation) != 1 or weight.shape[-2] != k or k not in (1, 3) or (stride, padding) not in ((1, k // 2), (2, 0)):
        raise NotImplementedError("conv2d_gradfix on havatar_b200 supports k in {1,3}, dilation 1, (stride, padding) in "
                                  "{(1, k//2), (2, 0)}; got k=%d stride=%d padding=%d dilation=%r" % (k, stride, padding, dilation))
    down = 2 if stride == 2 else 1
    if groups == 1:
        y = _one(input, weight, 1, down)
    else:
        B, C, H, W = input.shape
        cin, cout = C // groups, weight.shape[0] // groups
        if weight.shape[1] != cin:
            raise ValueError("grouped conv2d: weight must be [groups*Cout, Cin/groups, k, k]")
        ys = [_one(input[:, g * cin:(g + 1) * cin], weight[g * cout:(g + 1) * cout], 1, down) for g in range(groups)]
        y = torch.cat(ys, dim=1)
    return y if bias is None else y + bias.view(1, -1, 1, 1)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    k = int(weight.shape[-1])
    if _pair(dilation) != 1 or _pair(stride) != 2 or _pair(padding) != 0 or _pair(output_padding) != 0 or k != 3:
        raise NotImplementedError("conv_transpose2d on havatar_b200 supports the StyleUNet form only: k=3, stride 2, padding 0")
    if groups == 1:
        y = _one(input, weight.transpose(0, 1), 2, 1)
    else:
        B, C, H, W = input.shape
        cin = C // groups
        ys = [_one(input[:, g * cin:(g + 1) * cin], weight[g * cin:(g + 1) * cin].transpose(0, 1), 2, 1) for g in range(groups)]
        y = torch.cat(ys, dim=1)
    return y if bias is None else y + bias.view(1, -1, 1, 1)
