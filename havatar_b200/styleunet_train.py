"""Differentiable forward of the StyleUNet mirrors (havatar_b200/styleunet.py) for the training steps
(train_avatar.py:112-149, train_avatarHD.py:211-280).

styleunet.py's default forward is the inference path: fused tcgen05 convolutions on channels-last fp16 hand-over tensors,
no autograd graph.  When gradients are enabled the same modules (same parameters, same state_dict) run through the functions
below instead: the reference's layer formulas (model/styleUnet.py, cited per function) on NCHW fp32 tensors with torch
autograd.  What runs underneath:
  * upfirdn2d (Blur / Upsample / Downsample / Haar) and fused_leaky_relu: OUR sm_100a kernels with their first- and
    second-order autograd (havatar_b200/op, the `model/op` replacements),
  * the convolutions: conv.conv2d_autograd -- the tcgen05 forward kernel, and a backward made of the same kernel on the
    transposed weight image (data gradient), the tcgen05 weight-gradient kernel (hav_conv2d_wgrad) and one row kernel for the
    modulation / demodulation gradients, in ModulatedConv2d's shared-weight formulation (the reference's own non-fused branch,
    styleUnet.py:225-251).
    The three convolution kernels are closed under differentiation (conv.py: _Fwd / _Dgrad / _Wgrad), so second-order passes
    (the R1 penalty every 16th discriminator step, utils/styleUnet_util.py:72-79) run on them too.  `library_convs()` switches the
    same formulas to torch.nn.functional.conv2d / conv_transpose2d (cuDNN) -- used only by tests and timing comparisons.
"""
import contextlib

import torch
import torch.nn.functional as F

from . import conv as hconv
from .op import fused_leaky_relu
from .op.fused_act import noise_leaky_relu

_NATIVE = [True]


@contextlib.contextmanager
def library_convs():
    """Run the convolutions of the enclosed forward on torch's own ops (a comparison arm for tests / timing, not the product path)."""
    prev, _NATIVE[0] = _NATIVE[0], False
    try:
        yield
    finally:
        _NATIVE[0] = prev


def _native(x):
    return _NATIVE[0] and x.is_cuda and x.dtype == torch.float32


def equal_linear(m, x):
    """EqualLinear.forward (styleUnet.py:150-162)."""
    if m.activation:
        return fused_leaky_relu(F.linear(x, m.weight * m.scale), m.bias * m.lr_mul)
    return F.linear(x, m.weight * m.scale, bias=None if m.bias is None else m.bias * m.lr_mul)


def style_mlp(seq, x):
    for layer in seq:
        x = equal_linear(layer, x) if hasattr(layer, "lr_mul") else layer(x)
    return x


def conv_layer(m, x):
    """ConvLayer (styleUnet.py:326-368): [Blur] -> EqualConv2d (:108-118) -> [FusedLeakyReLU]."""
    mods = list(m)
    if len(mods) > 1 and hasattr(mods[0], "kernel"):
        x = mods[0](x)
        mods = mods[1:]
    conv = mods[0]
    k = conv.weight.shape[-1]
    if _native(x) and k in (1, 3) and ((conv.stride == 1 and conv.padding == k // 2) or (conv.stride == 2 and conv.padding == 0 and k == 3)):
        out = hconv.conv2d_autograd(x, conv.weight, None, None, conv.scale, down=conv.stride, cache=conv._cache.store)
        if conv.bias is not None:
            out = out + conv.bias[None, :, None, None]
    else:
        out = F.conv2d(x, conv.weight * conv.scale, bias=conv.bias, stride=conv.stride, padding=conv.padding)
    if m.activate:
        out = fused_leaky_relu(out, mods[1].bias)
    return out


def mod_conv(m, x, style):
    """ModulatedConv2d.forward (styleUnet.py:222-297), shared-weight form:  demod * conv(x * s, scale * W)."""
    s = equal_linear(m.modulation, style)                                  # [B,Cin]
    k = m.weight.shape[-1]
    if _native(x) and (k == 3 or (k == 1 and not m.upsample)) and m.padding == k // 2:
        w0 = m.weight[0]
        d = None
        if m.demodulate:                                                   # :256-258
            d = torch.rsqrt((s * s) @ (w0 * w0).sum(dim=(2, 3)).t() * (m.scale * m.scale) + m.eps)
        out = hconv.conv2d_autograd(x, w0, s, d, m.scale, up=2 if m.upsample else 1, cache=m._cache.store)
        return m.blur(out) if m.upsample else out                          # :271-277
    w = m.weight[0] * m.scale                                              # [Cout,Cin,k,k]
    x = x * s[:, :, None, None]
    if m.upsample:
        out = F.conv_transpose2d(x, w.transpose(0, 1), stride=2, padding=0)       # :264-270
    else:
        out = F.conv2d(x, w, padding=m.padding)                            # :289-291
    if m.demodulate:                                                       # :256-258
        d = torch.rsqrt((s * s) @ (w * w).sum(dim=(2, 3)).t() + m.eps)     # [B,Cout]
        out = out * d[:, :, None, None]
    if m.upsample:
        out = m.blur(out)                                                  # :271-277
    return out


def styled_conv(m, x, style, noise=None):
    """StyledConv.forward (styleUnet.py:593-599) with NoiseInjection (:300-310)."""
    out = mod_conv(m.conv, x, style)
    if noise is None:
        noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
    return noise_leaky_relu(out, noise, m.noise.weight, m.activate.bias)      # one launch; one backward pass for all three gradients


def to_rgb(m, x, style, skip=None):
    """ToRGB.forward (styleUnet.py:617-628)."""
    out = mod_conv(m.conv, x, style) + m.bias
    if skip is not None:
        skip = m.dwt(m.upsample(m.iwt(skip))) if m.use_wt else m.upsample(skip)
        out = out + skip
    return out


def from_rgb(m, x, skip=None):
    """FromRGB.forward (styleUnet.py:455-467)."""
    if m.downsample:
        x = m.dwt(m.downsample(m.iwt(x))) if m.use_wt else m.downsample(x)
    out = conv_layer(m.conv, x)
    if skip is not None:
        out = out + skip
    return x, out


def conv_block(m, x):
    return conv_layer(m.conv2, conv_layer(m.conv1, x))


def cond_encoder(net, cond_img):
    """styleUnet.py:1379-1388 / :847-856."""
    cond_out = conv_layer(net.conv_in, cond_img)
    feats = [cond_out]
    for fr, cc in zip(net.from_rgbs, net.cond_convs):
        cond_img, cond_out = from_rgb(fr, cond_img, cond_out)
        cond_out = conv_block(cc, cond_out)
        feats.append(cond_out)
    return feats


def swgan_unet_forward(net, latent, condition_img, noise):
    """SWGAN_unet.forward after the style / noise bookkeeping (styleUnet.py:1379-1410)."""
    from .styleunet import _SkipFork

    feats = cond_encoder(net, condition_img)
    i, skip, out = 0, None, None
    fork = _SkipFork(condition_img.device)          # the ToRGB pyramid on a side stream (its backward follows it there)
    for conv1, conv2, n1, n2, rgb in zip(net.convs[::2], net.convs[1::2], noise[::2], noise[1::2], net.to_rgbs):
        if i == 0:
            out = conv_layer(net.comb_convs[-1], feats[-1])
        elif i < 2 * len(net.comb_convs):
            out = conv_layer(net.comb_convs[-1 - (i // 2)], torch.cat([out, feats[-1 - (i // 2)]], dim=1))
        out = styled_conv(conv1, out, latent[:, i], n1)
        out = styled_conv(conv2, out, latent[:, i + 1], n2)
        with fork.branch(out):
            skip = to_rgb(rgb, out, latent[:, i + 2], skip)
        i += 2
    return net.iwt(fork.join(skip))


def stylegan_zxc_forward(net, latent, cond_feats, noise):
    """StyleGAN_zxc.forward after the style / noise bookkeeping (styleUnet.py:847-878), no_skip configuration."""
    feats = cond_encoder(net, cond_feats)
    out = styled_conv(net.conv1, net.input(latent), latent[:, 0], noise[0])
    i = 1
    for conv1, conv2, n1, n2 in zip(net.convs[::2], net.convs[1::2], noise[1::2], noise[2::2]):
        if 1 < i <= 2 * len(feats) + 1:
            out = conv_layer(net.comb_convs[-(i // 2)], torch.cat([out, feats[-(i // 2)]], dim=1))
        out = styled_conv(conv1, out, latent[:, i], n1)
        out = styled_conv(conv2, out, latent[:, i + 1], n2)
        i += 2
    return conv_layer(net.conv_out, out)


def discriminator_forward(net, x):
    """Discriminator.forward (styleUnet.py:524-562)."""
    x = net.dwt(x)
    out = None
    for fr, block in zip(net.from_rgbs, net.convs):
        x, out = from_rgb(fr, x, out)
        out = conv_block(block, out)
    _, out = from_rgb(net.from_rgbs[-1], x, out)
    b, c, h, w = out.shape
    group = min(b, net.stddev_group)                                       # minibatch standard deviation (:539-545)
    sd = out.view(group, -1, net.stddev_feat, c // net.stddev_feat, h, w)
    sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8).mean([2, 3, 4], keepdims=True).squeeze(2)
    out = torch.cat([out, sd.repeat(group, 1, h, w)], 1)
    out = conv_layer(net.final_conv, out)
    return style_mlp(net.final_linear, out.view(b, -1))
