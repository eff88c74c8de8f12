"""Backward of the tensor-core convolution (conv.conv2d_autograd: hav_conv2d_forward on the transposed weight image for the
data gradient, hav_conv2d_wgrad for the weight gradient, hav_rowscale_dot for the modulation gradients) against torch fp32
autograd over the same formula -- ModulatedConv2d's shared-weight branch / EqualConv2d (model/styleUnet.py:225-251, :108-118),
i.e. what the reference obtains from cuDNN through model/op/conv2d_gradfix.py.
Tolerance: bf16 operands (8-bit mantissa) with fp32 accumulation -> 2e-2 of each gradient tensor's range (stated)."""
import pytest
import torch
import torch.nn.functional as F

from havatar_b200 import conv

pytestmark = pytest.mark.gpu


def _ref(x, w, s, d, wscale, up, down):
    k = w.shape[-1]
    xs = x if s is None else x * s[:, :, None, None]
    if up == 2:
        y = F.conv_transpose2d(xs, (w * wscale).transpose(0, 1), stride=2, padding=0)
    elif down == 2:
        y = F.conv2d(xs, w * wscale, stride=2, padding=0)
    else:
        y = F.conv2d(xs, w * wscale, padding=k // 2)
    return y if d is None else y * d[:, :, None, None]


CASES = [
    # B, Cin, Cout, H, W, k, up, down, modulated
    (2, 64, 128, 16, 16, 3, 1, 1, True),
    (1, 512, 512, 16, 32, 3, 1, 1, True),        # several ci tiles (48) with a ragged last one, 4 co tiles
    (3, 40, 24, 13, 21, 3, 1, 1, True),          # ragged everything: tiles, channels, unaligned rows (scalar staging path)
    (2, 7, 32, 24, 24, 3, 1, 1, False),          # EqualConv2d from a 7-channel condition image
    (2, 128, 12, 32, 32, 1, 1, 1, True),         # ToRGB: 1x1, no demodulation in the reference (still tested with one)
    (2, 96, 64, 8, 8, 3, 2, 1, True),            # upsampling StyledConv: conv_transpose2d stride 2
    (1, 64, 64, 17, 16, 3, 2, 1, True),
    (2, 64, 96, 17, 17, 3, 1, 2, False),         # ConvLayer(downsample=True) after its blur: stride 2, pad 0
    (2, 32, 48, 18, 20, 3, 1, 2, False),         # even input size: the last row / column reaches no output
    (1, 256, 128, 64, 64, 3, 1, 1, True),        # many position tiles per CTA: the 3-stage ring wraps
]


@pytest.mark.parametrize("B,Cin,Cout,H,W,k,up,down,mod", CASES)
def test_conv_backward_matches_torch_autograd(B, Cin, Cout, H, W, k, up, down, mod):
    torch.manual_seed(B * 1000 + Cin + Cout + H)
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    x = torch.randn(B, Cin, H, W, device=dev, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, device=dev, requires_grad=True)
    s = (torch.rand(B, Cin, device=dev) + 0.5).requires_grad_(True) if mod else None
    d = (torch.rand(B, Cout, device=dev) + 0.5).requires_grad_(True) if mod else None
    wscale = 1.0 / (Cin * k * k) ** 0.5
    leaves = [t for t in (x, w, s, d) if t is not None]
    y_ref = _ref(x, w, s, d, wscale, up, down)
    go = torch.randn_like(y_ref) * 1e-4          # small cotangents, like a mean-reduced loss (bf16 keeps the exponent range)
    g_ref = torch.autograd.grad(y_ref, leaves, go)
    y = conv.conv2d_autograd(x, w, s, d, wscale, up=up, down=down)
    assert y.shape == y_ref.shape
    assert float((y - y_ref).detach().abs().max()) < 2e-2 * float(y_ref.detach().abs().max())
    g = torch.autograd.grad(y, leaves, go)
    torch.cuda.synchronize()
    for name, a, b in zip(("x", "w", "s", "d"), g, g_ref):
        assert a.shape == b.shape, name
        err = float((a - b).abs().max()) / float(b.abs().max())
        assert err < 2e-2, (name, err)


def test_wgrad_accumulates_and_linear():
    torch.manual_seed(3)
    x = torch.randn(2, 48, 16, 16, device="cuda")
    g = torch.randn(2, 32, 16, 16, device="cuda")
    a = conv.conv_wgrad(g, x, 3)
    b = conv.conv_wgrad(g, x, 3, out=a.clone())           # accumulate: a + a
    assert float((b - 2 * a).abs().max()) <= 1e-5 * float(a.abs().max())
    c = conv.conv_wgrad(2 * g, x, 3, wscale=0.5)
    assert float((c - a).abs().max()) <= 1e-5 * float(a.abs().max())


def test_rowscale_dot():
    torch.manual_seed(4)
    for n in (64 * 64, 17 * 13):
        a, x = torch.randn(3, 5, n, 1, device="cuda"), torch.randn(3, 5, n, 1, device="cuda")
        s = torch.randn(3, 5, device="cuda")
        out, dot = conv.rowscale_dot(a, x, s)
        assert torch.equal(out, a * s[:, :, None, None])
        assert float((dot - (a * x).sum((2, 3))).abs().max()) < 1e-3


@pytest.mark.parametrize("Cin,Cout,H,k,up,down,mod", [(32, 48, 16, 3, 1, 1, False), (24, 32, 17, 3, 1, 2, False), (16, 8, 12, 1, 1, 1, False),
                                                      (32, 32, 8, 3, 2, 1, True), (40, 24, 12, 3, 1, 1, True)])
def test_conv_double_backward_matches_torch(Cin, Cout, H, k, up, down, mod):
    """R1-style second order (utils/styleUnet_util.py:72-79): L = |d sum(phi(y)) / dx|^2, gradients of L w.r.t. the weight (and
    the modulation).  The input gradient is recorded with create_graph=True through _Dgrad, whose backward runs _Fwd / _Wgrad."""
    torch.manual_seed(Cin + Cout + H)
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    B = 2
    x = torch.randn(B, Cin, H, H, device=dev, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, device=dev, requires_grad=True)
    s = (torch.rand(B, Cin, device=dev) + 0.5).requires_grad_(True) if mod else None
    d = (torch.rand(B, Cout, device=dev) + 0.5).requires_grad_(True) if mod else None
    wscale = 1.0 / (Cin * k * k) ** 0.5
    leaves = [t for t in (w, s, d) if t is not None]

    def penalty(y):
        gx, = torch.autograd.grad((y * y).sum() * 0.5, x, create_graph=True)        # phi = y^2 / 2 so the first gradient depends on y
        return gx.pow(2).sum()

    ref = torch.autograd.grad(penalty(_ref(x, w, s, d, wscale, up, down)), leaves)
    got = torch.autograd.grad(penalty(conv.conv2d_autograd(x, w, s, d, wscale, up=up, down=down)), leaves)
    torch.cuda.synchronize()
    for name, a, b in zip(("w", "s", "d"), got, ref):
        err = float((a - b).abs().max()) / float(b.abs().max())
        assert err < 3e-2, (name, err)      # three chained bf16-operand convolutions
