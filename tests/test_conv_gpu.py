"""Tensor-core convolution (hav_conv2d_forward) against a plain PyTorch fp32 reference of the same op, and the fused
modulated form against the reference's own formulation (model/styleUnet.py:253-297 restated with torch ops).
Operands are rounded to fp16, accumulation is fp32: tolerance 1e-2 relative to the output scale.  pytest -m gpu."""
import math

import pytest
import torch
import torch.nn.functional as F

from havatar_b200 import conv

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    """The torch side of every comparison is a true fp32 reference: cuDNN / cuBLAS TF32 off (torch enables cuDNN TF32 by default)."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("B,Cin,Cout,H,W,k", [(1, 64, 128, 32, 32, 3), (2, 512, 512, 16, 16, 3), (1, 7, 256, 40, 24, 3),
                                              (2, 128, 12, 32, 32, 1), (1, 1024, 512, 8, 8, 3), (1, 64, 64, 13, 9, 3)])
def test_plain_conv_matches_torch(B, Cin, Cout, H, W, k):
    torch.manual_seed(0)
    x = torch.randn(B, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda")
    scale = 1 / math.sqrt(Cin * k * k)
    bias = torch.randn(Cout, device="cuda")
    ref = F.conv2d(x, w * scale, bias=bias, padding=k // 2)
    got = conv.conv2d(x, conv.pack_weights(w, scale), bias=bias)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-2
    # bias + leaky-relu * sqrt(2) epilogue (ConvLayer: EqualConv2d -> FusedLeakyReLU, styleUnet.py:352-366)
    ref2 = F.leaky_relu(F.conv2d(x, w * scale, padding=k // 2) + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    got2 = conv.conv2d(x, conv.pack_weights(w, scale), bias=bias, act=True)
    assert rel_err(got2, ref2) < 1e-2


@pytest.mark.parametrize("B,Cin,Cout,H", [(1, 64, 64, 16), (2, 512, 256, 8), (1, 128, 128, 33)])
def test_strided_and_transposed(B, Cin, Cout, H):
    torch.manual_seed(1)
    x = torch.randn(B, Cin, H, H, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda")
    scale = 1 / math.sqrt(Cin * 9)
    ref = F.conv2d(x, w * scale, stride=2, padding=0)
    got = conv.conv2d(x, conv.pack_weights(w, scale), down=2)
    assert got.shape == ref.shape and rel_err(got, ref) < 1e-2
    ref = F.conv_transpose2d(x, (w * scale).transpose(0, 1), stride=2, padding=0)
    got = conv.conv2d(x, conv.pack_weights(w, scale, up=2), up=2)
    assert got.shape == ref.shape and rel_err(got, ref) < 1e-2


@pytest.mark.parametrize("upsample", [False, True])
@pytest.mark.parametrize("demodulate", [True, False])
def test_modulated_conv_matches_reference_formulation(upsample, demodulate):
    """ModulatedConv2d fused branch (styleUnet.py:253-297): per-sample weights + grouped conv, in torch fp32."""
    torch.manual_seed(2)
    B, Cin, Cout, H, k = 2, 128, 64, 16, 3 if demodulate else 1
    if upsample and k == 1:
        pytest.skip("ToRGB is never an upsampling conv")
    x = torch.randn(B, Cin, H, H, device="cuda")
    W = torch.randn(1, Cout, Cin, k, k, device="cuda")
    s = torch.randn(B, Cin, device="cuda") * 0.3 + 1.0
    scale = 1 / math.sqrt(Cin * k * k)
    weight = scale * W * s.view(B, 1, Cin, 1, 1)
    if demodulate:
        demod = torch.rsqrt(weight.pow(2).sum([2, 3, 4]) + 1e-8)
        weight = weight * demod.view(B, Cout, 1, 1, 1)
    if upsample:
        wt = weight.transpose(1, 2).reshape(B * Cin, Cout, k, k)
        ref = F.conv_transpose2d(x.view(1, B * Cin, H, H), wt, padding=0, stride=2, groups=B)
        ref = ref.view(B, Cout, ref.shape[-2], ref.shape[-1])
    else:
        ref = F.conv2d(x.view(1, B * Cin, H, H), weight.view(B * Cout, Cin, k, k), padding=k // 2, groups=B).view(B, Cout, H, H)
    noise = torch.randn(1, 1, ref.shape[-2], ref.shape[-1], device="cuda")
    bias = torch.randn(Cout, device="cuda")
    ref = F.leaky_relu(ref + 0.37 * noise + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    d = conv.modconv_demod(W[0], s, scale) if demodulate else None
    if demodulate:
        assert torch.allclose(d, demod, rtol=1e-5, atol=1e-7)
    got = conv.conv2d(x, conv.pack_weights(W[0], scale, up=2 if upsample else 1), in_scale=s, out_scale=d, noise=noise, noise_weight=0.37,
                      bias=bias, act=True, up=2 if upsample else 1)
    assert got.shape == ref.shape and rel_err(got, ref) < 1e-2


def test_full_size_layer_linearity():
    """256x256, 128 -> 128 channels (the largest SWGAN_unet decoder layer at 128 -> 512): conv is linear in x."""
    torch.manual_seed(3)
    x1 = torch.randn(1, 128, 256, 256, device="cuda")
    x2 = torch.randn(1, 128, 256, 256, device="cuda")
    pw = conv.pack_weights(torch.randn(128, 128, 3, 3, device="cuda"), 1 / math.sqrt(128 * 9))
    y = conv.conv2d(x1 + x2, pw)
    y12 = conv.conv2d(x1, pw) + conv.conv2d(x2, pw)
    assert rel_err(y, y12) < 5e-3
    assert torch.isfinite(y).all()


@pytest.mark.parametrize("up,down", [(1, 1), (2, 1), (1, 2)])
def test_channels_last_fp16_hand_over(up, down):
    """NHWC fp16 in / out (the StyleUNet's internal layout) gives the same convolution as the NCHW fp32 path."""
    torch.manual_seed(4)
    B, Cin, Cout, H = 2, 128, 64, 24
    x = torch.randn(B, Cin, H, H, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda")
    s = torch.rand(B, Cin, device="cuda") + 0.5
    d = torch.rand(B, Cout, device="cuda") + 0.5
    bias = torch.randn(Cout, device="cuda")
    pw = conv.pack_weights(w, 1 / math.sqrt(Cin * 9), up=up)
    ref = conv.conv2d(x, pw, in_scale=s, out_scale=d, bias=bias, act=True, up=up, down=down)
    xcl = x.permute(0, 2, 3, 1).contiguous().half()
    got_cl = conv.conv2d(xcl, pw, in_scale=s, out_scale=d, bias=bias, act=True, up=up, down=down, out_cl=True)
    assert got_cl.dtype == torch.float16 and got_cl.shape == (B, ref.shape[2], ref.shape[3], Cout)
    assert rel_err(conv.to_nchw(got_cl), ref) < 5e-3
    got_mixed = conv.conv2d(xcl, pw, in_scale=s, out_scale=d, bias=bias, act=True, up=up, down=down, out_cl=False)
    assert rel_err(got_mixed, ref) < 5e-3


@pytest.mark.parametrize("up,down,pad", [(1, 1, (1, 1)), (1, 1, (2, 2)), (2, 1, (2, 1)), (1, 2, (1, 1))])
def test_upfirdn2d_channels_last_with_fused_tail(up, down, pad):
    from havatar_b200 import op

    torch.manual_seed(5)
    B, C, H, W = 2, 16, 13, 18
    x = torch.randn(B, C, H, W, device="cuda")
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
    k = k[None, :] * k[:, None]
    k = k / k.sum() * (up * up)
    ref = op.upfirdn2d(x.half().float(), k, up=up, down=down, pad=pad)
    noise = torch.randn(1, 1, ref.shape[2], ref.shape[3], device="cuda")
    bias = torch.randn(C, device="cuda")
    ref_tail = F.leaky_relu(ref + 0.3 * noise + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    xcl = x.permute(0, 2, 3, 1).contiguous().half()
    got = conv.upfirdn2d_cl(xcl, k, up=up, down=down, pad=pad)
    assert rel_err(conv.to_nchw(got), ref) < 2e-3
    got = conv.upfirdn2d_cl(xcl, k, up=up, down=down, pad=pad, noise=noise, noise_weight=0.3, bias=bias, act=True)
    assert rel_err(conv.to_nchw(got), ref_tail) < 2e-3


@pytest.mark.parametrize("C,H,W,pad,sep", [(64, 37, 45, (1, 1), True), (128, 16, 16, (2, 2), True), (64, 33, 70, (2, 1), True),
                                            (192, 19, 35, (1, 1), False), (64, 5, 3, (2, 2), True)])
def test_blur_channels_last_tma_tiles(C, H, W, pad, sep):
    """hav_upfirdn2d_cl's TMA-tiled 4x4 path (C % 64 == 0; csrc/fir_cl.cu): ragged tile edges, the zero padding supplied by the
    tensor map's out-of-bounds fill, several channel blocks / images per CTA, rank-one and general taps, per-sample noise."""
    from havatar_b200 import op

    torch.manual_seed(11)
    B = 3
    x = torch.randn(B, C, H, W, device="cuda")
    if sep:
        k = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
        k = k[None, :] * k[:, None]
        k = k / k.sum() * 4
    else:
        k = torch.randn(4, 4, device="cuda")
    ref = op.upfirdn2d(x.half().float(), k, pad=pad)
    noise = torch.randn(B, 1, ref.shape[2], ref.shape[3], device="cuda")
    bias = torch.randn(C, device="cuda")
    ref_tail = F.leaky_relu(ref + 0.3 * noise + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    xcl = x.permute(0, 2, 3, 1).contiguous().half()
    got = conv.upfirdn2d_cl(xcl, k, pad=pad)
    assert tuple(got.shape) == (B, ref.shape[2], ref.shape[3], C)
    assert rel_err(conv.to_nchw(got), ref) < 2e-3
    got = conv.upfirdn2d_cl(xcl, k, pad=pad, noise=noise, noise_weight=0.3, bias=bias, act=True)
    assert rel_err(conv.to_nchw(got), ref_tail) < 2e-3


@pytest.mark.parametrize("B,Cin,Cout,H,k,up,down", [(1, 512, 512, 16, 3, 1, 1), (2, 320, 128, 8, 3, 1, 1), (1, 512, 12, 32, 1, 1, 1),
                                                   (1, 512, 512, 8, 3, 2, 1), (1, 1024, 512, 17, 3, 1, 2), (1, 192, 72, 12, 3, 2, 1),
                                                   (1, 1024, 512, 16, 3, 1, 1)])
def test_cluster_split_k_layers(B, Cin, Cout, H, k, up, down):
    """Few output tiles x many channel blocks: the channel blocks of a tile are split over the CTAs of a cluster and the partial
    tiles reduced through distributed shared memory (conv_tc.cu).  Uneven splits (5 blocks over 4 CTAs), the four-accumulator
    transposed form, the subsampled stride-2 form, narrow 1x1 heads, NCHW fp32 and channels-last fp16 tensors, full fused tail."""
    torch.manual_seed(6)
    x = torch.randn(B, Cin, H, H, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda")
    s = torch.rand(B, Cin, device="cuda") + 0.5
    d = torch.rand(B, Cout, device="cuda") + 0.5
    bias = torch.randn(Cout, device="cuda")
    scale = 1 / math.sqrt(Cin * k * k)
    xs = x * s.view(B, Cin, 1, 1)
    if up == 2:
        ref = F.conv_transpose2d(xs, (w * scale).transpose(0, 1), stride=2, padding=0)
    elif down == 2:
        ref = F.conv2d(xs, w * scale, stride=2, padding=0)
    else:
        ref = F.conv2d(xs, w * scale, padding=k // 2)
    noise = torch.randn(B, 1, ref.shape[2], ref.shape[3], device="cuda")
    ref = F.leaky_relu(ref * d.view(B, Cout, 1, 1) + 0.21 * noise + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    pw = conv.pack_weights(w, scale, up=up)
    kw = dict(in_scale=s, out_scale=d, noise=noise, noise_weight=0.21, bias=bias, act=True, up=up, down=down)
    got = conv.conv2d(x, pw, **kw)
    assert got.shape == ref.shape and rel_err(got, ref) < 1e-2
    if Cout % 8 == 0:
        xcl = x.permute(0, 2, 3, 1).contiguous().half()
        got_cl = conv.conv2d(xcl, pw, out_cl=True, **kw)
        assert rel_err(conv.to_nchw(got_cl), ref) < 1e-2
        # the two layouts run the same MMAs in the same order on operands that agree to fp16 rounding
        assert rel_err(conv.conv2d(xcl, pw, out_cl=False, **kw), ref) < 1e-2


@pytest.mark.parametrize("B,Cin,Cout,H,k", [(2, 64, 128, 24, 1), (1, 512, 512, 16, 3), (1, 128, 12, 40, 1)])
def test_residual_added_in_the_epilogue(B, Cin, Cout, H, k):
    """hav_conv_args.residual: out = act(conv + bias) + residual in one launch (FromRGB / ToRGB skip additions), both layouts, with
    and without cluster split-K."""
    torch.manual_seed(8)
    x = torch.randn(B, Cin, H, H, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda")
    bias = torch.randn(Cout, device="cuda")
    res = torch.randn(B, Cout, H, H, device="cuda")
    scale = 1 / math.sqrt(Cin * k * k)
    ref = F.leaky_relu(F.conv2d(x, w * scale, padding=k // 2) + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2) + res
    pw = conv.pack_weights(w, scale)
    got = conv.conv2d(x, pw, bias=bias, act=True, residual=res)
    assert rel_err(got, ref) < 1e-2
    assert torch.equal(got, conv.conv2d(x, pw, bias=bias, act=True) + res)          # fp32 NCHW: the same addition, bit for bit
    if Cout % 8 == 0:
        xcl, rcl = x.permute(0, 2, 3, 1).contiguous().half(), res.permute(0, 2, 3, 1).contiguous().half()
        got_cl = conv.conv2d(xcl, pw, bias=bias, act=True, out_cl=True, residual=rcl)
        assert rel_err(conv.to_nchw(got_cl), F.leaky_relu(F.conv2d(x, w * scale, padding=k // 2) + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
                       + rcl.permute(0, 3, 1, 2).float()) < 1e-2
    with pytest.raises(Exception):
        conv.conv2d(x, pw, bias=bias, residual=res[:, :, 1:])
