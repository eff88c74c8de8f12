"""Parity of the `model/op` replacements (C ABI: hav_fused_bias_act, hav_upfirdn2d) with the oracle and the
reference-minted goldens, plus first/second-order autograd of the Python wrappers.  pytest -m gpu."""
import os

import numpy as np
import pytest
import torch

from havatar_b200 import op
from havatar_b200.op import fused as fused_mod
from havatar_b200.op import upfirdn2d_op
from oracle import ops_oracle as oo
from oracle.gen_golden import UFD_CASES, ufd_kernel

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_upfirdn2d_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ops.npz"))
    for name, shape, taps, gain, up, down, pad in UFD_CASES:
        x, ref = g["ufd_%s_x" % name], g["ufd_%s_y" % name]
        k = ufd_kernel(np, taps, gain)
        y = op.upfirdn2d(_t(x), _t(k), up=up, down=down, pad=pad).cpu().numpy()
        assert y.shape == ref.shape, name
        assert np.abs(y - ref).max() < 2e-6, name


def test_upfirdn2d_minor_dim_and_large_image():
    rs = np.random.RandomState(0)
    x = rs.standard_normal((3, 20, 33, 4)).astype(np.float32)            # [major,H,W,minor]
    k = ufd_kernel(np, [1, 3, 3, 1], 1.0)
    y = upfirdn2d_op.upfirdn2d(_t(x), _t(k), 2, 2, 1, 1, 2, 1, 2, 1).cpu().numpy()
    ref = oo.upfirdn2d(x.transpose(0, 3, 1, 2), k, 2, 2, 1, 1, 2, 1, 2, 1).transpose(0, 2, 3, 1)
    assert np.abs(y - ref).max() < 2e-6
    x = rs.standard_normal((1, 3, 512, 512)).astype(np.float32)          # full-size blur: linearity + oracle
    y = op.upfirdn2d(_t(x), _t(k), pad=(2, 1)).cpu().numpy()
    assert np.abs(y - oo.upfirdn2d(x, k, 1, 1, 1, 1, 2, 1, 2, 1)).max() < 2e-6


def test_fused_bias_act_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ops.npz"))
    assert np.array_equal(op.fused_leaky_relu(_t(g["act_x"]), _t(g["act_b"])).cpu().numpy(), g["act_y"])
    assert np.array_equal(op.fused_leaky_relu(_t(g["act_x"])).cpu().numpy(), g["act_y_nobias"])
    assert np.array_equal(op.fused_leaky_relu(_t(g["act2_x"]), _t(g["act2_b"])).cpu().numpy(), g["act2_y"])


@pytest.mark.parametrize("shape", [(2, 8, 16, 16), (3, 5, 7, 3), (4, 12), (1, 64, 128, 128)])
def test_fused_bias_act_all_modes(shape):
    rs = np.random.RandomState(1)
    x = rs.standard_normal(shape).astype(np.float32)
    b = rs.standard_normal(shape[1]).astype(np.float32)
    ref = rs.standard_normal(shape).astype(np.float32)
    e = torch.empty(0, device="cuda")
    for act, grad in ((3, 0), (3, 1), (3, 2), (1, 0), (1, 1), (1, 2)):
        got = fused_mod.fused_bias_act(_t(x), _t(b), _t(ref) if grad == 1 else e, act, grad, 0.2, 1.4142135).cpu().numpy()
        want = oo.fused_bias_act(x, b, ref, act, grad, 0.2, 1.4142135)
        assert np.array_equal(got, want), (shape, act, grad)


def test_fused_leaky_relu_first_and_second_order_grads():
    torch.manual_seed(0)
    x = torch.randn(2, 6, 5, 5, device="cuda", requires_grad=True)
    b = torch.randn(6, device="cuda", requires_grad=True)
    y = op.fused_leaky_relu(x, b)
    yr = torch.nn.functional.leaky_relu(x + b.view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    go = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, (x, b), go, create_graph=True)
    gxr, gbr = torch.autograd.grad(yr, (x, b), go, create_graph=True)
    assert torch.allclose(gx, gxr, atol=1e-6) and torch.allclose(gb, gbr, atol=1e-5)
    # R1-style double backward: d/d(go-like input) of |grad|^2
    go2 = torch.randn_like(y, requires_grad=True)
    (gx2,) = torch.autograd.grad(op.fused_leaky_relu(x, b), x, go2, create_graph=True)
    (gx2r,) = torch.autograd.grad(torch.nn.functional.leaky_relu(x + b.view(1, -1, 1, 1), 0.2) * 2 ** 0.5, x, go2,
                                  create_graph=True)
    (h,) = torch.autograd.grad(gx2.pow(2).sum(), go2)
    (hr,) = torch.autograd.grad(gx2r.pow(2).sum(), go2)
    assert torch.allclose(h, hr, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 6, 5, 5), (4, 512, 16, 16), (1, 64, 257, 257), (3, 33, 7), (5, 512), (2, 128, 64, 64), (1, 12, 1, 1)])
def test_fused_leaky_relu_backward_one_pass(shape):
    """hav_bias_act_backward (gradient + per-channel partial sums in one pass) against torch autograd: vector and scalar paths,
    splits over batch boundaries, 2-D inputs (the style MLP), bit-identical grad_input vs the two-launch form."""
    from havatar_b200.op import fused

    torch.manual_seed(1)
    x = torch.randn(*shape, device="cuda", requires_grad=True)
    b = torch.randn(shape[1], device="cuda", requires_grad=True)
    y = op.fused_leaky_relu(x, b)
    bshape = [1, -1] + [1] * (len(shape) - 2)
    yr = torch.nn.functional.leaky_relu(x + b.view(*bshape), 0.2) * 2 ** 0.5
    go = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, (x, b), go)
    gxr, gbr = torch.autograd.grad(yr, (x, b), go)
    assert torch.allclose(gx, gxr, atol=1e-6)
    n = x.numel() // shape[1]
    assert torch.allclose(gb, gbr, atol=2e-6 * max(1.0, n ** 0.5), rtol=1e-5)
    two_launch = fused.fused_bias_act(go.contiguous(), go.new_empty(0), y.detach(), 3, 1, 0.2, 2 ** 0.5)
    assert torch.equal(gx, two_launch)


@pytest.mark.parametrize("shape,per_sample", [((2, 6, 5, 5), False), ((4, 512, 16, 16), False), ((2, 64, 33, 33), True), ((1, 128, 64, 64), False)])
def test_noise_leaky_relu_matches_unfused_formulation(shape, per_sample):
    """StyledConv's tail in one launch (hav_noise_bias_act) and its one-pass backward (grad_x, grad_bias, grad of the noise weight)
    against the reference formulation fused_leaky_relu(x + weight * noise, bias) differentiated by torch (styleUnet.py:300-310, 596-598)."""
    from havatar_b200.op.fused_act import noise_leaky_relu

    torch.manual_seed(2)
    x = torch.randn(*shape, device="cuda", requires_grad=True)
    b = torch.randn(shape[1], device="cuda", requires_grad=True)
    w = torch.full((1,), 0.37, device="cuda", requires_grad=True)
    noise = torch.randn(shape[0] if per_sample else 1, 1, shape[2], shape[3], device="cuda")
    y = noise_leaky_relu(x, noise, w, b)
    yr = torch.nn.functional.leaky_relu(x + w * noise + b.view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    assert torch.allclose(y, yr, atol=1e-6, rtol=1e-6)
    go = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), go)
    gxr, gwr, gbr = torch.autograd.grad(yr, (x, w, b), go)
    n = x.numel() // shape[1]
    assert torch.allclose(gx, gxr, atol=1e-6)
    assert torch.allclose(gb, gbr, atol=2e-6 * max(1.0, n ** 0.5), rtol=1e-5)
    assert torch.allclose(gw, gwr, atol=2e-6 * max(1.0, x.numel() ** 0.5), rtol=1e-4)


def test_upfirdn2d_first_and_second_order_grads():
    torch.manual_seed(0)
    k = _t(ufd_kernel(np, [1, 3, 3, 1], 1.0))
    for up, down, pad in ((1, 1, (2, 1)), (2, 1, (2, 1)), (1, 2, (1, 1))):
        x = torch.randn(2, 3, 8, 10, device="cuda", requires_grad=True)
        y = op.upfirdn2d(x, k, up=up, down=down, pad=pad)
        go = torch.randn_like(y, requires_grad=True)
        (gx,) = torch.autograd.grad(y, x, go, create_graph=True)
        # adjoint identity <A x, g> == <x, A^T g>
        assert abs(float((y * go).sum() - (x * gx).sum())) < 1e-3
        # second order: gradient of <A^T go, v> w.r.t. go is A v
        v = torch.randn_like(x)
        (ggo,) = torch.autograd.grad((gx * v).sum(), go)
        assert torch.allclose(ggo, op.upfirdn2d(v, k, up=up, down=down, pad=pad), atol=1e-5)


def test_reference_module_names_are_installable():
    import sys

    op.install_reference_modules()
    assert sys.modules["fused"].fused_bias_act is fused_mod.fused_bias_act
    assert sys.modules["upfirdn2d"].upfirdn2d is upfirdn2d_op.upfirdn2d


@pytest.mark.parametrize("shape,pad", [((3, 2, 33, 70), (1, 2, 2, 1)), ((2, 3, 257, 257), (1, 1, 1, 1)), ((1, 1, 5, 3), (2, 2, 2, 2)),
                                       ((1, 2, 64, 1025), (2, 1, 2, 1)), ((70000, 1, 4, 4), (1, 1, 1, 1))])
@pytest.mark.parametrize("taps", ["blur", "random"])
def test_streaming_4x4_fir_matches_oracle(shape, pad, taps):
    """up == down == 1, 4x4 taps -> blur4x4_stream_kernel: separable path ([1,3,3,1] outer product) and general path (random
    taps), odd row pitches (every row has a different 16-byte alignment), edge / narrow / > 65535-image cases; and the round-1
    shared-memory kernel must agree with it."""
    rs = np.random.RandomState(sum(shape))
    if shape[0] > 1000:
        x = rs.standard_normal((shape[0], 1, 1, 1)).astype(np.float32) * np.ones(shape, np.float32) + rs.standard_normal(shape[1:]).astype(np.float32)
    else:
        x = rs.standard_normal(shape).astype(np.float32)
    k = ufd_kernel(np, [1, 3, 3, 1], 4.0) if taps == "blur" else rs.standard_normal((4, 4)).astype(np.float32)
    y = op.upfirdn2d(_t(x), _t(k), pad=pad).cpu().numpy()
    n = min(shape[0], 64)
    ref = oo.upfirdn2d(x[:n], k, 1, 1, 1, 1, *pad)
    assert y.shape[1:] == ref.shape[1:] and y.shape[0] == shape[0]
    assert np.abs(y[:n] - ref).max() < 4e-6 * max(1.0, float(np.abs(ref).max()))
    if shape[0] > 1000:
        ref_last = oo.upfirdn2d(x[-2:], k, 1, 1, 1, 1, *pad)
        assert np.abs(y[-2:] - ref_last).max() < 4e-6 * max(1.0, float(np.abs(ref_last).max()))
