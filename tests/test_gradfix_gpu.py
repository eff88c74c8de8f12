"""model/op/conv2d_gradfix.py drop-in (havatar_b200/op/conv2d_gradfix.py) against torch fp32 convolutions: the call forms
model/styleUnet.py issues, including the per-sample grouped form of the fused ModulatedConv2d branch, first-order gradients,
and the R1-style double backward with no_weight_gradients()."""
import pytest
import torch
import torch.nn.functional as F

from havatar_b200.op import conv2d_gradfix as gf

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("k,stride,groups", [(3, 1, 1), (1, 1, 1), (3, 2, 1), (3, 1, 2), (1, 1, 2)])
def test_conv2d_forms(k, stride, groups):
    torch.manual_seed(0)
    cin, cout, H = 64, 32, 17
    x = torch.randn(1 if groups > 1 else 2, groups * cin, H, H, device="cuda", requires_grad=True)
    w = (torch.randn(groups * cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5).requires_grad_(True)
    b = torch.randn(groups * cout, device="cuda")
    pad = k // 2 if stride == 1 else 0
    ref = F.conv2d(x, w, bias=b, stride=stride, padding=pad, groups=groups)
    got = gf.conv2d(x, w, bias=b, stride=stride, padding=pad, groups=groups)
    assert got.shape == ref.shape and _rel(got, ref) < 1e-2
    cot = torch.randn_like(ref)
    gr = torch.autograd.grad((ref * cot).sum(), (x, w))
    gg = torch.autograd.grad((got * cot).sum(), (x, w))
    assert _rel(gg[0], gr[0]) < 2e-2 and _rel(gg[1], gr[1]) < 2e-2


@pytest.mark.parametrize("groups", [1, 2])
def test_conv_transpose2d_form(groups):
    torch.manual_seed(1)
    cin, cout, H = 64, 32, 9
    x = torch.randn(1 if groups > 1 else 2, groups * cin, H, H, device="cuda", requires_grad=True)
    w = (torch.randn(groups * cin, cout, 3, 3, device="cuda") / (cin * 9) ** 0.5).requires_grad_(True)
    ref = F.conv_transpose2d(x, w, stride=2, padding=0, groups=groups)
    got = gf.conv_transpose2d(x, w, stride=2, padding=0, groups=groups)
    assert got.shape == ref.shape and _rel(got, ref) < 1e-2
    cot = torch.randn_like(ref)
    gr = torch.autograd.grad((ref * cot).sum(), (x, w))
    gg = torch.autograd.grad((got * cot).sum(), (x, w))
    assert _rel(gg[0], gr[0]) < 2e-2 and _rel(gg[1], gr[1]) < 2e-2


def test_r1_double_backward_with_no_weight_gradients():
    """utils/styleUnet_util.py:72-79 on a two-convolution toy discriminator."""
    torch.manual_seed(2)
    x = torch.randn(2, 16, 12, 12, device="cuda")
    w1 = (torch.randn(32, 16, 3, 3, device="cuda") / 12).requires_grad_(True)
    w2 = (torch.randn(8, 32, 3, 3, device="cuda") / 17).requires_grad_(True)

    def penalty(conv, ctx):
        xi = x.clone().requires_grad_(True)
        pred = conv(F.leaky_relu(conv(xi, w1, padding=1), 0.2), w2, stride=2).sum()
        with ctx():
            g, = torch.autograd.grad(pred, xi, create_graph=True)
        return g.pow(2).reshape(2, -1).sum(1).mean()

    import contextlib
    ref = penalty(F.conv2d, contextlib.nullcontext)
    got = penalty(gf.conv2d, gf.no_weight_gradients)
    assert abs(float(got) - float(ref)) < 3e-2 * abs(float(ref))
    gr = torch.autograd.grad(ref, (w1, w2))
    gg = torch.autograd.grad(got, (w1, w2))
    assert _rel(gg[0], gr[0]) < 5e-2 and _rel(gg[1], gr[1]) < 5e-2


def test_unsupported_forms_raise():
    x = torch.randn(1, 8, 8, 8, device="cuda")
    with pytest.raises(NotImplementedError):
        gf.conv2d(x, torch.randn(8, 8, 5, 5, device="cuda"), padding=2)
    with pytest.raises(NotImplementedError):
        gf.conv2d(x, torch.randn(8, 8, 3, 3, device="cuda"), padding=1, dilation=2)
    with pytest.raises(NotImplementedError):
        gf.conv_transpose2d(x, torch.randn(8, 8, 3, 3, device="cuda"), stride=1)
