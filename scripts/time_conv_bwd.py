"""Forward + backward time of the StyleUNet convolution shapes: conv.conv2d_autograd (tcgen05 forward / data-gradient / weight-
gradient kernels) next to torch autograd over cuDNN (TF32 allowed, its default) for the same formula."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from havatar_b200 import conv  # noqa: E402

SHAPES = [  # B, Cin, Cout, H, k, up, down
    (4, 512, 512, 16, 3, 1, 1), (4, 512, 512, 32, 3, 1, 1), (4, 512, 512, 64, 3, 1, 1), (4, 256, 256, 128, 3, 1, 1),
    (4, 512, 512, 32, 3, 2, 1), (4, 512, 256, 64, 3, 2, 1), (4, 256, 512, 131, 3, 1, 2), (1, 512, 512, 64, 3, 1, 1),
    (1, 256, 256, 128, 3, 1, 1), (1, 128, 128, 256, 3, 1, 1), (1, 128, 12, 256, 1, 1, 1), (1, 64, 64, 515, 3, 1, 2),
]


def ref(x, w, s, d, ws, up, down):
    xs = x * s[:, :, None, None]
    if up == 2:
        y = F.conv_transpose2d(xs, (w * ws).transpose(0, 1), stride=2)
    elif down == 2:
        y = F.conv2d(xs, w * ws, stride=2)
    else:
        y = F.conv2d(xs, w * ws, padding=w.shape[-1] // 2)
    return y * d[:, :, None, None]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for B, Cin, Cout, H, k, up, down in SHAPES:
    x = torch.randn(B, Cin, H, H, device="cuda", requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, device="cuda", requires_grad=True)
    s = torch.rand(B, Cin, device="cuda").add_(0.5).requires_grad_(True)
    d = torch.rand(B, Cout, device="cuda").add_(0.5).requires_grad_(True)
    ws = 1.0 / (Cin * k * k) ** 0.5
    y0 = conv.conv2d_autograd(x, w, s, d, ws, up=up, down=down)
    go = torch.randn_like(y0)

    def ours():
        y = conv.conv2d_autograd(x, w, s, d, ws, up=up, down=down)
        torch.autograd.grad(y, (x, w, s, d), go)

    def ours_fwd():
        with torch.no_grad():
            conv.conv2d_autograd(x, w, s, d, ws, up=up, down=down)

    def lib():
        y = ref(x, w, s, d, ws, up, down)
        torch.autograd.grad(y, (x, w, s, d), go)

    def wgrad_only():
        conv.conv_wgrad(go, x.detach(), k, in_scale=s.detach(), out_scale=d.detach(), wscale=ws, up=up, down=down)

    flop = 2.0 * B * y0.shape[2] * y0.shape[3] * Cin * Cout * k * k / (4 if up == 2 else 1)
    t_o, t_f, t_l, t_w = timeit(ours), timeit(ours_fwd), timeit(lib), timeit(wgrad_only)
    print("B%d %4d->%4d %3dx%-3d k%d up%d down%d | ours fwd+bwd %7.3f ms (fwd %6.3f, wgrad %6.3f = %6.1f TFLOP/s) | cuDNN autograd %7.3f ms | x%.2f"
          % (B, Cin, Cout, H, H, k, up, down, t_o, t_f, t_w, flop / t_w / 1e9, t_l, t_l / t_o))
