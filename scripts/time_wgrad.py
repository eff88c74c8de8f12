"""Time hav_conv2d_wgrad alone on one shape (HAV_WG_DEBUG=1: MMA side only, operand staging skipped after the first ring fill)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from havatar_b200 import conv  # noqa: E402

for B, Cin, Cout, H in ((4, 512, 512, 64), (4, 256, 256, 128), (1, 512, 512, 64)):
    x, g = torch.randn(B, Cin, H, H, device="cuda"), torch.randn(B, Cout, H, H, device="cuda")
    s, d = torch.rand(B, Cin, device="cuda") + 0.5, torch.rand(B, Cout, device="cuda") + 0.5
    out = torch.zeros(Cout, Cin, 3, 3, device="cuda")
    f = lambda: conv.conv_wgrad(g, x, 3, in_scale=s, out_scale=d, wscale=0.1, out=out)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        f()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print("B%d %d->%d %dx%d: %.1f us  %.0f TFLOP/s" % (B, Cin, Cout, H, H, ms * 1e3, 2.0 * B * H * H * Cin * Cout * 9 / ms / 1e9))
