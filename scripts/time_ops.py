"""HBM roofline of the two `model/op` replacements: achieved GB/s on algorithmic bytes (4 B per element read + written)
against MEASURED_PEAKS.json hbm_gbs.  Shapes are the large StyleUNet layers."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from havatar_b200 import op  # noqa: E402

peak = 6650.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


blur = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
blur = (blur[None, :] * blur[:, None])
blur = blur / blur.sum()
haar = torch.tensor([[1.0, 1.0], [1.0, 1.0]], device="cuda") / 2
with torch.no_grad():
    for name, shape in (("fused_leaky_relu", (4, 256, 256, 256)), ("fused_leaky_relu", (4, 512, 64, 64))):
        x, bias = torch.randn(shape, device="cuda"), torch.randn(shape[1], device="cuda")
        ms = timeit(lambda: op.fused_leaky_relu(x, bias))
        by = 8.0 * x.numel()
        print("%-34s %-20s %8.1f us  %7.0f GB/s  %.2f of peak" % (name, shape, ms * 1e3, by / ms / 1e6, by / ms / 1e6 / peak))
    for name, shape, k, up, down, pad in (("upfirdn2d blur 4x4 (after convT)", (4, 256, 257, 257), blur * 4, 1, 1, (1, 1)),
                                          ("upfirdn2d blur 4x4 (before down)", (4, 256, 256, 256), blur, 1, 1, (2, 2)),
                                          ("upfirdn2d Upsample x2", (4, 12, 256, 256), blur * 4, 2, 1, (2, 1)),
                                          ("upfirdn2d Haar down 2x2", (4, 12, 512, 512), haar, 1, 2, (0, 0))):
        x = torch.randn(shape, device="cuda")
        y = op.upfirdn2d(x, k, up=up, down=down, pad=pad)
        ms = timeit(lambda: op.upfirdn2d(x, k, up=up, down=down, pad=pad))
        by = 4.0 * (x.numel() + y.numel())
        print("%-34s %-20s %8.1f us  %7.0f GB/s  %.2f of peak" % (name, shape, ms * 1e3, by / ms / 1e6, by / ms / 1e6 / peak))
