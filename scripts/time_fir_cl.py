"""Channels-last 4x4 blur (hav_upfirdn2d_cl) on the StyleUNet's shapes: time (10x CUDA-graph replay) and algorithmic GB/s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from havatar_b200 import conv  # noqa: E402

k = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
k = k[None, :] * k[:, None]
k = k / k.sum()
for B, C, H, pad in [(1, 64, 513, (1, 1)), (1, 128, 257, (1, 1)), (1, 128, 256, (2, 2)), (1, 256, 129, (1, 1)), (1, 256, 128, (2, 2)),
                     (1, 512, 65, (1, 1)), (1, 512, 64, (2, 2)), (1, 512, 33, (1, 1)), (1, 512, 17, (1, 1)), (4, 64, 513, (1, 1)),
                     (4, 256, 129, (1, 1)), (4, 512, 65, (1, 1))]:
    x = torch.randn(B, H, H, C, device="cuda").half()
    noise = torch.randn(1, 1, H + pad[0] + pad[1] - 3, H + pad[0] + pad[1] - 3, device="cuda")
    bias = torch.randn(C, device="cuda")
    f = lambda: conv.upfirdn2d_cl(x, k, pad=pad, noise=noise, noise_weight=0.1, bias=bias, act=True)  # noqa: E731
    y = f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 100)
    by = (x.numel() + y.numel()) * 2
    print("B%d C%4d %4dx%-4d -> %4d: %7.1f us  %7.1f GB/s" % (B, C, H, H, y.shape[1], best, by / best / 1e3))
