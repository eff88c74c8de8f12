"""Micro-benchmark of conv.conv2d for a few shapes (CUDA events)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from havatar_b200 import conv
def bench(B, Cin, Cout, H, k, up=1, down=1, n=20):
    x = torch.randn(B, Cin, H, H, device="cuda"); w = torch.randn(Cout, Cin, k, k, device="cuda")
    pw = conv.pack_weights(w, 1 / math.sqrt(Cin * k * k), up=up)
    for _ in range(3): y = conv.conv2d(x, pw, up=up, down=down)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): y = conv.conv2d(x, pw, up=up, down=down)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * B * y.shape[2] * y.shape[3] * Cin * Cout * k * k
    kb = (Cin + 63) // 64
    print("B%d Cin%4d Cout%4d H%4d k%d up%d down%d: %8.1f us  %6.1f TFLOP/s(out-res count)  %.2f us/kblock %.2f us/(kblock*tap)" % (
        B, Cin, Cout, H, k, up, down, us, fl / us / 1e6, us / kb, us / kb / (k * k)))
for args in [(1, 512, 512, 16, 3), (1, 512, 512, 16, 1), (1, 64, 512, 16, 3), (1, 64, 128, 16, 3), (1, 64, 16, 16, 3), (1, 1024, 512, 16, 3),
             (1, 128, 128, 256, 3), (1, 256, 128, 128, 3, 2), (4, 512, 512, 64, 3)]:
    bench(*args)
