"""Micro-benchmark of conv.conv2d for a few shapes (CUDA events)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from havatar_b200 import conv
CL = len(sys.argv) > 1 and sys.argv[1] == "cl"
def bench(B, Cin, Cout, H, k, up=1, down=1, n=20):
    x = torch.randn(B, Cin, H, H, device="cuda"); w = torch.randn(Cout, Cin, k, k, device="cuda")
    if CL: x = x.permute(0, 2, 3, 1).contiguous().half()
    pw = conv.pack_weights(w, 1 / math.sqrt(Cin * k * k), up=up)
    for _ in range(3): y = conv.conv2d(x, pw, up=up, down=down, out_cl=CL)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): y = conv.conv2d(x, pw, up=up, down=down, out_cl=CL)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    hw = (y.shape[1] * y.shape[2]) if CL else (y.shape[2] * y.shape[3])
    fl = 2.0 * B * hw * Cin * Cout * k * k
    kb = (Cin + 63) // 64
    print("B%d Cin%4d Cout%4d H%4d k%d up%d down%d: %8.1f us  %6.1f TFLOP/s(out-res count)  %.2f us/kblock %.2f us/(kblock*tap)" % (
        B, Cin, Cout, H, k, up, down, us, fl / us / 1e6, us / kb, us / kb / (k * k)))
for args in [(1, 512, 512, 16, 3), (1, 512, 512, 16, 1), (1, 64, 512, 16, 3), (1, 64, 128, 16, 3), (1, 64, 16, 16, 3), (1, 1024, 512, 16, 3),
             (1, 128, 128, 256, 3), (1, 256, 128, 128, 3, 2), (4, 512, 512, 64, 3)]:
    bench(*args)
