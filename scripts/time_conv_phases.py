"""Phase stamps of CTA 0 of conv_tc_kernel (debug build: HAV_NVCC_DEFS=-DHAV_CONV_TIMING python -m havatar_b200.build --force)."""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from havatar_b200 import _lib, conv  # noqa: E402

names = ["start", "alloc+sync", "scale+bar", "staged kb0", "weights step0", "mma issued", "staged last", "acc ready", "epilogue", "cluster done", "dealloc"]
L = _lib.lib()
for B, Cin, Cout, H, k, up, down in [(1, 64, 512, 16, 1, 1, 1), (1, 512, 512, 16, 3, 1, 1), (1, 512, 512, 64, 3, 1, 1), (1, 64, 64, 512, 3, 1, 1)]:
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    pw = conv.pack_weights(torch.randn(Cout, Cin, k, k, device="cuda"), 1 / math.sqrt(Cin * k * k), up=up)
    bias = torch.randn(Cout, device="cuda")
    for _ in range(3):
        y = conv.conv2d(x, pw, bias=bias, act=True, up=up, down=down, out_cl=True)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    L.hav_conv_debug_stamps.restype = C.c_int
    assert L.hav_conv_debug_stamps(buf) == 0
    t = list(buf)
    print("Cin%d Cout%d %dx%d k%d up%d down%d:" % (Cin, Cout, H, H, k, up, down),
          "  ".join("%s %+d" % (names[i], t[i] - t[0]) for i in range(1, 11) if t[i]))
