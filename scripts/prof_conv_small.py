"""A few small / large channels-last convolution launches for ncu (launch list or --set full)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from havatar_b200 import conv  # noqa: E402

shapes = [(1, 64, 512, 16, 1, 1, 1), (1, 512, 512, 16, 3, 1, 1), (1, 512, 512, 32, 3, 1, 1), (1, 512, 512, 64, 3, 1, 1), (1, 64, 64, 512, 3, 1, 1),
          (1, 512, 512, 16, 3, 2, 1), (1, 256, 512, 129, 3, 1, 2), (1, 512, 256, 64, 3, 2, 1)]
if len(sys.argv) > 1:
    shapes = [shapes[int(sys.argv[1])]]
for B, Cin, Cout, H, k, up, down in shapes:
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    pw = conv.pack_weights(torch.randn(Cout, Cin, k, k, device="cuda"), 1 / math.sqrt(Cin * k * k), up=up)
    bias = torch.randn(Cout, device="cuda")
    for _ in range(3):
        y = conv.conv2d(x, pw, bias=bias, act=True, up=up, down=down, out_cl=True)
    torch.cuda.synchronize()
