"""One hav_conv2d_wgrad launch per shape for ncu (scripts/ncu_summary.py reads the report)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from havatar_b200 import conv  # noqa: E402

B, Cin, Cout, H, k, up, down = [int(v) for v in (sys.argv[1:8] if len(sys.argv) >= 8 else (4, 512, 512, 64, 3, 1, 1))]
x = torch.randn(B, Cin, H, H, device="cuda")
Ho = 2 * H + 1 if up == 2 else ((H - k) // 2 + 1 if down == 2 else H)
g = torch.randn(B, Cout, Ho, Ho, device="cuda")
s, d = torch.rand(B, Cin, device="cuda") + 0.5, torch.rand(B, Cout, device="cuda") + 0.5
for _ in range(4):
    conv.conv_wgrad(g, x, k, in_scale=s, out_scale=d, wscale=0.1, up=up, down=down)
torch.cuda.synchronize()
