"""One channels-last blur launch on the largest StyleUNet shape for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from havatar_b200 import conv  # noqa: E402

k = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
k = k[None, :] * k[:, None]
k = k / k.sum()
x = torch.randn(1, 513, 513, 64, device="cuda").half()
for _ in range(4):
    y = conv.upfirdn2d_cl(x, k, pad=(1, 1))
torch.cuda.synchronize()
