import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from havatar_b200 import conv
x = torch.randn(4, 512, 64, 64, device="cuda"); w = torch.randn(512, 512, 3, 3, device="cuda")
pw = conv.pack_weights(w, 1 / math.sqrt(512 * 9))
for _ in range(4): y = conv.conv2d(x, pw)
torch.cuda.synchronize()
