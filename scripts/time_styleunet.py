"""Scratch timing of the StyleUNet forward on the GPU box.  python scripts/time_styleunet.py [inp out] [batch]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from havatar_b200 import styleunet
inp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
out = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.manual_seed(0)
torch.set_grad_enabled(False)   # inference kernels (grad mode selects the differentiable formulation)
net = styleunet.SWGAN_unet(inp_size=inp, inp_ch=64, out_ch=3, out_size=out, style_dim=64, n_mlp=4, middle_size=8).cuda()
x = torch.randn(B, 64, inp, inp, device="cuda"); s = torch.randn(B, 64, device="cuda")
noise = net.make_noise("cuda")
for _ in range(3):
    y = net([s], x, noise=noise)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    y = net([s], x, noise=noise)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
gf = {(128, 512): 176.3, (512, 1024): 352.3}.get((inp, out), 0) * B
from havatar_b200.graph import GraphedForward
g = GraphedForward(lambda st, c: net([st], c, noise=noise), s, x)
yg = g(s, x); torch.cuda.synchronize()
print("graph vs eager max diff", float((yg - y).abs().max()))
e0.record()
for _ in range(n):
    yg = g(s, x)
e1.record(); torch.cuda.synchronize()
msg = e0.elapsed_time(e1) / n
print("  CUDA graph replay: %.3f ms  %.1f frames/s  %.1f TFLOP/s" % (msg, B * 1e3 / msg, gf / msg))
print("SWGAN_unet %d->%d B=%d: %.3f ms/frame-batch  %.1f frames/s  %.1f TFLOP/s (reference FLOP count)  out %s finite=%s" % (
    inp, out, B, ms, B * 1e3 / ms, gf / ms, tuple(y.shape), bool(torch.isfinite(y).all())))
