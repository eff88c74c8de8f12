"""HD frame timing: plane generators + full-frame render + StyleUNet upsampler.  python scripts/time_hd.py [render_size out_size]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from havatar_b200 import pipeline, synth
rs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
out = int(sys.argv[2]) if len(sys.argv) > 2 else 512
torch.manual_seed(0)
sc = synth.scene(batch=1, height=rs, width=rs, seed=0)
net = pipeline.AvatarHD(sc["weights"], sc["wvol"], render_size=rs, out_size=out).cuda()
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
args = (dev(sc["ray_batch"]), dev(sc["background_prior"]), torch.zeros(1, 32, device="cuda"), dev(sc["inv_head_T"]),
        torch.rand(1, 7, 256, 256, device="cuda"), torch.rand(1, 7, 256, 256, device="cuda"), torch.rand(1, 7, 256, 256, device="cuda"),
        torch.randn(1, 64, device="cuda"))
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
img, lr = net.frame(*args)
print("image", tuple(img.shape), "finite", bool(torch.isfinite(img).all()), "lowres", tuple(lr.shape))
print("eager frame: %.3f ms" % timeit(lambda: net.frame(*args)))
print("  planes only: %.3f ms" % timeit(lambda: net.planes(args[2], args[3], args[4], args[5], args[6])))
g = net.graphed(*args)
ms = timeit(lambda: g(*args))
print("graph frame (%d^2 render -> %d^2): %.3f ms = %.1f HD frames/s" % (rs, out, ms, 1e3 / ms))
