"""Scratch timing of render_rays on the GPU box (CUDA events).  python scripts/time_render.py [prec] [H] [S] [nfine]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from havatar_b200 import render, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
nf = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sc = synth.scene(batch=1, height=H, width=H, seed=0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
args = [dev(sc[k]) for k in ("ray_batch", "background_prior", "inv_head_T", "planes", "wvol")]
w = {k: dev(v) for k, v in sc["weights"].items()}
for _ in range(3):
    out = render.render_rays(*args, w, S, nf, precision=prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    out = render.render_rays(*args, w, S, nf, precision=prec)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
R = H * H
print("prec=%s %dx%d S=%d nf=%d: %.3f ms/frame  %.2f Mrays/s  %.1f TFLOP/s  acc=%.3f" % (
    prec, H, H, S, nf, ms, R / ms / 1e3, R * S * 94848 / ms / 1e9, float(out.acc_coarse.mean())))
