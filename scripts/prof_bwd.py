"""One forward + backward at the training shape, for ncu (python scripts/prof_bwd.py [B] [patch])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from havatar_b200 import render, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc = synth.scene(batch=B, crop=(224, 224, P, P), seed=60)
R = P * P
r = synth.randoms(B, R, 64, 16, seed=67)
rnd = {k: torch.from_numpy(r[k]).cuda() for k in ("t_rand", "noise_coarse", "u_rand", "noise_fine")}
cot = synth.cotangents(B, R, True, seed=71, scale=1.0 / (B * R))
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
w = {k: dev(v) for k, v in sc["weights"].items()}
inp = [dev(sc[k]) for k in ("ray_batch", "background_prior", "inv_head_T", "planes", "wvol")]
g = {k: dev(v) for k, v in cot.items()}
for _ in range(reps):
    out, ctx = render.render_rays(*inp, w, 64, 16, precision="fp16", want_z_fine=True, return_ctx=True, **rnd)
    render.render_backward(ctx, g_rgb_coarse=g["rgb_coarse"], g_depth_coarse=g["depth_coarse"], g_acc_coarse=g["acc_coarse"],
                           g_rgb_fine=g["rgb_fine"], g_depth_fine=g["depth_fine"], g_acc_fine=g["acc_fine"])
torch.cuda.synchronize()
print("done")
