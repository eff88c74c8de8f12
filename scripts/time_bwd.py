"""Timing of the fused render forward + backward at the stage-one training shape (BASELINE configs[2]: batch of 4 frames x one
64x64 patch, 64 + 16 hierarchical samples, perturb + noise) next to torch-CUDA autograd over the reference's ATen sequence."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from havatar_b200 import render, synth  # noqa: E402
from oracle import render_oracle as ro  # noqa: E402
from oracle import render_oracle_torch as rot  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sc = synth.scene(batch=B, crop=(224, 224, P, P), seed=60)
R = P * P
r = synth.randoms(B, R, 64, 16, seed=67)
rnd = {k: torch.from_numpy(r[k]).cuda() for k in ("t_rand", "noise_coarse", "u_rand", "noise_fine")}
cot = synth.cotangents(B, R, True, seed=71, scale=1.0 / (B * R))
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
w = {k: dev(v) for k, v in sc["weights"].items()}
inp = [dev(sc[k]) for k in ("ray_batch", "background_prior", "inv_head_T", "planes", "wvol")]
g = {k: dev(v) for k, v in cot.items()}


def step():
    out, ctx = render.render_rays(*inp, w, 64, 16, precision="fp16", want_z_fine=True, return_ctx=True, **rnd)
    return ctx


def bwd(ctx):
    return render.render_backward(ctx, g_rgb_coarse=g["rgb_coarse"], g_depth_coarse=g["depth_coarse"], g_acc_coarse=g["acc_coarse"],
                                  g_rgb_fine=g["rgb_fine"], g_depth_fine=g["depth_fine"], g_acc_fine=g["acc_fine"])


for _ in range(3):
    ctx = step()
    bwd(ctx)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
n = 10
tf = tb = 0.0
for _ in range(n):
    ev[0].record()
    ctx = step()
    ev[1].record()
    bwd(ctx)
    ev[2].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1])
    tb += ev[1].elapsed_time(ev[2])
samples = B * R * (64 + 48)
print("fused  B=%d R=%d: forward %.3f ms  backward %.3f ms  (%.1f M samples/s fwd+bwd)" % (B, R, tf / n, tb / n, samples / (tf + tb) * n / 1e3))

# the reference's formulation on the same GPU: ATen ops + autograd, fp32 (4096-ray chunks)
for it in range(3):
    torch.cuda.synchronize()
    ev[0].record()
    rot.render_rays_grad(sc["ray_batch"], sc["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"], sc["weights"],
                         ro.default_boxes(), 64, 16, cotangents=cot, device="cuda", **{k: r[k] for k in rnd})
    ev[1].record()
    torch.cuda.synchronize()
    t_ref = ev[0].elapsed_time(ev[1])
print("ATen autograd port (fp32, incl. H2D of inputs): %.1f ms fwd+bwd -> fused is %.1fx" % (t_ref, t_ref / ((tf + tb) / n)))
