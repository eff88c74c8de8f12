"""Small invocations of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from havatar_b200 import render, synth, conv, op
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
sc = synth.scene(batch=2, crop=(250, 200, 1, 200), seed=1)       # 400 rays: ragged tiles, 2 frames
w = {k: dev(v) for k, v in sc["weights"].items()}
rnd = synth.randoms(2, 200, 16, 6, seed=3)
for prec in ("fp32", "fp16"):
    o = render.render_rays(dev(sc["ray_batch"]), dev(sc["background_prior"]), dev(sc["inv_head_T"]), dev(sc["planes"]), dev(sc["wvol"]), w,
                           16, 6, precision=prec, t_rand=dev(rnd["t_rand"]), noise_coarse=dev(rnd["noise_coarse"]),
                           u_rand=dev(rnd["u_rand"]), noise_fine=dev(rnd["noise_fine"]), want_z_fine=True)
    torch.cuda.synchronize()
    print(prec, float(o.acc_fine.mean()))
x = torch.randn(2, 72, 20, 12, device="cuda")
wt = torch.randn(40, 72, 3, 3, device="cuda")
for up, down in ((1, 1), (2, 1), (1, 2)):
    y = conv.conv2d(x, conv.pack_weights(wt, 0.05, up=up), up=up, down=down, in_scale=torch.rand(2, 72, device="cuda"),
                    out_scale=torch.rand(2, 40, device="cuda"), bias=torch.randn(40, device="cuda"), act=True)
    torch.cuda.synchronize()
    print("conv", up, down, tuple(y.shape), float(y.abs().mean()))
k = torch.tensor([1., 3., 3., 1.], device="cuda"); k = k[None] * k[:, None]; k = k / k.sum()
for up, down, pad in ((1, 1, (2, 1)), (2, 1, (2, 1)), (1, 2, (1, 1)), (3, 2, (1, 2))):
    y = op.upfirdn2d(torch.randn(2, 3, 19, 23, device="cuda"), k, up=up, down=down, pad=pad)
    torch.cuda.synchronize()
    print("ufd", up, down, tuple(y.shape))
y = op.fused_leaky_relu(torch.randn(2, 5, 7, 3, device="cuda"), torch.randn(5, device="cuda"))
rays = render.get_rays(9, 11, (700., 690., 0.5, 0.5), np.eye(4)[:3], 2.4, 5.0)
torch.cuda.synchronize()
print("ok")
