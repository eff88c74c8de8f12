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
# round 2: CTA-pair kernel (16-bit and split precision; 4 ray blocks = an odd share per cluster -> dummy blocks), in-kernel ray
# generation with a pixel selection, the range report, sample_pdf alone, the condition-image kernel, the streaming blur
args = (dev(sc["background_prior"]), dev(sc["inv_head_T"]), dev(sc["planes"]), dev(sc["wvol"]), w, 16, 6)
rk = dict(t_rand=dev(rnd["t_rand"]), noise_coarse=dev(rnd["noise_coarse"]), u_rand=dev(rnd["u_rand"]), noise_fine=dev(rnd["noise_fine"]))
for kw in (dict(precision="fp16", cta_pairs=True), dict(precision="bf16", cta_pairs=True), dict(precision="fp16x3"),
           dict(precision="fp16", check_range=True)):
    o = render.render_rays(dev(sc["ray_batch"]), *args, want_pdf_inds=True, **kw, **rk)
    torch.cuda.synchronize()
    print(kw, float(o.acc_fine.mean()), int(o.pdf_inds.max()))
cam = render.camera_block(np.array([700., 690., 0.5, 0.5], np.float32), np.stack([np.eye(4)[:3]] * 2).astype(np.float32), 2.4, 5.0)
pix = torch.randint(0, 48 * 40, (2, 200), device="cuda", dtype=torch.int32)
o = render.render_rays(None, *args, precision="fp16", camera=cam, img_hw=(48, 40), pixel_index=pix)
torch.cuda.synchronize()
print("camera", float(o.acc_coarse.mean()))
smp, inds = render.sample_pdf(torch.rand(37, 63, device="cuda").sort(-1).values, torch.rand(37, 62, device="cuda"), 16, torch.rand(37, 16, device="cuda"))
from havatar_b200 import data as hdata
c = hdata.make_render_cond(torch.randint(0, 256, (3, 24, 24, 3), dtype=torch.uint8), torch.randint(0, 256, (3, 24, 24, 3), dtype=torch.uint8))
torch.cuda.synchronize()
print("pdf / cond", tuple(smp.shape), tuple(c.shape))
kb = torch.tensor([1., 3., 3., 1.], device="cuda"); kb = kb[None] * kb[:, None]
for shp, pad in (((2, 3, 33, 70), (1, 2, 2, 1)), ((1, 2, 5, 3), (2, 2)), ((3, 1, 257, 257), (1, 1))):
    for kk in (kb, torch.randn(4, 4, device="cuda")):
        y = op.upfirdn2d(torch.randn(*shp, device="cuda"), kk, pad=pad)
torch.cuda.synchronize()
print("blur ok")
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
# end of round 2: TMA-tiled channels-last blur (ragged tiles, out-of-bounds fill), cluster split-K convolutions (uneven split,
# transposed form, channels-last), the style plan, the one-pass activation backward and the fused noise tail
xcl = torch.randn(2, 19, 35, 128, device="cuda").half()
for pad in ((1, 1), (2, 2), (2, 1)):
    for kk in (k, torch.randn(4, 4, device="cuda")):
        y = conv.upfirdn2d_cl(xcl, kk, pad=pad, noise=torch.randn(2, 1, 19 + sum(pad) - 3, 35 + sum(pad) - 3, device="cuda"),
                              noise_weight=0.2, bias=torch.randn(128, device="cuda"), act=True)
torch.cuda.synchronize()
print("blur cl tma", tuple(y.shape), float(y.float().abs().mean()))
for (B, cin, cout, H, ks, up, down) in ((1, 512, 512, 16, 3, 1, 1), (2, 320, 128, 8, 3, 1, 1), (1, 512, 512, 8, 3, 2, 1), (1, 1024, 512, 17, 3, 1, 2),
                                        (1, 512, 12, 32, 1, 1, 1)):
    xx = torch.randn(B, H, H, cin, device="cuda").half()
    pw = conv.pack_weights(torch.randn(cout, cin, ks, ks, device="cuda"), 1 / math.sqrt(cin * ks * ks), up=up)
    y = conv.conv2d(xx, pw, in_scale=torch.rand(B, cin, device="cuda"), out_scale=torch.rand(B, cout, device="cuda"),
                    bias=torch.randn(cout, device="cuda"), act=True, up=up, down=down, out_cl=cout % 8 == 0)
    torch.cuda.synchronize()
    print("conv split-K", cin, cout, H, up, down, tuple(y.shape), float(y.float().abs().mean()))
from havatar_b200 import styleunet as su
net = su.SWGAN_unet(inp_size=32, inp_ch=16, out_ch=3, out_size=64, style_dim=64, n_mlp=2, middle_size=8).cuda().eval()
with torch.no_grad():
    img = net([torch.randn(2, 64, device="cuda")], torch.randn(2, 16, 32, 32, device="cuda"), noise=net.make_noise("cuda"))
torch.cuda.synchronize()
print("style plan + unet", tuple(img.shape), float(img.abs().mean()))
from havatar_b200.op.fused_act import noise_leaky_relu
for shp in ((2, 6, 5, 5), (3, 40, 16, 16)):
    xg = torch.randn(*shp, device="cuda", requires_grad=True)
    bg = torch.randn(shp[1], device="cuda", requires_grad=True)
    wg = torch.full((1,), 0.3, device="cuda", requires_grad=True)
    (op.fused_leaky_relu(xg, bg).sum() + noise_leaky_relu(xg, torch.randn(1, 1, shp[2], shp[3], device="cuda"), wg, bg).sum()).backward()
torch.cuda.synchronize()
print("act backward ok", float(bg.grad.abs().mean()), float(wg.grad))
