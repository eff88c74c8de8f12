"""Mint the golden fixtures under tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden
Inputs are regenerated from seeds by havatar_b200/synth.py, so each fixture stores reference
OUTPUTS only (+ the case parameters).  tests/test_oracle_golden.py checks oracle/ against them;
the gpu tests check the CUDA path against them too.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from havatar_b200 import synth  # noqa: E402

# case name -> parameters.  `crop` = (y0, x0, h, w) window of a 512x512 frame.
RENDER_CASES = {
    # BASELINE.json configs[0] scaled to fixture size: 32x32 crop, 32 samples, coarse only
    "render_c32_s32": dict(batch=1, crop=(240, 240, 32, 32), num_coarse=32, num_fine=0, rand=False, seed=0),
    # hierarchical 64 + 16 (-> 48 fine samples), deterministic (perturb off), B=2
    "render_hier_det": dict(batch=2, crop=(200, 260, 16, 16), num_coarse=64, num_fine=16, rand=False, seed=10),
    # train-mode randoms supplied explicitly: perturb + sigma noise 0.1 + stratified sample_pdf
    "render_hier_rand": dict(batch=1, crop=(300, 180, 16, 16), num_coarse=64, num_fine=16, rand=True, seed=20),
    # non-square planes / non-cubic volume pin the axis conventions; window at the frame corner leaves the box
    "render_oddshape": dict(batch=1, crop=(0, 0, 8, 16), num_coarse=16, num_fine=0, rand=False, seed=30,
                            plane_hw=(24, 40), vol_dhw=(6, 10, 14)),
    # BASELINE.json configs[0] LITERALLY (round 2): single 64 x 64 crop, 32 samples per ray, coarse only
    "render_c64_s32": dict(batch=1, crop=(224, 224, 64, 64), num_coarse=32, num_fine=0, rand=False, seed=60),
}
ROUND2_RENDER_CASES = ("render_c64_s32",)


# backward cases: gradients of  sum_k <cotangent_k, out_k>  from the reference's own autograd (small planes / volume keep the
# fixtures small; every learnable input of predict_and_render_radiance gets a gradient)
RENDER_BWD_CASES = {
    "render_bwd_coarse": dict(batch=1, crop=(248, 240, 7, 15), num_coarse=24, num_fine=0, rand=False, seed=40,
                              plane_hw=(32, 32), vol_dhw=(16, 16, 16)),
    "render_bwd_hier_rand": dict(batch=2, crop=(230, 250, 8, 16), num_coarse=64, num_fine=16, rand=True, seed=50,
                                 plane_hw=(32, 48), vol_dhw=(12, 16, 20)),
}


def _reference_render(trainer, torch, case, grad=False, inds_out=None):
    sc = synth.scene(batch=case["batch"], crop=case["crop"], seed=case["seed"],
                     plane_hw=case.get("plane_hw", (128, 128)), vol_dhw=case.get("vol_dhw", (64, 64, 64)))
    B, R = sc["ray_batch"].shape[:2]
    mc = trainer.model_coarse
    sd = {k: torch.from_numpy(v) for k, v in sc["weights"].items()}
    mc.load_state_dict(sd, strict=False)
    mc.triPlane_embeddings = torch.from_numpy(sc["planes"]).requires_grad_(grad)
    hs = trainer.headpose_skin_net
    hs.fix_canoW = True
    hs.canonical_W = torch.from_numpy(sc["wvol"]).requires_grad_(grad)
    opt = trainer.cfg.nerf.train
    opt.num_coarse, opt.num_fine = case["num_coarse"], case["num_fine"]
    opt.perturb = bool(case["rand"])
    opt.radiance_field_noise_std = 0.1 if case["rand"] else 0.0

    rays = torch.from_numpy(sc["ray_batch"])
    viewdirs = rays[..., 3:6] / rays[..., 3:6].norm(p=2, dim=-1).unsqueeze(-1)   # nerf_trainer.py:52
    rays11 = torch.cat((rays, viewdirs), dim=-1)                                  # nerf_trainer.py:63

    orig_rand, orig_randn = torch.rand, torch.randn
    if case["rand"]:
        rnd = synth.randoms(B, R, case["num_coarse"], case["num_fine"], seed=case["seed"] + 7)
        q_rand = [rnd["t_rand"], rnd["u_rand"].reshape(B * R, -1)]
        q_randn = [rnd["unit_coarse"].reshape(B * R, -1), rnd["unit_fine"].reshape(B * R, -1)]

        def fake_rand(shape, *a, **k):
            v = q_rand.pop(0)
            assert tuple(shape) == v.shape, (tuple(shape), v.shape)
            return torch.from_numpy(v.copy())

        def fake_randn(shape, *a, **k):
            v = q_randn.pop(0)
            assert tuple(shape) == v.shape, (tuple(shape), v.shape)
            return torch.from_numpy(v.copy())

        torch.rand, torch.randn = fake_rand, fake_randn
    orig_ss = torch.searchsorted
    if inds_out is not None:      # integer bookkeeping of sample_pdf (utils/nerf_util.py:102), recorded as the reference computes it
        def spy(*a, **k):
            r = orig_ss(*a, **k)
            inds_out.append(r.detach().numpy().copy())
            return r
        torch.searchsorted = spy
    try:
        with torch.set_grad_enabled(grad):
            out = trainer.predict_and_render_radiance("train", rays11, torch.from_numpy(sc["background_prior"]),
                                                      inv_head_T=torch.from_numpy(sc["inv_head_T"]))
    finally:
        torch.rand, torch.randn, torch.searchsorted = orig_rand, orig_randn, orig_ss
    names = ["rgb_coarse", "depth_coarse", "acc_coarse", "weights_max", "rgb_fine", "depth_fine", "acc_fine"]
    res = {n: o.detach().numpy() for n, o in zip(names, out) if o is not None}
    if not grad:
        return res
    outs = dict(zip(names, out))
    cot = synth.cotangents(B, R, case["num_fine"] > 0, seed=case["seed"] + 11)
    loss = 0.0
    for k, c in cot.items():
        loss = loss + (outs[k] * torch.from_numpy(c).reshape(outs[k].shape)).sum()
    trainer.zero_grad()
    loss.backward()
    res = {"out_" + k: v for k, v in res.items()}
    res["g_planes"] = mc.triPlane_embeddings.grad.numpy()
    res["g_wvol"] = hs.canonical_W.grad.numpy()
    params = dict(mc.named_parameters())
    for k in sc["weights"]:
        res["g_" + k] = params[k].grad.numpy().copy()
    return res


def gen_render_bwd(trainer, torch):
    for name, case in RENDER_BWD_CASES.items():
        out = _reference_render(trainer, torch, case, grad=True)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), case=json.dumps(case), **out)
        print(name, {k: (v.shape, "%.3g" % np.abs(v).max()) for k, v in out.items() if k.startswith("g_")})


def gen_pdf_inds(trainer, torch):
    """sample_pdf's searchsorted indices exactly as the reference computes them: inside the whole hierarchical render (both
    hierarchical RENDER_CASES) and for the stand-alone sample_pdf inputs of stages.npz (recomputed from the same seed)."""
    from utils import nerf_util

    out = {}
    for name, case in RENDER_CASES.items():
        if case["num_fine"] == 0:
            continue
        inds = []
        _reference_render(trainer, torch, case, inds_out=inds)
        assert len(inds) == 1
        out[name] = inds[0].reshape(case["batch"], -1, case["num_fine"]).astype(np.int32)
    st = np.load(os.path.join(GOLD, "stages.npz"))
    bins, wts, u = st["pdf_bins"], st["pdf_w"], st["pdf_u"]
    orig_ss, orig_rand = torch.searchsorted, torch.rand
    rec = []

    def spy(*a, **k):
        r = orig_ss(*a, **k)
        rec.append(r.numpy().copy())
        return r

    torch.searchsorted = spy
    torch.rand = lambda shape, *a, **k: torch.from_numpy(u.copy())
    try:
        det = nerf_util.sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 16, det=True).numpy()
        rnd = nerf_util.sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 16, det=False).numpy()
    finally:
        torch.searchsorted, torch.rand = orig_ss, orig_rand
    assert np.array_equal(det, st["pdf_det"]) and np.array_equal(rnd, st["pdf_rand"])
    out["stage_det"], out["stage_rand"] = rec[0].astype(np.int32), rec[1].astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "pdf_inds.npz"), **out)
    print("pdf_inds", {k: (v.shape, int(v.min()), int(v.max())) for k, v in out.items()})


def gen_render(trainer, torch, only=None):
    for name, case in RENDER_CASES.items():
        if only is not None and name not in only:
            continue
        out = _reference_render(trainer, torch, case)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), case=json.dumps(case), **out)
        print(name, {k: v.shape for k, v in out.items()}, "acc mean %.4f" % out["acc_coarse"].mean())


def gen_stages(torch):
    """Stage-level goldens straight from the reference functions (no Trainer needed)."""
    from model.network.embedder import get_embedder
    from utils import nerf_util
    from utils.util import sample_from_triplane_new, voxel_feature
    from dataloader import data_util

    rs = np.random.RandomState(100)
    out = {}
    # positional encoding (embedder.py:32-61) on canonical-space magnitudes
    x = rs.uniform(-1.7, 1.7, size=(257, 3)).astype(np.float32)
    emb, dim = get_embedder(multires=8, input_dims=3, include_input=False)
    out["pe_x"], out["pe_y"] = x, emb(torch.from_numpy(x)).numpy()
    # bi-plane fetch incl. out-of-range coordinates (zeros padding)
    pl = rs.standard_normal((2, 1, 5, 7, 9)).astype(np.float32)
    q = rs.uniform(-1.3, 1.3, size=(1, 300, 3)).astype(np.float32)
    out["tp_planes"], out["tp_q"] = pl, q
    out["tp_y"] = sample_from_triplane_new(torch.from_numpy(q), torch.from_numpy(pl)).numpy()
    # trilinear border fetch
    vol = rs.uniform(0, 1, size=(1, 1, 4, 5, 6)).astype(np.float32)
    out["vx_vol"], out["vx_q"] = vol, q
    out["vx_y"] = voxel_feature(torch.from_numpy(q), torch.from_numpy(vol)).numpy()
    # sample_pdf, both modes
    bins = np.sort(rs.uniform(2.4, 5.0, size=(64, 63)).astype(np.float32), axis=-1)
    wts = rs.uniform(0, 1, size=(64, 62)).astype(np.float32) ** 4
    wts[:4] = 0.0                                      # all-zero weights: uniform pdf from the 1e-5 floor
    out["pdf_bins"], out["pdf_w"] = bins, wts
    out["pdf_det"] = nerf_util.sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 16, det=True).numpy()
    u = rs.uniform(0, 1, size=(64, 16)).astype(np.float32)
    orig = torch.rand
    torch.rand = lambda shape, *a, **k: torch.from_numpy(u.copy())
    try:
        out["pdf_rand"] = nerf_util.sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 16, det=False).numpy()
    finally:
        torch.rand = orig
    out["pdf_u"] = u
    # composite
    rf = rs.standard_normal((33, 17, 68)).astype(np.float32)
    z = np.sort(rs.uniform(2.4, 5.0, size=(33, 17)).astype(np.float32), axis=-1)
    rd = rs.standard_normal((33, 3)).astype(np.float32)
    bg = rs.uniform(0, 1, size=(33, 3)).astype(np.float32)
    res = nerf_util.volume_render_radiance_field(torch.from_numpy(rf.copy()), torch.from_numpy(z), torch.from_numpy(rd),
                                                 background_prior=torch.from_numpy(bg), act_feat=False)
    out["cmp_rf"], out["cmp_z"], out["cmp_rd"], out["cmp_bg"] = rf, z, rd, bg
    for n, r in zip(("rgb", "disp", "acc", "w", "depth"), res):
        out["cmp_" + n] = r.numpy()
    # ray generation (data_util.py:28-56)
    intr = np.array([700.0, 690.0, 0.49, 0.52], dtype=np.float32)
    c2w = np.array([[0.96, 0.05, -0.27, 0.3], [0.0, -0.98, -0.19, 0.1], [-0.28, 0.18, -0.94, 3.9]], dtype=np.float32)
    o, d = data_util.get_rays(6, 8, intr, torch.from_numpy(c2w))
    out["ray_intr"], out["ray_c2w"] = intr, c2w
    out["ray_o"], out["ray_d"] = o.reshape(-1, 3).numpy().copy(), d.reshape(-1, 3).numpy()
    np.savez_compressed(os.path.join(GOLD, "stages.npz"), **out)
    print("stages", sorted(out))


# upfirdn2d call shapes of the reference (model/styleUnet.py): Blur after convT-up, Blur before stride-2 conv,
# Upsample, Downsample, Haar analysis / synthesis, plus asymmetric / cropping / non-square cases.
UFD_CASES = [
    # name, shape NCHW, taps (1-D list or "haar_xx"), gain, up, down, pad(x0,x1,y0,y1)
    ("blur_up", (2, 5, 17, 17), [1, 3, 3, 1], 4.0, (1, 1), (1, 1), (1, 1, 1, 1)),
    ("blur_down", (2, 3, 16, 16), [1, 3, 3, 1], 1.0, (1, 1), (1, 1), (2, 2, 2, 2)),
    ("upsample", (1, 4, 9, 12), [1, 3, 3, 1], 4.0, (2, 2), (1, 1), (2, 1, 2, 1)),
    ("downsample", (1, 4, 14, 10), [1, 3, 3, 1], 1.0, (1, 1), (2, 2), (1, 1, 1, 1)),
    ("haar_ll_down", (1, 3, 8, 12), "ll", 1.0, (1, 1), (2, 2), (0, 0, 0, 0)),
    ("haar_lh_down", (1, 3, 8, 12), "lh", 1.0, (1, 1), (2, 2), (0, 0, 0, 0)),
    ("haar_hl_up", (1, 3, 6, 5), "hl", 1.0, (2, 2), (1, 1), (1, 0, 1, 0)),
    ("haar_hh_up", (1, 3, 6, 5), "hh", 1.0, (2, 2), (1, 1), (1, 0, 1, 0)),
    ("asym_crop", (1, 2, 11, 13), [1, 2, 4, 2, 1], 1.0, (3, 2), (2, 3), (-1, 2, 3, -2)),
    ("wide_tile", (1, 1, 40, 150), [1, 3, 3, 1], 1.0, (1, 1), (1, 1), (2, 1, 2, 1)),
]


def ufd_kernel(np_mod, taps, gain):
    if isinstance(taps, str):
        s = 1 / (2 ** 0.5)
        lo, hi = np_mod.array([[s, s]], dtype=np_mod.float32), np_mod.array([[-s, s]], dtype=np_mod.float32)
        a, b = {"ll": (lo, lo), "lh": (hi, lo), "hl": (lo, hi), "hh": (hi, hi)}[taps]
        return (a.T * b).astype(np_mod.float32)            # model/styleUnet.py:371-381
    k = np_mod.asarray(taps, dtype=np_mod.float32)
    k = k[None, :] * k[:, None]
    return (k / k.sum() * np_mod.float32(gain)).astype(np_mod.float32)


def gen_ops(torch):
    """model/op goldens from the reference's own CPU fallbacks (upfirdn2d.py:172-213, fused_act.py:107-119)."""
    from model.op.fused_act import fused_leaky_relu
    from model.op.upfirdn2d import upfirdn2d_native

    rs = np.random.RandomState(200)
    out = {}
    for name, shape, taps, gain, up, down, pad in UFD_CASES:
        x = rs.standard_normal(shape).astype(np.float32)
        k = ufd_kernel(np, taps, gain)
        y = upfirdn2d_native(torch.from_numpy(x), torch.from_numpy(k), up[0], up[1], down[0], down[1], *pad)
        out["ufd_%s_x" % name], out["ufd_%s_y" % name] = x, y.numpy()
    x = rs.standard_normal((3, 6, 5, 7)).astype(np.float32)
    b = rs.standard_normal(6).astype(np.float32)
    out["act_x"], out["act_b"] = x, b
    out["act_y"] = fused_leaky_relu(torch.from_numpy(x), torch.from_numpy(b)).numpy()
    out["act_y_nobias"] = fused_leaky_relu(torch.from_numpy(x)).numpy()
    x2 = rs.standard_normal((4, 10)).astype(np.float32)
    b2 = rs.standard_normal(10).astype(np.float32)
    out["act2_x"], out["act2_b"] = x2, b2
    out["act2_y"] = fused_leaky_relu(torch.from_numpy(x2), torch.from_numpy(b2)).numpy()
    np.savez_compressed(os.path.join(GOLD, "ops.npz"), **out)
    print("ops", len(out))


STYLEUNET_CASES = {
    # SWGAN_unet at reduced size (same code path as the shipped 128 -> 512: encoder 16,8 -> decoder 16,32,64 -> IWT 128)
    "swgan_32_128": dict(net="SWGAN_unet", kw=dict(inp_size=32, inp_ch=8, out_ch=3, out_size=128, style_dim=64, n_mlp=4), batch=2, seed=1),
    # plane generator configuration of model/nerf_model.py:39-42 at reduced size (cond 64x64, planes 32x32)
    # wavelet discriminator (stage two), reduced size
    "disc_64": dict(net="Discriminator", kw=dict(size=64, img_channel=3), batch=4, seed=4),
    "zxc_64_32": dict(net="StyleGAN_zxc", kw=dict(out_ch=16, out_size=32, style_dim=44, middle_size=16, zero_latent=False,
                                                   zero_noise=True, no_skip=True, n_mlp=4, inp_size=64, inp_ch=7), batch=2, seed=2),
    # FULL-SIZE networks (round 2): the shipped upsampler 128 -> 512 (train_avatarHD.py:110, config/singleview_512_HD_base.yml:33-37),
    # BASELINE.json configs[3]'s 512 -> 1024, and the 512 x 512 wavelet discriminator (train_avatarHD.py:112).  `stride` = the
    # fixture stores every stride-th pixel of the output image (the whole image is produced and compared at those pixels).
    "swgan_128_512": dict(net="SWGAN_unet", kw=dict(inp_size=128, inp_ch=64, out_ch=3, out_size=512, style_dim=64, n_mlp=4), batch=1, seed=6,
                          stride=8),
    "swgan_512_1024": dict(net="SWGAN_unet", kw=dict(inp_size=512, inp_ch=64, out_ch=3, out_size=1024, style_dim=64, n_mlp=4), batch=1, seed=7,
                           stride=16),
    "disc_512": dict(net="Discriminator", kw=dict(size=512, img_channel=3), batch=2, seed=8),
}
FULL_SIZE_CASES = ("swgan_128_512", "swgan_512_1024", "disc_512")


def styleunet_inputs(case, net):
    """Deterministic inputs of a STYLEUNET_CASES entry (shared with tests/test_styleunet_gpu.py)."""
    kw, B, seed = case["kw"], case["batch"], case["seed"]
    sd = synth.styleunet_state({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed)
    if case["net"] == "Discriminator":
        return sd, None, synth.named_normal("input.image", (B, kw["img_channel"], kw["size"], kw["size"]), seed), None
    style = synth.named_normal("input.style", (B, kw["style_dim"]), seed)
    cond = synth.named_normal("input.cond", (B, kw["inp_ch"], kw["inp_size"], kw["inp_size"]), seed)
    if case["net"] == "SWGAN_unet":
        noise = [synth.named_normal("input.noise%d" % i, (1, 1, 2 ** r, 2 ** r), seed)
                 for i, r in enumerate(r for r in range(net.middle_log_size + 1, net.log_size + 1) for _ in range(2))]
    else:
        noise = [synth.named_normal("input.noise0", (1, 1, 2 ** net.middle_log_size, 2 ** net.middle_log_size), seed)]
    return sd, style, cond, noise


def gen_styleunet(torch, only=None):
    import model.styleUnet as ref

    for name, case in STYLEUNET_CASES.items():
        if only is not None and name not in only:
            continue
        net = getattr(ref, case["net"])(**case["kw"])
        sd, style, cond, noise = styleunet_inputs(case, net)
        missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not missing.unexpected_keys and all(k.endswith(synth._FIXED) for k in missing.missing_keys), missing
        with torch.no_grad():
            if case["net"] == "SWGAN_unet":
                out = net([torch.from_numpy(style)], torch.from_numpy(cond), noise=[torch.from_numpy(n) for n in noise])
            elif case["net"] == "Discriminator":
                out = net(torch.from_numpy(cond))
            else:
                net.zero_noise[0] = torch.from_numpy(noise[0])        # the fixed first-layer draw (styleUnet.py:748)
                out, _ = net([torch.from_numpy(style)], torch.from_numpy(cond))
        keys = json.dumps({k: list(v.shape) for k, v in net.state_dict().items()})
        st = case.get("stride", 1)
        out = out[..., ::st, ::st] if st > 1 else out
        np.savez_compressed(os.path.join(GOLD, "styleunet_%s.npz" % name), out=out.numpy(), state_dict_shapes=keys)
        print(name, tuple(out.shape), "abs mean %.3f max %.3f" % (out.abs().mean(), out.abs().max()))


def trainer_inputs(seed=3, render_size=128, batch=1):
    """Deterministic inputs of the Trainer golden (shared with tests/test_trainer.py)."""
    sc = synth.scene(batch=batch, height=render_size, width=render_size, seed=seed)
    conds = {k: np.abs(synth.named_normal("input." + k, (batch, 7, 256, 256), seed)).astype(np.float32) % np.float32(1.0)
             for k in ("front_render_cond", "left_render_cond", "right_render_cond")}
    noise0 = {g: synth.named_normal("input.%s.noise0" % g, (1, 1, 16, 16), seed) for g in ("XY_gen", "YZ_gen")}
    return sc, conds, noise0


def gen_trainer(torch):
    """Whole-orchestrator golden: the unmodified reference Trainer.forward (mode 'validation', full 128x128 frame,
    hierarchical 64 + 16, perturb off) on CPU.  Stores every 4th pixel of the [1,67,128,128] render and the mask."""
    from oracle import ref_shim
    from model.nerf_trainer import Trainer

    cfg = ref_shim.load_cfg()
    cfg.nerf.validation.perturb = False
    net = Trainer(cfg, 4)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth.trainer_state(shapes, seed=3)
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not missing.unexpected_keys, missing
    sc, conds, noise0 = trainer_inputs()
    net.model_coarse.XY_gen.zero_noise[0] = torch.from_numpy(noise0["XY_gen"])
    net.model_coarse.YZ_gen.zero_noise[0] = torch.from_numpy(noise0["YZ_gen"])
    with torch.no_grad():
        render, mask, lat = net(mode="validation", fidx=None, render_full_img=True, ray_batch=torch.from_numpy(sc["ray_batch"]),
                                background_prior=torch.from_numpy(sc["background_prior"]), inv_head_T=torch.from_numpy(sc["inv_head_T"]),
                                **{k: torch.from_numpy(v) for k, v in conds.items()})
    print("trainer", tuple(render.shape), tuple(mask.shape), float(lat), "acc mean %.3f" % float(mask.mean()))
    np.savez_compressed(os.path.join(GOLD, "trainer_validation.npz"), render=render.numpy()[:, :, ::4, ::4], mask=mask.numpy()[:, :, ::4, ::4],
                        latent_code_loss=np.float32(lat), state_dict_shapes=json.dumps({k: list(v) for k, v in shapes.items()}))


TRAIN_GRAD_KEYS = ["latent_codes", "model_coarse.fc_alpha.weight", "model_coarse.fc_rgb.bias", "model_coarse.layers_xyz.0.bias",
                   "model_coarse.layers_xyz.1.weight", "model_coarse.XY_gen.conv_out.0.weight", "model_coarse.XY_gen.style.1.weight",
                   "model_coarse.YZ_gen.conv1.conv.modulation.bias", "model_coarse.YZ_gen.convs.5.activate.bias",
                   "model_coarse.YZ_gen.conv_in.1.weight", "headpose_skin_net.canonical_Wvolume.final_conv.weight",
                   "headpose_skin_net.canonical_Wvolume.filters.0.up.1.bias"]


def train_step_inputs(seed=5, batch=2, patch=16):
    """Deterministic inputs of the stage-one training-step golden (shared with tests/test_trainer.py): B patches of
    patch x patch rays (train_avatar.py:61-62 uses 64 x 64), targets, masks, the reference's random draws."""
    sc = synth.scene(batch=batch, crop=(248, 240, patch, patch), seed=seed)
    R = patch * patch
    conds = {k: np.abs(synth.named_normal("input." + k, (batch, 7, 256, 256), seed)).astype(np.float32) % np.float32(1.0)
             for k in ("front_render_cond", "left_render_cond", "right_render_cond")}
    noise0 = {g: synth.named_normal("input.%s.noise0" % g, (1, 1, 16, 16), seed) for g in ("XY_gen", "YZ_gen")}
    rnd = synth.randoms(batch, R, 64, 16, seed=seed + 7)
    rs = np.random.RandomState(seed + 1)
    target = rs.uniform(0, 1, size=(batch, R, 3)).astype(np.float32)
    mask = (rs.uniform(0, 1, size=(batch, R, 1)) > 0.4).astype(np.float32)
    return sc, conds, noise0, rnd, target, mask


def train_step_loss(torch, out, target, mask, mask_weight=0.01):
    """The differentiable part of the stage-one loss that needs no external weights (train_avatar.py:131-146 minus LPIPS and
    the volume-smoothness term): mse + mask BCE on the coarse and the fine pass + the latent-code regulariser."""
    import torch.nn.functional as F

    rgb_c, _, acc_c, _, rgb_f, _, acc_f, lat = out
    loss = F.mse_loss(rgb_c[..., :3], target) + mask_weight * F.binary_cross_entropy(acc_c.clip(1e-3, 1.0 - 1e-3), mask)
    loss = loss + F.mse_loss(rgb_f[..., :3], target) + mask_weight * F.binary_cross_entropy(acc_f.clip(1e-3, 1.0 - 1e-3), mask)
    return loss + lat


def gen_trainer_train(torch):
    """Whole training-step golden: the unmodified reference Trainer.forward(mode='train') + loss.backward() on CPU; stores the
    loss and the gradients of a dozen parameters spread over every sub-network."""
    from oracle import ref_shim
    from model.nerf_trainer import Trainer

    cfg = ref_shim.load_cfg()
    cfg.nerf.train.perturb, cfg.nerf.train.radiance_field_noise_std = True, 0.1
    cfg.nerf.train.num_coarse, cfg.nerf.train.num_fine = 64, 16
    net = Trainer(cfg, 4)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth.trainer_state(shapes, seed=3)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    with torch.no_grad():
        net.latent_codes.copy_(torch.from_numpy(synth.named_normal("latent_codes", (4, 32), 5) * np.float32(0.1)))
    sc, conds, noise0, rnd, target, mask = train_step_inputs()
    net.model_coarse.XY_gen.zero_noise[0] = torch.from_numpy(noise0["XY_gen"])
    net.model_coarse.YZ_gen.zero_noise[0] = torch.from_numpy(noise0["YZ_gen"])
    B, R = sc["ray_batch"].shape[:2]
    q_rand = [rnd["t_rand"], rnd["u_rand"].reshape(B * R, -1)]
    q_randn = [rnd["unit_coarse"].reshape(B * R, -1), rnd["unit_fine"].reshape(B * R, -1)]
    orig_rand, orig_randn = torch.rand, torch.randn

    def fake_rand(shape, *a, **k):
        v = q_rand.pop(0)
        assert tuple(shape) == v.shape, (tuple(shape), v.shape)
        return torch.from_numpy(v.copy())

    def fake_randn(shape, *a, **k):
        v = q_randn.pop(0)
        assert tuple(shape) == v.shape, (tuple(shape), v.shape)
        return torch.from_numpy(v.copy())

    torch.rand, torch.randn = fake_rand, fake_randn
    try:
        out = net(mode="train", fidx=torch.tensor([1, 3]), render_full_img=False, ray_batch=torch.from_numpy(sc["ray_batch"]),
                  background_prior=torch.from_numpy(sc["background_prior"]), inv_head_T=torch.from_numpy(sc["inv_head_T"]),
                  **{k: torch.from_numpy(v) for k, v in conds.items()})
    finally:
        torch.rand, torch.randn = orig_rand, orig_randn
    assert not q_rand and not q_randn
    loss = train_step_loss(torch, out, torch.from_numpy(target), torch.from_numpy(mask))
    loss.backward()
    params = dict(net.named_parameters())
    res = {"g_" + k: params[k].grad.numpy() for k in TRAIN_GRAD_KEYS}
    print("trainer_train loss %.6f" % float(loss), {k: "%.3g" % np.abs(v).max() for k, v in res.items()})
    np.savez_compressed(os.path.join(GOLD, "trainer_train_step.npz"), loss=np.float32(loss.item()), rgb_fine=out[4].detach().numpy(),
                        acc_fine=out[6].detach().numpy(), **res)


# ------------------------------------------------------------------------------------------------
# stage-two training step (train_avatarHD.py:201-303) at reduced size, round 2
# ------------------------------------------------------------------------------------------------
STAGE_TWO_CASE = dict(batch=2, render_size=32, gen_size=128, num_coarse=32, num_fine=8, seed=9, latent=64, n_mlp=4)
STAGE_TWO_GRAD_KEYS = {
    # gradients recorded right after the backward of each phase (before the optimiser step)
    "d": ["convs.0.conv1.0.weight", "convs.2.conv2.1.weight", "final_conv.0.weight", "final_linear.1.weight", "from_rgbs.1.conv.0.weight"],
    "r1": ["convs.0.conv1.0.weight", "convs.1.conv2.1.weight", "final_linear.0.weight", "from_rgbs.0.conv.0.weight"],
    "g": ["convs.0.conv.weight", "convs.3.conv.modulation.weight", "to_rgbs.1.conv.weight", "style.2.weight", "comb_convs.0.0.weight",
          "cond_convs.0.conv1.0.weight", "convs.1.noise.weight"],
    "nerf": ["latent_codes", "model_coarse.layers_xyz.0.weight", "model_coarse.fc_rgbFeat.weight", "model_coarse.XY_gen.conv_out.0.weight",
             "model_coarse.YZ_gen.convs.5.conv.weight", "headpose_skin_net.canonical_Wvolume.final_conv.weight"],
}


def subsample(v, cap=4096):
    """Every k-th element of the flattened array, k = ceil(size / cap): keeps the fixtures small while touching the whole tensor."""
    flat = np.asarray(v).reshape(-1)
    return flat[::max(1, -(-flat.size // cap))].copy()


def stage_two_inputs(case=None):
    """Deterministic inputs of the stage-two step golden (shared with tests/test_stage_two_parity_gpu.py)."""
    c = case or STAGE_TWO_CASE
    B, rs, gs, seed = c["batch"], c["render_size"], c["gen_size"], c["seed"]
    sc = synth.scene(batch=B, height=rs, width=rs, seed=seed)
    R = rs * rs
    conds = {k: np.abs(synth.named_normal("input." + k, (B, 7, 256, 256), seed)).astype(np.float32) % np.float32(1.0)
             for k in ("front_render_cond", "left_render_cond", "right_render_cond")}
    noise0 = {g: synth.named_normal("input.%s.noise0" % g, (1, 1, 16, 16), seed) for g in ("XY_gen", "YZ_gen")}
    rnd = {"d": synth.randoms(B, R, c["num_coarse"], c["num_fine"], seed=seed + 7),
           "g": synth.randoms(B, R, c["num_coarse"], c["num_fine"], seed=seed + 8)}
    z = {ph: synth.named_normal("input.z_" + ph, (B, c["latent"]), seed) for ph in ("d", "g")}
    res = [r for r in range(4, int(np.log2(gs))) for _ in range(2)]      # SWGAN_unet.make_noise: middle_log_size+1 .. log_size
    gnoise = {ph: [synth.named_normal("input.gnoise_%s%d" % (ph, i), (1, 1, 2 ** r, 2 ** r), seed) for i, r in enumerate(res)]
              for ph in ("d", "g")}
    rs_ = np.random.RandomState(seed + 1)
    gt_hr = (rs_.uniform(0, 1, size=(B, 3, gs, gs)) * 2 - 1).astype(np.float32)
    gt_mask = (rs_.uniform(0, 1, size=(B, 1, rs, rs)) > 0.4).astype(np.float32)
    return dict(scene=sc, conds=conds, noise0=noise0, rnd=rnd, z=z, gnoise=gnoise, gt_hr=gt_hr, gt_mask=gt_mask)


def stage_two_states(shapes_nerf, shapes_gen, shapes_disc, seed):
    """Name-keyed initial weights of the three networks (the discriminator's first convolutions are scaled down so that its
    logits stay O(1) on random inputs)."""
    sd_n = synth.trainer_state(shapes_nerf, seed=seed)
    sd_n["latent_codes"] = (synth.named_normal("latent_codes", shapes_nerf["latent_codes"], seed) * np.float32(0.1)).astype(np.float32)
    return sd_n, synth.styleunet_state(shapes_gen, seed + 1), synth.styleunet_state(shapes_disc, seed + 2)


def gen_skin(torch):
    """Deformation_Field_new.pretrain_wc / sample_volume (model/Skinning_Field.py:65-68,101-125) of the unmodified reference:
    two Adam iterations of the head-box fit from seeded weights and a seeded torch generator (make_volume_pts' jitter)."""
    from oracle import ref_shim
    from model.nerf_trainer import Trainer

    cfg = ref_shim.load_cfg()
    net = Trainer(cfg, 2)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth.trainer_state(shapes, seed=11)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    skin = net.headpose_skin_net
    torch.manual_seed(77)
    import tqdm as _tq
    skin.pretrain_wc(num_iter=2, vol_thr=cfg.models.coarse.Head_bounding)
    with torch.no_grad():
        vol = skin.canonical_Wvolume()
        pts = torch.from_numpy(np.random.RandomState(5).uniform(-1.4, 1.4, size=(64, 3)).astype(np.float32))
        smp = skin.sample_volume(pts)
    np.savez_compressed(os.path.join(GOLD, "skin_pretrain.npz"), vol=vol.numpy()[:, :, ::8, ::8, ::8], pts=pts.numpy(), sample=smp.numpy(),
                        b0=dict(skin.named_parameters())["canonical_Wvolume.final_conv.bias"].detach().numpy())
    print("skin", vol.shape, float(vol[0, 1].mean()), smp.shape)


def make_fixture_dataset(dst):
    """A tiny dataset in the reference's on-disk formats (data_preprocessing/fit_video.py:336-339,353-418): split JSON, two
    frames x two views (one of them view_name '8', which the loader skips), 16 x 16 images / masks, 32 x 32 condition PNGs."""
    import cv2

    rs = np.random.RandomState(300)
    os.makedirs(dst, exist_ok=True)
    frames = []
    for f, fidx in enumerate((7, 3)):
        inst = "inst_%d" % fidx
        os.makedirs(os.path.join(dst, inst), exist_ok=True)
        for v in ("front", "left", "right"):
            yy, xx = np.mgrid[0:32, 0:32]
            normal = np.stack([(xx * 7 + f * 13) % 256, (yy * 5 + 31) % 256, (xx + yy) * 3 % 256], -1).astype(np.uint8)
            normal[(xx - 16) ** 2 + (yy - 16) ** 2 > 150 + 20 * f] = 0                     # outside the head: |normal| == 0
            render = rs.randint(0, 256, size=(32, 32, 3)).astype(np.uint8)
            cv2.imwrite(os.path.join(dst, inst, "ortho_%s_normal_256_baseGama.png" % v), cv2.cvtColor(normal, cv2.COLOR_RGB2BGR))
            cv2.imwrite(os.path.join(dst, inst, "ortho_%s_render_256_baseGama.png" % v), cv2.cvtColor(render, cv2.COLOR_RGB2BGR))
        views = []
        for vi, name in enumerate(("0", "8", "2")):
            ang = 0.2 * vi + 0.1 * f
            c, s_ = np.cos(ang), np.sin(ang)
            pose = np.eye(4)
            pose[:3, :3] = np.array([[c, 0, s_], [0, -1, 0], [s_, 0, -c]])
            pose[:3, 3] = [0.1 * vi, -0.05 * f, 4.0 + 0.1 * vi]
            ori = pose.copy()
            ori[:3, 3] += [0.01, 0.02, -0.03]
            img = rs.randint(0, 256, size=(16, 16, 3)).astype(np.uint8)
            mask = (rs.uniform(size=(16, 16)) > 0.5).astype(np.uint8) * 255
            ip, mp = "%s/img_%s.png" % (inst, name), "%s/mask_%s.png" % (inst, name)
            cv2.imwrite(os.path.join(dst, ip), img)
            cv2.imwrite(os.path.join(dst, mp), np.repeat(mask[..., None], 3, -1))
            views.append({"view_name": name, "transform_matrix": pose.tolist(), "transform_matrix_ori": ori.tolist(), "file_path": ip,
                          "mask_path": mp})
        ht = np.eye(4)
        a = 0.15 + 0.1 * f
        ht[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        ht[:3, 3] = [0.02, -0.01 * f, 0.03]
        frames.append({"fidx": fidx, "inst_dir": inst, "head_transformation": ht.tolist(), "mutiview_info_ls": views})
    meta = {"img_res": 16, "mutiview_intr_ls": [[24.0, 23.0, 0.5, 0.48], [25.0, 25.0, 0.51, 0.5], [22.0, 24.0, 0.49, 0.52]], "frames": frames}
    with open(os.path.join(dst, "sv_v31_all.json"), "w") as fjs:
        json.dump(meta, fjs)


def gen_dataset(torch):
    """The unmodified reference MultiView_ImgDataset (dataloader/dataloader.py:38-230, mode 'val') on the fixture dataset:
    every item's ray tensor, colours, condition tensors and inv_head_T."""
    from oracle import ref_shim

    dst = os.path.join(GOLD, "dataset")
    make_fixture_dataset(dst)
    cfg = ref_shim.load_cfg()
    cfg.dataset.cond_render_res = 32
    from dataloader.dataloader import MultiView_ImgDataset

    out = {}
    cwd = os.getcwd()
    os.chdir(dst)                       # the reference opens the JSON's paths as they are
    try:
        for tag, res in (("r32", 32), ("r24", 24)):          # 24: the cv2.INTER_LINEAR resize branch (dataloader.py:220-225)
            cfg.dataset.cond_render_res = res
            ds = MultiView_ImgDataset("sv_v31_all.json", "val", cfg, down_sample=1.0, white_bg=True)
            out["len"] = np.int32(len(ds))
            for i in range(len(ds)):
                idx, d = ds[i]
                if tag == "r32":
                    out["item%d_mv_rays" % i] = d["mv_rays"].numpy()
                    out["item%d_gt" % i] = d["mv_rays_gt_color"].numpy()
                    out["item%d_inv_head_T" % i] = d["inv_head_T"].numpy()
                    out["item%d_fidx" % i] = np.int32(ds.frames[i]["fidx"])
                for v in ("front", "left", "right"):
                    out["item%d_%s_%s" % (i, v, tag)] = d["%s_render_cond" % v].numpy()
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "dataset_items.npz"), **out)
    print("dataset", int(out["len"]), sorted(k for k in out if k.startswith("item0")))


def gen_stage_two(torch):
    """The loop body of train_avatarHD.py:201-303 (iteration i = 0, so the R1 branch runs) restated around the UNMODIFIED
    reference modules and loss functions, LPIPS term omitted (its weights are not available offline), random draws and mixing
    noise supplied.  Records the losses of each phase, a handful of gradients per network taken right after each backward,
    and a few weights after the optimiser steps."""
    import torch.nn.functional as F
    from oracle import ref_shim
    from model.nerf_trainer import Trainer
    from model.styleUnet import SWGAN_unet, Discriminator
    from utils.styleUnet_util import d_logistic_loss, d_r1_loss, g_nonsaturating_loss, requires_grad, accumulate

    c = STAGE_TWO_CASE
    B, rs, gs = c["batch"], c["render_size"], c["gen_size"]
    cfg = ref_shim.load_cfg("singleview_512_HD_base.yml")
    cfg.models.StyleUnet.inp_size, cfg.models.StyleUnet.out_size = rs, gs
    cfg.nerf.train.num_coarse, cfg.nerf.train.num_fine = c["num_coarse"], c["num_fine"]
    nerf = Trainer(cfg, 4)
    gen = SWGAN_unet(inp_size=rs, inp_ch=64, out_size=gs, out_ch=3, style_dim=c["latent"], c_dim=0, n_mlp=c["n_mlp"], channel_multiplier=2)
    g_ema = SWGAN_unet(inp_size=rs, inp_ch=64, out_size=gs, out_ch=3, style_dim=c["latent"], c_dim=0, n_mlp=c["n_mlp"], channel_multiplier=2)
    disc = Discriminator(gs, 3, channel_multiplier=2, c_dim=0)
    shp = lambda m: {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd_n, sd_g, sd_d = stage_two_states(shp(nerf), shp(gen), shp(disc), c["seed"])
    for m, sd in ((nerf, sd_n), (gen, sd_g), (disc, sd_d)):
        miss = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not miss.unexpected_keys, miss
    accumulate(g_ema, gen, 0)                                                                              # train_avatarHD.py:115
    inp = stage_two_inputs(c)
    nerf.model_coarse.XY_gen.zero_noise[0] = torch.from_numpy(inp["noise0"]["XY_gen"])
    nerf.model_coarse.YZ_gen.zero_noise[0] = torch.from_numpy(inp["noise0"]["YZ_gen"])
    g_ratio, d_ratio, lr = 4 / 5, 16 / 17, 1e-3                                                            # :116-118
    g_optim = torch.optim.Adam(gen.parameters(), lr=lr * g_ratio, betas=(0 ** g_ratio, 0.99 ** g_ratio))
    d_optim = torch.optim.Adam(disc.parameters(), lr=lr * d_ratio, betas=(0 ** d_ratio, 0.99 ** d_ratio))
    n_optim = torch.optim.Adam([{"params": nerf.parameters()}], lr=cfg.optimizer.lr)
    sc = inp["scene"]
    Rn = rs * rs
    t = torch.from_numpy
    gt_hr, gt_mask = t(inp["gt_hr"]), t(inp["gt_mask"])
    inp_data = dict(mode="train", fidx=torch.tensor([1, 3]), render_full_img=True, ray_batch=t(sc["ray_batch"]),
                    background_prior=t(sc["background_prior"]), inv_head_T=t(sc["inv_head_T"]), **{k: t(v) for k, v in inp["conds"].items()})
    orig_rand, orig_randn = torch.rand, torch.randn

    def with_draws(rnd, fn):
        q_rand = [rnd["t_rand"], rnd["u_rand"].reshape(B * Rn, -1)]
        q_randn = [rnd["unit_coarse"].reshape(B * Rn, -1), rnd["unit_fine"].reshape(B * Rn, -1)]

        def fake(q):
            def f(shape, *a, **k):
                v = q.pop(0)
                assert tuple(shape) == v.shape, (tuple(shape), v.shape)
                return torch.from_numpy(v.copy())
            return f

        torch.rand, torch.randn = fake(q_rand), fake(q_randn)
        try:
            r = fn()
        finally:
            torch.rand, torch.randn = orig_rand, orig_randn
        assert not q_rand and not q_randn
        return r

    out = {}
    grads = lambda m, keys, tag: out.update({"g_%s_%s" % (tag, k): subsample(dict(m.named_parameters())[k].grad.numpy()) for k in keys})
    i = 0
    gan_w = min(1e-3 * 1.1 ** (i // 500), 0.1)                                                            # :205-206
    gt_lr = F.interpolate(F.interpolate(gt_hr, size=(rs, rs), mode="bilinear", align_corners=True), size=(gs, gs), mode="bilinear",
                          align_corners=True)                                                              # :202-204
    # ---- D step (:211-231)
    requires_grad(nerf, False), requires_grad(gen, False), requires_grad(disc, True)
    with torch.no_grad():
        render, _, _ = with_draws(inp["rnd"]["d"], lambda: nerf(**inp_data))
        fake_img = gen([t(inp["z"]["d"])], render[:, 3:], noise=[t(n) for n in inp["gnoise"]["d"]])
    out["render_d"] = render.numpy()[:, :, ::2, ::2].copy()
    out["fake_d"] = fake_img.numpy()[:, :, ::4, ::4].copy()
    fake_pred, real_pred = disc(fake_img, flat_pose=None), disc(gt_hr, flat_pose=None)
    d_loss = d_logistic_loss(real_pred, fake_pred) * gan_w
    out["d_loss"] = np.float32((d_loss / gan_w).item())
    out["real_pred"], out["fake_pred"] = real_pred.detach().numpy().copy(), fake_pred.detach().numpy().copy()
    disc.zero_grad()
    d_loss.backward()
    grads(disc, STAGE_TWO_GRAD_KEYS["d"], "d")
    d_optim.step()
    # ---- R1 (:233-240), i % d_reg_every == 0
    gt_req = gt_hr.clone().requires_grad_(True)
    real_pred = disc(gt_req, flat_pose=None)
    r1_loss = d_r1_loss(real_pred, gt_req) * gan_w
    out["r1"] = np.float32((r1_loss / gan_w).item())
    disc.zero_grad()
    (10.0 / 2 * r1_loss * 16 + 0 * real_pred[0]).backward()
    grads(disc, STAGE_TWO_GRAD_KEYS["r1"], "r1")
    d_optim.step()
    # ---- G step (:243-280)
    requires_grad(nerf, True)
    render, mask, lat = with_draws(inp["rnd"]["g"], lambda: nerf(**inp_data))
    lr_img = F.interpolate(render[:, :3], size=(gs, gs), mode="bilinear", align_corners=True)
    rgb_loss = F.mse_loss(lr_img, gt_lr)
    nerf_loss = rgb_loss + 1.0 * lat
    mask_loss = cfg.experiment.mask_weight * F.binary_cross_entropy(mask.clip(1e-3, 1.0 - 1e-3), gt_mask)
    nerf_loss = nerf_loss + mask_loss
    g_loss = nerf_loss
    requires_grad(gen, True)
    fake_img = gen([t(inp["z"]["g"])], render[:, 3:], noise=[t(n) for n in inp["gnoise"]["g"]])
    requires_grad(disc, False)
    fake_pred = disc(fake_img, flat_pose=None)
    g_ns = g_nonsaturating_loss(fake_pred)
    hr_l1 = F.l1_loss(fake_img, gt_hr)
    g_loss = g_loss + g_ns * gan_w + hr_l1
    out.update(rgb_loss=np.float32(rgb_loss.item()), mask_loss=np.float32(mask_loss.item()), lat=np.float32(lat.item()),
               g_nonsat=np.float32(g_ns.item()), hr_l1=np.float32(hr_l1.item()), g_loss=np.float32(g_loss.item()))
    nerf.zero_grad()
    gen.zero_grad()
    g_loss.backward()
    grads(gen, STAGE_TWO_GRAD_KEYS["g"], "g")
    grads(nerf, STAGE_TWO_GRAD_KEYS["nerf"], "nerf")
    g_optim.step()
    n_optim.step()
    accumulate(g_ema, gen, 0.5 ** (32 / (10 * 1000)))                                                     # :162, :303
    # a few weights after the whole iteration
    out["w_disc_final_linear.1.weight"] = subsample(dict(disc.named_parameters())["final_linear.1.weight"].detach().numpy())
    out["w_gen_to_rgbs.1.conv.weight"] = subsample(dict(gen.named_parameters())["to_rgbs.1.conv.weight"].detach().numpy())
    out["w_ema_to_rgbs.1.conv.weight"] = subsample(dict(g_ema.named_parameters())["to_rgbs.1.conv.weight"].detach().numpy())
    out["w_nerf_model_coarse.fc_rgb.weight"] = subsample(dict(nerf.named_parameters())["model_coarse.fc_rgb.weight"].detach().numpy())
    np.savez_compressed(os.path.join(GOLD, "stage_two_step.npz"), **out)
    print("stage_two", {k: (float(v) if v.ndim == 0 else "%s max %.3g" % (v.shape, np.abs(v).max())) for k, v in out.items()})


def main():
    from oracle import ref_shim

    cfg = ref_shim.load_cfg()
    import torch

    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(GOLD, exist_ok=True)
    if "--train-only" in sys.argv:
        return gen_trainer_train(torch)
    if "--round2" in sys.argv:          # fixtures added in round 2 only (the round-1 files stay byte-identical)
        from model.nerf_trainer import Trainer

        which = [a for a in sys.argv[2:] if not a.startswith("--")]
        if not which or "styleunet" in which:
            gen_styleunet(torch, only=FULL_SIZE_CASES)
        if not which or "render" in which:
            trainer = Trainer(cfg, 4)
            gen_render(trainer, torch, only=ROUND2_RENDER_CASES)
            gen_pdf_inds(trainer, torch)
        if not which or "stage_two" in which:
            gen_stage_two(torch)
        if not which or "skin" in which:
            gen_skin(torch)
        if not which or "dataset" in which:
            gen_dataset(torch)
        return
    if "--bwd-only" not in sys.argv:
        gen_stages(torch)
        gen_ops(torch)
        gen_styleunet(torch)
        gen_trainer(torch)
    from model.nerf_trainer import Trainer

    trainer = Trainer(cfg, 4)
    if "--bwd-only" not in sys.argv:
        gen_render(trainer, torch)
    gen_render_bwd(trainer, torch)
    gen_trainer_train(torch)


if __name__ == "__main__":
    main()
