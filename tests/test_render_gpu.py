"""Parity of the CUDA render path (through the C ABI) with the oracle and with the golden vectors
minted from the unmodified reference.  Run on the B200 box: pytest -m gpu."""
import json
import os

import numpy as np
import pytest
import torch

from havatar_b200 import render, synth
from oracle import render_oracle as ro

pytestmark = pytest.mark.gpu

# fp32 CUDA-core mode: same arithmetic as the reference, only summation order / libm differ.
TOL_FP32 = dict(rgb=1e-4, depth=2e-4, acc=1e-4, wmax=1e-4)
# 16-bit tensor-core operands, fp32 accumulate.  SURVEY.md section 8d allows 2e-2 abs; fp16 measures ~1e-4 on these scenes
# (bench.py reports 83 dB PSNR against the fp32 mode on the full frame), so the gate is 10x tighter than the allowance
TOL_TC = dict(rgb=2e-3, depth=4e-3, acc=2e-3, wmax=2e-3)
# render_c64_s32 = BASELINE.json configs[0] literally: one 64 x 64 crop, 32 samples per ray, coarse only
CASES = ["render_c32_s32", "render_hier_det", "render_hier_rand", "render_oddshape", "render_c64_s32"]


def _dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_cuda(sc, num_coarse, num_fine, rnd=None, precision="fp32", want_z_fine=False, boxes=None):
    kw = {}
    if rnd is not None:
        kw = {k: _dev(rnd[k]) for k in ("t_rand", "noise_coarse", "u_rand", "noise_fine")}
    w = {k: _dev(v) for k, v in sc["weights"].items()}
    out = render.render_rays(_dev(sc["ray_batch"]), _dev(sc["background_prior"]), _dev(sc["inv_head_T"]),
                             _dev(sc["planes"]), _dev(sc["wvol"]), w, num_coarse, num_fine, boxes=boxes,
                             precision=precision, want_z_fine=want_z_fine, **kw)
    torch.cuda.synchronize()
    return {k: (None if v is None else v.cpu().numpy()) for k, v in out._asdict().items()}


def case_scene(case):
    sc = synth.scene(batch=case["batch"], crop=tuple(case["crop"]), seed=case["seed"],
                     plane_hw=tuple(case.get("plane_hw", (128, 128))), vol_dhw=tuple(case.get("vol_dhw", (64, 64, 64))))
    B, R = sc["ray_batch"].shape[:2]
    rnd = synth.randoms(B, R, case["num_coarse"], case["num_fine"], seed=case["seed"] + 7) if case["rand"] else None
    return sc, rnd


def _tol(key, tol):
    return tol["rgb"] if key.startswith("rgb") else tol["depth"] if key.startswith("depth") else \
        tol["acc"] if key.startswith("acc") else tol["wmax"]


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_render_matches_reference_golden(golden_dir, name, precision):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    case = json.loads(str(z["case"]))
    sc, rnd = case_scene(case)
    got = run_cuda(sc, case["num_coarse"], case["num_fine"], rnd, precision)
    tol = TOL_FP32 if precision == "fp32" else TOL_TC
    if precision == "bf16":
        tol = {k: 20 * v for k, v in tol.items()}   # 8-bit mantissa operands: 4e-2
    for k in z.files:
        if k == "case":
            continue
        ref = z[k]
        err = np.abs(got[k].reshape(ref.shape) - ref).max()
        assert err < _tol(k, tol), (name, precision, k, float(err))


def test_fine_depths_and_sort_bookkeeping_match_oracle():
    """z_fine = sort(cat(z[::2], sample_pdf(...))) (nerf_trainer.py:166-170): same values, sorted, same count."""
    case = dict(batch=2, crop=(200, 260, 16, 16), num_coarse=64, num_fine=16, seed=10)
    sc = synth.scene(batch=2, crop=case["crop"], seed=10)
    rnd = synth.randoms(2, 256, 64, 16, seed=17)
    got = run_cuda(sc, 64, 16, rnd, "fp32", want_z_fine=True)
    ref = ro.render_rays(sc["ray_batch"], sc["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"],
                         sc["weights"], ro.default_boxes(), 64, 16, t_rand=rnd["t_rand"],
                         noise_coarse=rnd["noise_coarse"], u_rand=rnd["u_rand"], noise_fine=rnd["noise_fine"])
    zf = got["z_fine"]
    assert zf.shape == (2, 256, 48)
    assert np.all(np.diff(zf, axis=-1) >= 0), "fine depths must be sorted"
    assert np.abs(zf - ref["z_fine"]).max() < 5e-4
    # every retained coarse depth z[::2] is present (to 1 ulp: ATen's own CPU and CUDA linspace/jitter kernels
    # differ at that level through FMA contraction, so the last bit is not defined by the reference)
    zc = ro.coarse_z(sc["ray_batch"][0, :, 6], sc["ray_batch"][0, :, 7], 64, rnd["t_rand"][0])[:, ::2]
    for r in range(0, 256, 37):
        d = np.abs(zc[r][:, None] - zf[0, r][None, :]).min(axis=1)
        assert d.max() <= 2e-6, (r, float(d.max()))


def test_ragged_and_edge_sizes_match_oracle():
    """ray counts that are not a multiple of the 128-ray tile, 1 ray, and the minimum sample count."""
    for R, S, nf in ((1, 2, 0), (5, 3, 2), (129, 16, 0), (300, 9, 5)):
        sc = synth.scene(batch=1, crop=(250, 250, 1, R), seed=40 + R) if R <= 512 else None
        got = run_cuda(sc, S, nf, None, "fp32")
        ref = ro.render_rays(sc["ray_batch"], sc["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"],
                             sc["weights"], ro.default_boxes(), S, nf)
        for k in ("rgb_coarse", "depth_coarse", "acc_coarse", "weights_max") + (("rgb_fine", "acc_fine") if nf else ()):
            err = np.abs(got[k].reshape(ref[k].shape) - ref[k]).max()
            assert err < 2e-4, (R, S, nf, k, float(err))


def test_empty_ray_batch_is_a_noop():
    sc = synth.scene(batch=1, crop=(0, 0, 1, 4), seed=1)
    sc["ray_batch"] = sc["ray_batch"][:, :0]
    sc["background_prior"] = sc["background_prior"][:, :0]
    got = run_cuda(sc, 8, 0)
    assert got["rgb_coarse"].shape == (1, 0, 67)


def test_no_background_and_background_identity():
    """rgb[:3] += (1 - acc) * bg (utils/nerf_util.py:70-71): with-bg minus without-bg == 1 - acc for bg == 1."""
    sc = synth.scene(batch=1, crop=(240, 240, 16, 16), seed=3)
    a = run_cuda(sc, 24, 0)
    sc2 = dict(sc, background_prior=None)
    b = run_cuda(sc2, 24, 0)
    d = a["rgb_coarse"][..., :3] - b["rgb_coarse"][..., :3]
    assert np.abs(d - (1.0 - a["acc_coarse"])).max() < 1e-6
    assert np.array_equal(a["rgb_coarse"][..., 3:], b["rgb_coarse"][..., 3:])


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_full_frame_ray_bookkeeping_is_bit_exact(precision):
    """BASELINE.json configs[1] size (512x512 rays x 64 samples).  Rays are independent, so rendering any
    gathered subset of the frame must reproduce the full render's rows bit for bit (the reference's chunk
    loop relies on the same property, model/nerf_trainer.py:65-71); plus physical invariants."""
    sc = synth.scene(batch=1, height=512, width=512, seed=0)
    full = run_cuda(sc, 64, 0, None, precision)
    R = 512 * 512
    assert np.isfinite(full["rgb_coarse"]).all()
    acc = full["acc_coarse"]
    assert acc.min() >= 0.0 and acc.max() <= 1.0 + 1e-5
    assert 0.2 < acc.mean() < 0.98
    assert (full["weights_max"] <= acc + 1e-6).all()
    near, far = sc["ray_batch"][0, :, 6], sc["ray_batch"][0, :, 7]
    dz = full["depth_coarse"][0, :, 0]
    assert (dz <= far * acc[0, :, 0] + 1e-3).all() and (dz >= near * acc[0, :, 0] - 1e-3).all()
    rs = np.random.RandomState(5)
    idx = np.sort(rs.choice(R, size=4099, replace=False))      # ragged: not a multiple of the tile
    sub = dict(sc, ray_batch=sc["ray_batch"][:, idx], background_prior=sc["background_prior"][:, idx])
    part = run_cuda(sub, 64, 0, None, precision)
    for k in ("rgb_coarse", "depth_coarse", "acc_coarse", "weights_max"):
        assert np.array_equal(part[k], full[k][:, idx]), k
    # against the oracle on a strided subset the CPU finishes in seconds
    idx2 = np.arange(0, R, 257)
    sub2 = dict(sc, ray_batch=sc["ray_batch"][:, idx2], background_prior=sc["background_prior"][:, idx2])
    ref = ro.render_rays(sub2["ray_batch"], sub2["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"],
                         sc["weights"], ro.default_boxes(), 64, 0)
    tol = 1e-4 if precision == "fp32" else 2e-3
    assert np.abs(full["rgb_coarse"][:, idx2] - ref["rgb_coarse"]).max() < tol
    assert np.abs(full["acc_coarse"][:, idx2, 0] - ref["acc_coarse"]).max() < tol


def test_get_rays_matches_oracle_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "stages.npz"))
    rays = render.get_rays(6, 8, g["ray_intr"], g["ray_c2w"], near=2.4, far=5.0).cpu().numpy()
    assert np.array_equal(rays[:, :3], g["ray_o"])           # origins: bit-exact
    assert np.abs(rays[:, 3:6] - g["ray_d"]).max() < 1e-6
    assert np.all(rays[:, 6] == np.float32(2.4)) and np.all(rays[:, 7] == np.float32(5.0))
    o, d = ro.get_rays(512, 512, g["ray_intr"], g["ray_c2w"])
    big = render.get_rays(512, 512, g["ray_intr"], g["ray_c2w"], near=2.4, far=5.0).cpu().numpy()
    assert np.abs(big[:, 3:6] - d).max() < 1e-6              # ray r <-> pixel (r // W, r % W)


def test_bad_arguments_raise():
    sc = synth.scene(batch=1, crop=(0, 0, 2, 2), seed=1)
    with pytest.raises(RuntimeError):
        run_cuda(sc, 1, 0)                                    # num_coarse < 2
    with pytest.raises(RuntimeError):
        render.render_rays(torch.zeros(1, 4, 8), None, torch.zeros(1, 4, 3), torch.zeros(2, 1, 64, 8, 8),
                           torch.zeros(1, 2, 4, 4, 4), {}, 8)  # CPU tensors


def test_sample_pdf_indices_are_bit_exact(golden_dir):
    """SURVEY.md section 8 a9 "int indices bit-exact": hav_sample_pdf runs the device function the render kernels use between
    their two passes on the stage fixture's inputs; its searchsorted indices must equal the ones the unmodified reference
    computed (tests/golden/pdf_inds.npz, recorded by wrapping torch.searchsorted at utils/nerf_util.py:102) and the oracle's."""
    g, gi = np.load(os.path.join(golden_dir, "stages.npz")), np.load(os.path.join(golden_dir, "pdf_inds.npz"))
    bins, w, u = _dev(g["pdf_bins"]), _dev(g["pdf_w"]), _dev(g["pdf_u"])
    smp, inds = render.sample_pdf(bins, w, 16, u)
    torch.cuda.synchronize()
    assert np.array_equal(inds.cpu().numpy(), gi["stage_rand"])
    assert np.array_equal(inds.cpu().numpy(), ro.sample_pdf(g["pdf_bins"], g["pdf_w"], 16, u_rand=g["pdf_u"])[1].astype(np.int32))
    assert np.abs(smp.cpu().numpy() - g["pdf_rand"]).max() < 1e-4
    smp, inds = render.sample_pdf(bins, w, 16, None)
    torch.cuda.synchronize()
    # det=True: u[-1] == 1.0 exactly against cdf[-1] == 1 +- 1 ulp is a knife-edge of the reference itself (its CPU and CUDA
    # cumsum kernels disagree there); every other column is exact
    assert np.array_equal(inds.cpu().numpy()[:, :-1], gi["stage_det"][:, :-1])
    assert np.abs(inds.cpu().numpy()[:, -1] - gi["stage_det"][:, -1]).max() <= 1
    assert np.abs(smp.cpu().numpy() - g["pdf_det"])[:, :-1].max() < 2e-5   # depths ~2.4-5: 4 ulp


@pytest.mark.parametrize("name", ["render_hier_det", "render_hier_rand"])
def test_whole_path_sample_pdf_indices(golden_dir, name):
    """The same indices exported from inside the fused render (hav_render_args.pdf_inds, fp32 mode) against the reference's.
    The cdf is built from the coarse weights, which carry 1e-6-class summation-order noise, so an index may move by one where u
    falls within that distance of a cdf entry: exact on >= 99.9 % of the samples, never off by more than one."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    gi = np.load(os.path.join(golden_dir, "pdf_inds.npz"))[name]
    case = json.loads(str(z["case"]))
    sc, rnd = case_scene(case)
    kw = {} if rnd is None else {k: _dev(rnd[k]) for k in ("t_rand", "noise_coarse", "u_rand", "noise_fine")}
    w = {k: _dev(v) for k, v in sc["weights"].items()}
    out = render.render_rays(_dev(sc["ray_batch"]), _dev(sc["background_prior"]), _dev(sc["inv_head_T"]), _dev(sc["planes"]),
                             _dev(sc["wvol"]), w, case["num_coarse"], case["num_fine"], precision="fp32", want_pdf_inds=True, **kw)
    torch.cuda.synchronize()
    got = out.pdf_inds.cpu().numpy()
    assert got.shape == gi.shape and got.dtype == np.int32
    a, b = (got, gi) if case["rand"] else (got[..., :-1], gi[..., :-1])
    assert (a != b).mean() <= 1e-3 and np.abs(a.astype(np.int64) - b).max() <= 1


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_in_kernel_ray_generation_equals_uploaded_rays(golden_dir, precision):
    """rays generated inside the render kernel from an 18-float camera block (hav_render_args.camera)
    give the render of hav_get_rays' ray tensor bit for bit -- full frame, and a patch of pixels through pixel_index
    (the dataloader's select_inds, dataloader/dataloader.py:160-170)."""
    g = np.load(os.path.join(golden_dir, "stages.npz"))
    H, W = 48, 40
    sc = synth.scene(batch=2, height=H, width=W, seed=70)
    intr = np.stack([g["ray_intr"] * np.float32([0.1, 0.1, 1, 1]), g["ray_intr"] * np.float32([0.12, 0.11, 1.02, 0.97])])
    c2w = np.stack([g["ray_c2w"], g["ray_c2w"]])
    c2w[1, :, 3] += np.float32([0.1, -0.05, 0.2])
    near, far = np.float32([2.3, 2.5]), np.float32([4.9, 5.1])
    rays = torch.stack([render.get_rays(H, W, intr[b], c2w[b], float(near[b]), float(far[b])) for b in range(2)])
    cam = render.camera_block(intr, c2w, near, far)
    w = {k: _dev(v) for k, v in sc["weights"].items()}
    common = (_dev(sc["inv_head_T"]), _dev(sc["planes"]), _dev(sc["wvol"]), w, 16, 4)
    a = render.render_rays(rays, _dev(sc["background_prior"]), *common, precision=precision)
    b = render.render_rays(None, _dev(sc["background_prior"]), *common, precision=precision, camera=cam, img_hw=(H, W))
    torch.cuda.synchronize()
    for k in ("rgb_coarse", "depth_coarse", "acc_coarse", "weights_max", "rgb_fine", "depth_fine", "acc_fine"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    rs = np.random.RandomState(3)
    pix = np.stack([np.sort(rs.choice(H * W, size=131, replace=False)) for _ in range(2)]).astype(np.int32)
    sub = torch.stack([rays[i, torch.from_numpy(pix[i]).long().cuda()] for i in range(2)])
    bg = _dev(sc["background_prior"][:, :131])
    a = render.render_rays(sub, bg, *common, precision=precision)
    b = render.render_rays(None, bg, *common, precision=precision, camera=cam, img_hw=(H, W), pixel_index=_dev(pix))
    torch.cuda.synchronize()
    for k in ("rgb_coarse", "acc_coarse", "rgb_fine", "depth_fine"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k


def test_trained_magnitude_inputs_saturate_loudly_in_fp16_and_render_in_bf16():
    """Range stress (round-1 VERDICT weak #1).  (a) planes and both hidden layers' weights x16 (trained-like magnitudes, hidden
    activations in the thousands): fp16 must still track the fp32 mode and report a clean range status.  (b) x128: hidden
    activations pass 65504; fp16 conversions saturate (cvt.rn.satfinite), i.e. the render would be finite but WRONG --
    render_rays(check_range=True) must raise instead of returning it, and bf16 (fp32 exponent range) must still render."""
    def scaled(f, head_div=1.0):
        sc = synth.scene(batch=1, crop=(240, 240, 16, 16), seed=5)
        sc["planes"] = (sc["planes"] * np.float32(f)).astype(np.float32)
        for k in ("layers_xyz.0.weight", "layers_xyz.1.weight"):
            sc["weights"][k] = (sc["weights"][k] * np.float32(f)).astype(np.float32)
        # keep the outputs O(1)-O(10) so that absolute tolerances mean something: undo (most of) the f^3 gain in the linear heads
        for k in ("fc_alpha.weight", "fc_rgbFeat.weight"):
            sc["weights"][k] = (sc["weights"][k] / np.float32(head_div)).astype(np.float32)
        w = {k: _dev(v) for k, v in sc["weights"].items()}
        return (_dev(sc["ray_batch"]), _dev(sc["background_prior"]), _dev(sc["inv_head_T"]), _dev(sc["planes"]), _dev(sc["wvol"]), w, 24, 0)

    args = scaled(16.0, 256.0)        # head weights stay fp16-normal (3e-4); hidden activations reach ~1e3
    ref = render.render_rays(*args, precision="fp32")
    h = render.render_rays(*args, precision="fp16", check_range=True)          # must not raise
    torch.cuda.synchronize()
    assert float((h.rgb_coarse - ref.rgb_coarse).abs().max()) < 5e-3
    assert float((h.acc_coarse - ref.acc_coarse).abs().max()) < 5e-3
    args = scaled(128.0, 128.0 ** 3)  # hidden activations ~1e5 > 65504
    ref = render.render_rays(*args, precision="fp32")
    with pytest.raises(RuntimeError, match="fp16 operand range"):
        render.render_rays(*args, precision="fp16", check_range=True)
    b = render.render_rays(*args, precision="bf16", check_range=True)
    torch.cuda.synchronize()
    assert torch.isfinite(b.rgb_coarse).all()
    assert float((b.rgb_coarse - ref.rgb_coarse).abs().max()) < 1e-1           # 8-bit mantissa operands, no clipping
    # texels beyond 65504 are reported too (bit 0)
    big = list(scaled(1.0))
    big[3] = big[3] * 2.0e5
    with pytest.raises(RuntimeError, match="fp16 operand range"):
        render.render_rays(*big, precision="fp16", check_range=True)
