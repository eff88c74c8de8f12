"""Pins oracle/ (the numpy restatement) against outputs of the unmodified reference
(tests/golden/*.npz, minted by oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from havatar_b200 import synth
from oracle import render_oracle as ro

# fp32 restatement vs fp32 reference: only summation order differs (BLAS / cumprod / sum).
ATOL = 2e-5


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


def oracle_render_case(case):
    sc = synth.scene(batch=case["batch"], crop=tuple(case["crop"]), seed=case["seed"],
                     plane_hw=tuple(case.get("plane_hw", (128, 128))), vol_dhw=tuple(case.get("vol_dhw", (64, 64, 64))))
    B, R = sc["ray_batch"].shape[:2]
    kw = {}
    if case["rand"]:
        rnd = synth.randoms(B, R, case["num_coarse"], case["num_fine"], seed=case["seed"] + 7)
        kw = dict(t_rand=rnd["t_rand"], noise_coarse=rnd["noise_coarse"], u_rand=rnd["u_rand"], noise_fine=rnd["noise_fine"])
    return ro.render_rays(sc["ray_batch"], sc["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"],
                          sc["weights"], ro.default_boxes(), case["num_coarse"], case["num_fine"], **kw)


@pytest.mark.parametrize("name", ["render_c32_s32", "render_hier_det", "render_hier_rand", "render_oddshape", "render_c64_s32"])
def test_render_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    case = json.loads(str(g.pop("case")))
    out = oracle_render_case(case)
    assert 0.2 < g["acc_coarse"].mean() < 0.98, "degenerate golden scene"
    for k, ref in g.items():
        got = out[k].reshape(ref.shape)
        err = np.abs(got - ref).max()
        assert err < ATOL, (name, k, err)


def test_stages_match_reference(golden_dir):
    g = _load(golden_dir, "stages")
    assert np.abs(ro.positional_encode(g["pe_x"]) - g["pe_y"]).max() < 1e-6
    y = np.stack([ro.bilinear_zeros(g["tp_planes"][0, 0], g["tp_q"][0][:, [0, 1]]),
                  ro.bilinear_zeros(g["tp_planes"][1, 0], g["tp_q"][0][:, [2, 1]])], axis=-1)
    assert np.abs(y - g["tp_y"][0]).max() < 1e-6
    v = ro.trilinear_border(g["vx_vol"][0, 0], g["vx_q"][0])
    assert np.abs(v - g["vx_y"][0, :, 0]).max() < 1e-6
    s_det, _ = ro.sample_pdf(g["pdf_bins"], g["pdf_w"], 16)
    # det=True puts the last sample at u == 1.0 exactly, where searchsorted(right=True) lands on either
    # side of cdf[-1] (1 +- 1ulp depending on the cumsum order): a knife-edge of the reference itself
    # (utils/nerf_util.py:102-115).  Both sides are within (1 - t) * bin width of each other.
    assert np.abs(s_det - g["pdf_det"])[:, :-1].max() < 1e-5
    assert np.abs(s_det - g["pdf_det"])[:, -1].max() < 2e-2
    s_rnd, _ = ro.sample_pdf(g["pdf_bins"], g["pdf_w"], 16, u_rand=g["pdf_u"])
    assert np.abs(s_rnd - g["pdf_rand"]).max() < 1e-4   # small-denominator bins amplify cumsum rounding
    rgb, disp, acc, w, depth = ro.composite(g["cmp_rf"], g["cmp_z"], g["cmp_rd"], g["cmp_bg"])
    for got, key in ((rgb, "rgb"), (acc, "acc"), (w, "w"), (depth, "depth")):
        assert np.abs(got - g["cmp_" + key]).max() < 1e-5, key
    o, d = ro.get_rays(6, 8, g["ray_intr"], g["ray_c2w"])
    assert np.array_equal(o, g["ray_o"])
    assert np.abs(d - g["ray_d"]).max() < 1e-6


def test_feature_interleave_order(golden_dir):
    """feature index = 2*c + plane (utils/util.py:388 stack on the last dim, nerf_model.py:99 flatten)."""
    g = _load(golden_dir, "stages")
    f = ro.plane_features(g["tp_q"][0], g["tp_planes"][:, 0], np.ones(3, np.float32), np.zeros(3, np.float32))
    assert np.abs(f.reshape(-1, 5, 2) - g["tp_y"][0]).max() < 1e-6


@pytest.mark.parametrize("name", ["render_c32_s32", "render_hier_det", "render_hier_rand", "render_oddshape"])
def test_torch_port_matches_reference(golden_dir, name):
    """oracle/render_oracle_torch.py (the CPU baseline bench.py times) against the same goldens."""
    from oracle import render_oracle_torch as rt

    g = _load(golden_dir, name)
    case = json.loads(str(g.pop("case")))
    sc = synth.scene(batch=case["batch"], crop=tuple(case["crop"]), seed=case["seed"],
                     plane_hw=tuple(case.get("plane_hw", (128, 128))), vol_dhw=tuple(case.get("vol_dhw", (64, 64, 64))))
    B, R = sc["ray_batch"].shape[:2]
    kw = {}
    if case["rand"]:
        rnd = synth.randoms(B, R, case["num_coarse"], case["num_fine"], seed=case["seed"] + 7)
        kw = dict(t_rand=rnd["t_rand"], noise_coarse=rnd["noise_coarse"], u_rand=rnd["u_rand"], noise_fine=rnd["noise_fine"])
    out = rt.render_rays(sc["ray_batch"], sc["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"],
                         sc["weights"], ro.default_boxes(), case["num_coarse"], case["num_fine"], chunk=100, **kw)
    for k, ref in g.items():
        err = np.abs(out[k].reshape(ref.shape) - ref).max()
        assert err < ATOL, (name, k, err)


def test_sample_pdf_indices_match_reference_bit_for_bit(golden_dir):
    """Integer bookkeeping: the searchsorted(right=True) indices of utils/nerf_util.py:102, recorded
    while the unmodified reference ran (oracle/gen_golden.py::gen_pdf_inds), against the oracle on identical inputs."""
    g, gi = _load(golden_dir, "stages"), _load(golden_dir, "pdf_inds")
    _, inds = ro.sample_pdf(g["pdf_bins"], g["pdf_w"], 16, u_rand=g["pdf_u"])
    assert np.array_equal(inds.astype(np.int32), gi["stage_rand"])
    _, inds = ro.sample_pdf(g["pdf_bins"], g["pdf_w"], 16)
    # det=True: the last sample sits at u == 1.0 exactly, a knife-edge of the reference itself (cdf[-1] = 1 +- 1 ulp, see above)
    assert np.array_equal(inds.astype(np.int32)[:, :-1], gi["stage_det"][:, :-1])
    assert np.abs(inds[:, -1] - gi["stage_det"][:, -1]).max() <= 1


@pytest.mark.parametrize("name", ["render_hier_det", "render_hier_rand"])
def test_whole_path_sample_pdf_indices(golden_dir, name):
    """Indices inside the whole hierarchical render.  The coarse weights feeding the cdf carry summation-order noise (1e-6
    class), so an index may move by one where u falls within that distance of a cdf entry: exact on >= 99.9 % of the samples,
    never off by more than one."""
    g, gi = _load(golden_dir, name), _load(golden_dir, "pdf_inds")
    case = json.loads(str(g.pop("case")))
    out = oracle_render_case(case)
    got, ref = out["pdf_inds"].reshape(gi[name].shape), gi[name]
    skip_last = not case["rand"]
    a, b = (got[..., :-1], ref[..., :-1]) if skip_last else (got, ref)
    assert (a != b).mean() <= 1e-3 and np.abs(a.astype(np.int64) - b).max() <= 1
