"""The CTA-pair render kernel (render_tc3.cu: tcgen05 cta_group::2) and its split-precision mode (HAV_PREC_FP16X3) against
the reference goldens and against the other kernels.  pytest -m gpu."""
import json
import os

import numpy as np
import pytest
import torch

from havatar_b200 import render, synth
from oracle import render_oracle as ro
from test_render_gpu import CASES, TOL_FP32, TOL_TC, _dev, _tol, case_scene

pytestmark = pytest.mark.gpu


def run(sc, nc, nf, rnd, **kw):
    r = {} if rnd is None else {k: _dev(rnd[k]) for k in ("t_rand", "noise_coarse", "u_rand", "noise_fine")}
    w = {k: _dev(v) for k, v in sc["weights"].items()}
    out = render.render_rays(_dev(sc["ray_batch"]), _dev(sc["background_prior"]), _dev(sc["inv_head_T"]), _dev(sc["planes"]),
                             _dev(sc["wvol"]), w, nc, nf, **r, **kw)
    torch.cuda.synchronize()
    return {k: (None if v is None else v.cpu().numpy()) for k, v in out._asdict().items()}


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_cta_pair_kernel_matches_reference_golden(golden_dir, name, precision):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    case = json.loads(str(z["case"]))
    sc, rnd = case_scene(case)
    got = run(sc, case["num_coarse"], case["num_fine"], rnd, precision=precision, cta_pairs=True)
    tol = TOL_TC if precision == "fp16" else {k: 20 * v for k, v in TOL_TC.items()}
    for k in z.files:
        if k != "case":
            err = np.abs(got[k].reshape(z[k].shape) - z[k]).max()
            assert err < _tol(k, tol), (name, precision, k, float(err))


@pytest.mark.parametrize("name", CASES)
def test_split_precision_mode_matches_reference_golden_at_fp32_tolerance(golden_dir, name):
    """HAV_PREC_FP16X3: the SAME tolerances as the fp32 CUDA-core mode (1e-4 rgb / acc, 2e-4 depth)."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    case = json.loads(str(z["case"]))
    sc, rnd = case_scene(case)
    got = run(sc, case["num_coarse"], case["num_fine"], rnd, precision="fp16x3")
    worst = {}
    for k in z.files:
        if k != "case":
            worst[k] = float(np.abs(got[k].reshape(z[k].shape) - z[k]).max())
    print(name, worst)
    for k, err in worst.items():
        assert err < _tol(k, TOL_FP32), (name, k, err)


def test_cta_pair_kernel_equals_single_cta_kernel_bit_for_bit():
    """Same operands, same accumulation order per output element: the two 16-bit kernels must agree exactly, including a ragged
    number of ray blocks (odd block count -> one CTA of the last cluster runs a masked dummy block)."""
    for R, nc, nf in ((128 * 5 + 17, 24, 0), (128 * 3, 16, 6)):
        sc = synth.scene(batch=1, crop=(100, 0, 2, 512), seed=90)
        sc["ray_batch"], sc["background_prior"] = sc["ray_batch"][:, :R], sc["background_prior"][:, :R]
        a = run(sc, nc, nf, None, precision="fp16")
        b = run(sc, nc, nf, None, precision="fp16", cta_pairs=True)
        for k in ("rgb_coarse", "depth_coarse", "acc_coarse", "weights_max") + (("rgb_fine", "acc_fine") if nf else ()):
            assert np.array_equal(a[k], b[k]), (R, k, float(np.abs(a[k] - b[k]).max()))


@pytest.mark.parametrize("mode", ["pairs", "split"])
def test_full_frame_subset_invariance(mode):
    """512 x 512 x 64 (BASELINE.json configs[1]): any gathered subset reproduces the full render's rows bit for bit; the split
    mode additionally agrees with the oracle at fp32 tolerance on a strided subset."""
    kw = dict(precision="fp16", cta_pairs=True) if mode == "pairs" else dict(precision="fp16x3")
    sc = synth.scene(batch=1, height=512, width=512, seed=0)
    full = run(sc, 64, 0, None, **kw)
    R = 512 * 512
    assert np.isfinite(full["rgb_coarse"]).all() and 0.2 < full["acc_coarse"].mean() < 0.98
    idx = np.sort(np.random.RandomState(5).choice(R, size=4099, replace=False))
    sub = dict(sc, ray_batch=sc["ray_batch"][:, idx], background_prior=sc["background_prior"][:, idx])
    part = run(sub, 64, 0, None, **kw)
    for k in ("rgb_coarse", "depth_coarse", "acc_coarse", "weights_max"):
        assert np.array_equal(part[k], full[k][:, idx]), k
    idx2 = np.arange(0, R, 257)
    sub2 = dict(sc, ray_batch=sc["ray_batch"][:, idx2], background_prior=sc["background_prior"][:, idx2])
    ref = ro.render_rays(sub2["ray_batch"], sub2["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"], sc["weights"],
                         ro.default_boxes(), 64, 0)
    tol = 2e-3 if mode == "pairs" else 1e-4
    assert np.abs(full["rgb_coarse"][:, idx2] - ref["rgb_coarse"]).max() < tol
    assert np.abs(full["acc_coarse"][:, idx2, 0] - ref["acc_coarse"]).max() < tol
