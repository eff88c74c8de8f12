"""The autograd oracle of the render backward (oracle/render_oracle_torch.render_rays_grad) against the gradients the
UNMODIFIED reference produced with its own autograd (tests/golden/render_bwd_*.npz, minted by oracle/gen_golden.py)."""
import json
import os

import numpy as np
import pytest

from havatar_b200 import synth
from oracle import render_oracle as ro
from oracle import render_oracle_torch as rot

BWD_CASES = ["render_bwd_coarse", "render_bwd_hier_rand"]


def bwd_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    case = json.loads(str(z["case"]))
    sc = synth.scene(batch=case["batch"], crop=tuple(case["crop"]), seed=case["seed"], plane_hw=tuple(case["plane_hw"]),
                     vol_dhw=tuple(case["vol_dhw"]))
    B, R = sc["ray_batch"].shape[:2]
    rnd = None
    if case["rand"]:
        r = synth.randoms(B, R, case["num_coarse"], case["num_fine"], seed=case["seed"] + 7)
        rnd = {k: r[k] for k in ("t_rand", "noise_coarse", "u_rand", "noise_fine")}
    cot = synth.cotangents(B, R, case["num_fine"] > 0, seed=case["seed"] + 11)
    return z, case, sc, rnd, cot


@pytest.mark.parametrize("name", BWD_CASES)
def test_autograd_oracle_matches_reference_gradients(golden_dir, name):
    z, case, sc, rnd, cot = bwd_case(golden_dir, name)
    out, g = rot.render_rays_grad(sc["ray_batch"], sc["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"],
                                  sc["weights"], ro.default_boxes(), case["num_coarse"], case["num_fine"], cotangents=cot,
                                  **(rnd or {}))
    for k in z.files:
        if k.startswith("out_"):
            assert np.abs(out[k[4:]].reshape(z[k].shape) - z[k]).max() < 2e-5, k
    for k in z.files:
        if not k.startswith("g_"):
            continue
        ref, got = z[k], g[k[2:]].reshape(z[k].shape)
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-9, (k, float(np.abs(got - ref).max()), float(np.abs(ref).max()))
