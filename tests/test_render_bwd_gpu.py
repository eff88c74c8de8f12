"""Parity of the CUDA render backward (hav_render_backward through the C ABI) with the gradients of the UNMODIFIED
reference's autograd (tests/golden/render_bwd_*.npz) and with the autograd oracle.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from havatar_b200 import render, synth
from oracle import render_oracle as ro
from oracle import render_oracle_torch as rot
from test_render_bwd_oracle import BWD_CASES, bwd_case

pytestmark = pytest.mark.gpu

# 16-bit MLP / gradient operands with fp32 accumulation: max error relative to the largest entry of each gradient tensor.
# Measured on B200 (scripts/check_bwd.py): fp16 1e-4 .. 9e-3 on the MLP tensors, 1.0 .. 1.4e-2 on the planes, 1.7 .. 2.1e-2
# on the skinning volume (its gradient is made of differences of 16-bit texels); bf16 about 3x .. 5x that.
REL_TOL = {"fp16": 3e-2, "bf16": 1.5e-1}


def _dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def cuda_grads(sc, case, rnd, cot, precision, via_autograd=False):
    w = {k: _dev(v) for k, v in sc["weights"].items()}
    kw = {k: _dev(v) for k, v in (rnd or {}).items()}
    planes, wvol = _dev(sc["planes"]), _dev(sc["wvol"])
    fine = case["num_fine"] > 0
    if via_autograd:
        planes.requires_grad_(True), wvol.requires_grad_(True)
        for v in w.values():
            v.requires_grad_(True)
        out = render.render_rays_autograd(_dev(sc["ray_batch"]), _dev(sc["background_prior"]), _dev(sc["inv_head_T"]), planes, wvol,
                                          w, case["num_coarse"], case["num_fine"], precision=precision, **kw)
        loss = sum((getattr(out, k) * _dev(c).reshape(getattr(out, k).shape)).sum() for k, c in cot.items())
        loss.backward()
        g = {"planes": planes.grad, "wvol": wvol.grad}
        g.update({k: v.grad for k, v in w.items()})
    else:
        out, ctx = render.render_rays(_dev(sc["ray_batch"]), _dev(sc["background_prior"]), _dev(sc["inv_head_T"]), planes, wvol, w,
                                      case["num_coarse"], case["num_fine"], precision=precision, want_z_fine=True, return_ctx=True, **kw)
        g = render.render_backward(ctx, g_rgb_coarse=_dev(cot["rgb_coarse"]), g_depth_coarse=_dev(cot["depth_coarse"]),
                                   g_acc_coarse=_dev(cot["acc_coarse"]), g_rgb_fine=_dev(cot.get("rgb_fine")),
                                   g_depth_fine=_dev(cot.get("depth_fine")), g_acc_fine=_dev(cot.get("acc_fine")))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in g.items()}


def rel_errors(got, ref):
    return {k: float(np.abs(got[k].reshape(ref[k].shape) - ref[k]).max() / (np.abs(ref[k]).max() + 1e-30)) for k in ref}


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("name", BWD_CASES)
def test_backward_matches_reference_gradients(golden_dir, name, precision):
    z, case, sc, rnd, cot = bwd_case(golden_dir, name)
    got = cuda_grads(sc, case, rnd, cot, precision)
    ref = {k[2:]: z[k] for k in z.files if k.startswith("g_")}
    err = rel_errors(got, ref)
    bad = {k: e for k, e in err.items() if not e < REL_TOL[precision]}
    assert not bad, (name, precision, err)


def test_backward_through_autograd_function(golden_dir):
    """Same numbers when the call goes through torch autograd (render_rays_autograd -> loss.backward())."""
    z, case, sc, rnd, cot = bwd_case(golden_dir, "render_bwd_hier_rand")
    a = cuda_grads(sc, case, rnd, cot, "fp16", via_autograd=False)
    b = cuda_grads(sc, case, rnd, cot, "fp16", via_autograd=True)
    for k in a:   # atomics: summation order differs between runs
        assert np.abs(a[k] - b[k].reshape(a[k].shape)).max() <= 1e-4 * np.abs(a[k]).max() + 1e-12, k


def test_backward_linear_in_the_upstream_gradient(golden_dir):
    """Size-independent property: gradients are linear in the cotangents (scaling them by 64 scales every gradient by 64,
    across the device-chosen loss scale), and zero cotangents give exactly zero."""
    z, case, sc, rnd, cot = bwd_case(golden_dir, "render_bwd_coarse")
    g1 = cuda_grads(sc, case, rnd, cot, "fp16")
    g64 = cuda_grads(sc, case, rnd, {k: v * np.float32(64) for k, v in cot.items()}, "fp16")
    g0 = cuda_grads(sc, case, rnd, {k: v * np.float32(0) for k, v in cot.items()}, "fp16")
    for k in g1:
        assert np.abs(g64[k] - 64 * g1[k]).max() <= 1e-3 * np.abs(64 * g1[k]).max() + 1e-12, k
        assert not g0[k].any(), k


def test_backward_full_train_batch_against_oracle_on_gpu():
    """The stage-one training shape (B=2 x 64x64 patch, 64 + 16 hierarchical, perturb + noise; train_avatar.py:61-62,
    config/singleview_512_base.yml:105-123) against the torch-CUDA autograd of the reference's ATen sequence."""
    sc = synth.scene(batch=2, crop=(224, 224, 64, 64), seed=60)
    B, R = sc["ray_batch"].shape[:2]
    r = synth.randoms(B, R, 64, 16, seed=67)
    rnd = {k: r[k] for k in ("t_rand", "noise_coarse", "u_rand", "noise_fine")}
    cot = synth.cotangents(B, R, True, seed=71, scale=1.0 / (B * R))
    case = dict(num_coarse=64, num_fine=16)
    got = cuda_grads(sc, case, rnd, cot, "fp16")
    _, ref = rot.render_rays_grad(sc["ray_batch"], sc["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"],
                                  sc["weights"], ro.default_boxes(), 64, 16, cotangents=cot, device="cuda", **rnd)
    err = rel_errors(got, ref)
    bad = {k: e for k, e in err.items() if not e < 3e-2}
    assert not bad, err
