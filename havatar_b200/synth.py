"""Seeded synthetic inputs for the render hot path (numpy only, platform-stable RandomState streams).

Shared by tests, bench.py, __graft_entry__.smoke() and oracle/gen_golden.py so that the golden
fixtures under tests/golden/ only have to store reference OUTPUTS: every input below is regenerated
bit-identically from its seed.  Shapes and value ranges follow SURVEY.md section 8(d).
"""
import math

import numpy as np

F32 = np.float32

MLP_SHAPES = {
    "layers_xyz.0.weight": (128, 176), "layers_xyz.0.bias": (128,),
    "layers_xyz.1.weight": (128, 128), "layers_xyz.1.bias": (128,),
    "fc_alpha.weight": (1, 128), "fc_alpha.bias": (1,),
    "fc_rgbFeat.weight": (64, 128), "fc_rgbFeat.bias": (64,),
    "fc_rgb.weight": (3, 64), "fc_rgb.bias": (3,),
}


def mlp_weights(seed=0, alpha_bias=0.5, gain=1.0):
    """nn.Linear-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) init (reference: model/nerf_model.py:46-51),
    drawn from numpy so it is reproducible across torch versions.  fc_alpha.bias += alpha_bias gives a
    non-degenerate opacity (acc ~ 0.7; random init alone renders acc ~ 1e-3)."""
    rs = np.random.RandomState(seed)
    w = {}
    for name, shape in MLP_SHAPES.items():
        layer = name.rsplit(".", 1)[0]
        fan_in = MLP_SHAPES[layer + ".weight"][1]
        bound = gain / math.sqrt(fan_in)
        w[name] = rs.uniform(-bound, bound, size=shape).astype(F32)
    w["fc_alpha.bias"] = (w["fc_alpha.bias"] + F32(alpha_bias)).astype(F32)
    return w


def planes(seed=1, batch=1, ch=64, h=128, w=128, scale=0.5):
    """Bi-plane feature maps [2,B,C,H,W] (reference layout: model/nerf_model.py:85)."""
    rs = np.random.RandomState(seed)
    return (rs.standard_normal((2, batch, ch, h, w)) * scale).astype(F32)


def _upsample_linear(a, n, axis):
    m = a.shape[axis]
    x = np.linspace(0, m - 1, n)
    i0 = np.minimum(np.floor(x).astype(np.int64), m - 2)
    t = (x - i0).astype(F32)
    a0, a1 = np.take(a, i0, axis=axis), np.take(a, i0 + 1, axis=axis)
    shp = [1] * a.ndim
    shp[axis] = n
    t = t.reshape(shp)
    return (a0 * (1 - t) + a1 * t).astype(F32)


def skin_volume(seed=2, d=64, h=64, w=64, coarse=6):
    """Smooth synthetic skinning-weight volume [1,2,D,H,W] = cat[x, 1-x] like the VolumeDecoder output
    (reference: model/network/voxel_encoder.py:150-179, consumed at model/Skinning_Field.py:79)."""
    rs = np.random.RandomState(seed)
    x = rs.uniform(0.0, 1.0, size=(coarse, coarse, coarse)).astype(F32)
    for ax, n in enumerate((d, h, w)):
        x = _upsample_linear(x, n, ax)
    return np.stack([x, (F32(1.0) - x).astype(F32)], axis=0)[None].astype(F32)


def head_pose(batch=1, seed=3):
    """inv_head_T [B,4,3] = [R^-1 ; -t] (reference: dataloader/dataloader.py:215-216)."""
    rs = np.random.RandomState(seed)
    out = np.zeros((batch, 4, 3), dtype=F32)
    for b in range(batch):
        yaw, pitch = math.radians(10.0 * (b + 1)), math.radians(5.0)
        ry = np.array([[math.cos(yaw), 0, math.sin(yaw)], [0, 1, 0], [-math.sin(yaw), 0, math.cos(yaw)]])
        rx = np.array([[1, 0, 0], [0, math.cos(pitch), -math.sin(pitch)], [0, math.sin(pitch), math.cos(pitch)]])
        rot = ry @ rx
        t = rs.uniform(-0.05, 0.05, size=3)
        out[b, :3] = np.linalg.inv(rot).astype(F32)
        out[b, 3] = (-t).astype(F32)
    return out


def camera_rays(height, width, crop=None, dist=4.0, focal_mul=1.5, near=-1.6, far=1.0):
    """Pinhole camera at (0,0,dist) looking down -z, rays via the reference formula
    (dataloader/data_util.py:28-56); near/far = dist + {near,far} (dataloader/dataloader.py:174-177).
    crop = (y0, x0, h, w) selects a pixel window.  -> ray_batch [R,8] = o3 d3 near far, row-major pixels."""
    K = np.eye(3, dtype=F32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = focal_mul * width, focal_mul * height, 0.5 * width, 0.5 * height
    Kinv = np.linalg.inv(K)
    y0, x0, h, w = crop if crop is not None else (0, 0, height, width)
    jj, ii = np.meshgrid(np.arange(y0, y0 + h, dtype=F32), np.arange(x0, x0 + w, dtype=F32), indexing="ij")
    pix = np.stack([ii, jj, np.ones_like(ii)], axis=-1).reshape(-1, 3)
    rot = np.diag([1.0, -1.0, -1.0]).astype(F32)
    d = (pix @ Kinv.T.astype(F32)) @ rot.T
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(F32)
    o = np.broadcast_to(np.array([0.0, 0.0, dist], dtype=F32), d.shape)
    nf = np.broadcast_to(np.array([dist + near, dist + far], dtype=F32), (d.shape[0], 2))
    return np.concatenate([o, d, nf], axis=-1).astype(F32)


def camera_params(height, width, dist=4.0, focal_mul=1.5, near=-1.6, far=1.0):
    """The camera of camera_rays in the dataloader's own parameterisation (dataloader/data_util.py:28-56, dataloader.py:174-177):
    intr = (fx, fy, cx, cy) with the principal point as a fraction of the image size, c2w [3,4], near, far."""
    intr = np.array([focal_mul * width, focal_mul * height, 0.5, 0.5], dtype=F32)
    c2w = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, dist]], dtype=F32)
    return intr, c2w, F32(dist + near), F32(dist + far)


def scene(batch=1, height=512, width=512, crop=None, seed=0, plane_hw=(128, 128), vol_dhw=(64, 64, 64)):
    """Everything one render call needs, as a dict of float32 numpy arrays."""
    rays = camera_rays(height, width, crop)
    rays = np.broadcast_to(rays[None], (batch,) + rays.shape).copy()
    return {
        "ray_batch": rays,
        "background_prior": np.ones((batch, rays.shape[1], 3), dtype=F32),
        "inv_head_T": head_pose(batch, seed + 3),
        "planes": planes(seed + 1, batch, 64, plane_hw[0], plane_hw[1]),
        "wvol": skin_volume(seed + 2, *vol_dhw),
        "weights": mlp_weights(seed),
    }


def randoms(batch, rays, num_coarse, num_fine, seed=7, noise_std=0.1):
    """The random tensors the reference draws, as explicit inputs (SURVEY.md section 8a quirk v):
    t_rand ~ U[0,1) [B,R,Sc] (model/nerf_trainer.py:137), noise ~ N(0,std^2) [B,R,S]
    (utils/nerf_util.py:49-56), u_rand ~ U[0,1) [B,R,num_fine] (utils/nerf_util.py:95)."""
    rs = np.random.RandomState(seed)
    nf_total = num_coarse // 2 + num_fine
    t_rand = rs.uniform(0, 1, size=(batch, rays, num_coarse)).astype(F32)
    n_coarse = rs.standard_normal((batch, rays, num_coarse)).astype(F32)
    u_rand = rs.uniform(0, 1, size=(batch, rays, num_fine)).astype(F32)
    n_fine = rs.standard_normal((batch, rays, nf_total)).astype(F32)
    return {
        "t_rand": t_rand, "u_rand": u_rand,
        "unit_coarse": n_coarse, "unit_fine": n_fine,          # what torch.randn returned
        "noise_coarse": (n_coarse * F32(noise_std)).astype(F32),  # ... * radiance_field_noise_std
        "noise_fine": (n_fine * F32(noise_std)).astype(F32),
    }


# ------------------------------------------------------------------------------------------------
# StyleUNet fixtures: order-independent, name-keyed parameter values so that the reference module (built by
# oracle/gen_golden.py) and havatar_b200.styleunet get bit-identical weights without shipping a state_dict
# ------------------------------------------------------------------------------------------------
_FIXED = (".kernel", ".ll", ".lh", ".hl", ".hh")


def cotangents(batch, rays, fine, seed=11, scale=1e-3):
    """Upstream gradients for the backward parity cases: d(loss)/d(rgb [B,R,67]), /d(depth), /d(acc) of each pass.  `scale`
    mimics a mean-reduced loss (small values exercise the loss scaling of the 16-bit gradient operands)."""
    rs = np.random.RandomState(seed)
    out = {}
    for p in (("coarse", "fine") if fine else ("coarse",)):
        out["rgb_" + p] = (rs.standard_normal((batch, rays, 67)) * scale).astype(F32)
        out["depth_" + p] = (rs.standard_normal((batch, rays)) * scale).astype(F32)
        out["acc_" + p] = (rs.standard_normal((batch, rays)) * scale).astype(F32)
    return out


def named_normal(name, shape, seed=0):
    import zlib

    rs = np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    return rs.standard_normal(tuple(shape)).astype(F32)


def styleunet_state(shapes, seed=0):
    """shapes: {state_dict key: shape}.  Returns {key: float32 array} for every learnable / random entry (the fixed FIR
    and Haar buffers are left at their constructor values).  Scales follow the reference initialisers."""
    out = {}
    for name, shape in shapes.items():
        if name.endswith(_FIXED):
            continue
        n = named_normal(name, shape, seed)
        if name.endswith("modulation.bias"):
            v = 1.0 + 0.1 * n                      # EqualLinear(bias_init=1), model/styleUnet.py:214
        elif name.endswith("noise.weight"):
            v = 0.1 * n                            # NoiseInjection.weight (zero-initialised in the reference; nonzero so it is exercised)
        elif (name.startswith("style.") or ".style." in name) and name.endswith(".weight"):
            v = 100.0 * n                          # randn / lr_mul, lr_mlp = 0.01 (styleUnet.py:131)
        elif name.endswith("bias"):
            v = 0.1 * n
        else:
            v = n
        out[name] = v.astype(F32)
    return out


def trainer_state(shapes, seed=0):
    """Name-keyed values for every entry of a Trainer state_dict (reference model/nerf_trainer.py:12-36): StyleGAN_zxc
    generators via styleunet_state, the radiance MLP via mlp_weights, VolumeDecoder convolutions ~N(0, 0.05^2),
    latent codes ~N(0, 0.1^2).  Box-warp and identity buffers keep their constructor values."""
    gen = {k: v for k, v in shapes.items() if ".XY_gen." in k or ".YZ_gen." in k}
    out = styleunet_state(gen, seed)
    mlp = mlp_weights(seed)
    for name, shape in shapes.items():
        if name in gen or name.endswith(("scale_factor", "trans_factor", "identity_trans")):
            continue
        short = name.replace("model_coarse.", "")
        if short in mlp:
            out[name] = mlp[short]
        elif name == "latent_codes":
            out[name] = (0.1 * named_normal(name, shape, seed)).astype(F32)
        elif name.endswith("init_lc"):
            out[name] = np.abs(named_normal(name, shape, seed)).astype(F32) % F32(1.0)
        elif name.endswith("bias"):
            out[name] = (0.05 * named_normal(name, shape, seed)).astype(F32)
        else:
            out[name] = (0.05 * named_normal(name, shape, seed)).astype(F32)
    return out
