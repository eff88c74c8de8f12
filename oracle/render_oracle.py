"""CPU oracle for the HAvatar volumetric-render hot path (numpy, fp32).

TEST INFRASTRUCTURE ONLY -- a plain restatement of the reference algorithm used as the parity
checker.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The product path (havatar_b200/) never does: it fails loudly when the CUDA
library is missing.

Parity pin: the reference ships no tests or golden vectors, so this oracle
is pinned against outputs of the unmodified reference itself, executed on CPU in the build
container by oracle/gen_golden.py and committed under tests/golden/ (tests/test_oracle_golden.py).

Every function cites the reference file:line (relative to the reference root) it restates.
All arithmetic is float32 like the reference (no autocast anywhere).
"""
import math

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# box warps
# ----------------------------------------------------------------------------------------------
def box_warp_param(xb, yb, zb):
    """utils/util.py:179-186 get_box_warp_param -> (scales[3], trans[3]) python floats."""
    fx = 2 / (xb[1] - xb[0])
    cx = fx * (xb[0] + xb[1]) * 0.5
    fy = 2 / (yb[1] - yb[0])
    cy = fy * (yb[0] + yb[1]) * 0.5
    fz = 2 / (zb[1] - zb[0])
    cz = fz * (zb[0] + zb[1]) * 0.5
    return (float(fx), float(fy), float(fz)), (float(-cx), float(-cy), float(-cz))


def default_boxes(xyz_bounding=((-1.5, 1.5), (-1.6, 1.4), (-1.6, 1.2))):
    """Plane box = models.coarse.XYZ_bounding (config/singleview_512_base.yml:51, nerf_model.py:44);
    skin box = same with Y[0] = 0.3*Y[1] (model/nerf_trainer.py:29-34).
    Returns (plane_scale, plane_trans, skin_scale, skin_trans) as float32[3] arrays."""
    xb, yb, zb = [np.asarray(b, dtype=np.float64) for b in xyz_bounding]
    ps, pt = box_warp_param(xb, yb, zb)
    yb2 = yb.copy()
    yb2[0] = 0.3 * yb2[1]
    ss, st = box_warp_param(xb, yb2, zb)
    f = lambda t: np.asarray(t, dtype=F32)
    return f(ps), f(pt), f(ss), f(st)


# ----------------------------------------------------------------------------------------------
# grid sampling (ATen grid_sampler semantics for the exact arguments the reference uses)
# ----------------------------------------------------------------------------------------------
def _unnormalize(coord, size):
    # align_corners=True: ((x + 1) / 2) * (size - 1)
    return ((coord + F32(1.0)) / F32(2.0)) * F32(size - 1)


def trilinear_border(vol, xyz):
    """utils/util.py:409-418 voxel_feature -> F.grid_sample(5-D, bilinear, border, align_corners=True).
    vol: [D,H,W] float32 (one channel), xyz: [N,3] normalised coords (x->W, y->H, z->D). -> [N]."""
    D, H, W = vol.shape
    ix = np.clip(_unnormalize(xyz[:, 0], W), F32(0), F32(W - 1))
    iy = np.clip(_unnormalize(xyz[:, 1], H), F32(0), F32(H - 1))
    iz = np.clip(_unnormalize(xyz[:, 2], D), F32(0), F32(D - 1))
    x0f, y0f, z0f = np.floor(ix), np.floor(iy), np.floor(iz)
    x0, y0, z0 = x0f.astype(np.int64), y0f.astype(np.int64), z0f.astype(np.int64)
    fx1, fy1, fz1 = ix - x0f, iy - y0f, iz - z0f       # weight of the +1 corner
    fx0, fy0, fz0 = (x0f + F32(1)) - ix, (y0f + F32(1)) - iy, (z0f + F32(1)) - iz
    out = np.zeros(xyz.shape[0], dtype=F32)
    for dz, wz in ((0, fz0), (1, fz1)):
        for dy, wy in ((0, fy0), (1, fy1)):
            for dx, wx in ((0, fx0), (1, fx1)):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                ok = (xi <= W - 1) & (yi <= H - 1) & (zi <= D - 1)
                v = vol[np.minimum(zi, D - 1), np.minimum(yi, H - 1), np.minimum(xi, W - 1)]
                out += np.where(ok, v * (wx * wy * wz).astype(F32), F32(0)).astype(F32)
    return out


def bilinear_zeros(plane, xy):
    """utils/util.py:395-406 sample_from_2dgrid -> F.grid_sample(4-D, bilinear, zeros, align_corners=True).
    plane: [C,H,W], xy: [N,2] normalised (x->W, y->H). -> [N,C]."""
    C, H, W = plane.shape
    ix = _unnormalize(xy[:, 0], W)
    iy = _unnormalize(xy[:, 1], H)
    x0f, y0f = np.floor(ix), np.floor(iy)
    # keep indices finite/in int range for far-away points (weights there are masked anyway)
    x0 = np.clip(x0f, -2, W + 1).astype(np.int64)
    y0 = np.clip(y0f, -2, H + 1).astype(np.int64)
    wx1, wy1 = ix - x0f, iy - y0f
    wx0, wy0 = (x0f + F32(1)) - ix, (y0f + F32(1)) - iy
    out = np.zeros((xy.shape[0], C), dtype=F32)
    pl = np.ascontiguousarray(plane.transpose(1, 2, 0))  # [H,W,C]
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi, yi = x0 + dx, y0 + dy
            ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
            v = pl[np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)]       # [N,C]
            w = np.where(ok, (wx * wy).astype(F32), F32(0)).astype(F32)
            out += v * w[:, None]
    return out


# ----------------------------------------------------------------------------------------------
# the per-sample stages
# ----------------------------------------------------------------------------------------------
def skin_warp(pts, inv_T, wvol, skin_scale, skin_trans):
    """model/Skinning_Field.py:70-98 Deformation_Field_new.forward for one batch element.
    pts [N,3]; inv_T [4,3] (rows 0-2 rotation, row 3 translation); wvol [2,D,H,W].
    Bone 0 is the identity transform (Skinning_Field.py:50).  -> canonical pts [N,3]."""
    p0 = pts                                                        # (p + 0) @ I
    p1 = ((pts + inv_T[3][None, :]).astype(F32) @ inv_T[:3, :3]).astype(F32)   # :83
    w0 = trilinear_border(wvol[0], (p0 * skin_scale + skin_trans).astype(F32))  # :85
    w1 = trilinear_border(wvol[1], (p1 * skin_scale + skin_trans).astype(F32))
    den = (w0 + w1) + F32(1e-8)                                                  # :87
    w0n, w1n = w0 / den, w1 / den
    return (w0n[:, None] * p0 + w1n[:, None] * p1).astype(F32)                   # :90,95


def plane_features(pts_c, planes, plane_scale, plane_trans):
    """model/nerf_model.py:88-99 + utils/util.py:359-392: planes [2,C,H,W]; feature index = 2*c+plane."""
    q = (pts_c * plane_scale + plane_trans).astype(F32)           # utils/util.py:232-236
    f0 = bilinear_zeros(planes[0], q[:, [0, 1]])                  # util.py:378 (x,y)
    f1 = bilinear_zeros(planes[1], q[:, [2, 1]])                  # util.py:381 (z,y)
    return np.stack([f0, f1], axis=-1).reshape(pts_c.shape[0], -1)  # util.py:388, nerf_model.py:99


def positional_encode(x, num_freqs=8):
    """model/network/embedder.py:32-61: [N,3] -> [N, F*2*3], order [f][sin|cos][xyz]; cos = sin(a + pi/2)."""
    freqs = (F32(2.0) ** np.linspace(0.0, num_freqs - 1, num_freqs).astype(F32)).astype(F32)
    ang = (x[:, None, :] * freqs[None, :, None]).astype(F32)      # [N,F,3]
    feats = np.stack([ang, (ang + F32(math.pi / 2)).astype(F32)], axis=-2)  # [N,F,2,3]
    return np.sin(feats).astype(F32).reshape(x.shape[0], -1)


def mlp(pts_c, feat, w):
    """model/nerf_model.py:101-117 forward at sh_deg=0 -> [N,68] = rgb3 | feat64 | alpha1 (pre-activation)."""
    x = np.concatenate([feat, positional_encode(pts_c)], axis=-1)
    x = np.maximum(x @ w["layers_xyz.0.weight"].T + w["layers_xyz.0.bias"], F32(0)).astype(F32)
    x = np.maximum(x @ w["layers_xyz.1.weight"].T + w["layers_xyz.1.bias"], F32(0)).astype(F32)
    alpha = (x @ w["fc_alpha.weight"].T + w["fc_alpha.bias"]).astype(F32)
    f = (x @ w["fc_rgbFeat.weight"].T + w["fc_rgbFeat.bias"]).astype(F32)
    rgb = (f @ w["fc_rgb.weight"].T + w["fc_rgb.bias"]).astype(F32)
    return np.concatenate([rgb, f, alpha], axis=-1)


def composite(rf, z, rd, bg, noise=None):
    """utils/nerf_util.py:28-73 volume_render_radiance_field(act_feat=False) + cumprod_exclusive (:4-25).
    rf [R,S,68]; z [R,S]; rd [R,3]; bg [R,3] or None; noise [R,S] (already multiplied by std) or None.
    -> rgb [R,67], disp [R], acc [R], weights [R,S], depth [R]."""
    d = z[:, 1:] - z[:, :-1]
    d = np.concatenate([d, d[:, -1:]], axis=-1)
    d = (d * np.sqrt(np.sum(rd * rd, axis=-1, dtype=F32))[:, None]).astype(F32)   # :38
    rf = rf.copy()
    rf[..., :3] = F32(1) / (F32(1) + np.exp(-rf[..., :3]))                          # :45-46
    a = rf[..., -1] if noise is None else (rf[..., -1] + noise).astype(F32)
    sigma = np.maximum(a, F32(0))                                                   # :58
    alpha = (F32(1.0) - np.exp(-sigma * d)).astype(F32)                             # :59
    t = ((F32(1.0) - alpha) + F32(1e-10)).astype(F32)
    T = np.cumprod(t, axis=-1, dtype=F32)
    T = np.concatenate([np.ones_like(T[:, :1]), T[:, :-1]], axis=-1)              # :19-23
    w = (alpha * T).astype(F32)                                                     # :60
    rgb = np.sum(w[..., None] * rf[..., :-1], axis=-2, dtype=F32)                   # :62-63
    depth = np.sum(w * z, axis=-1, dtype=F32)                                       # :64-65
    acc = np.sum(w, axis=-1, dtype=F32)                                             # :67
    with np.errstate(divide="ignore", invalid="ignore"):
        disp = F32(1.0) / np.maximum(F32(1e-10), depth / acc)                       # :68
    if bg is not None:
        rgb[:, :3] = rgb[:, :3] + (F32(1.0) - acc[:, None]) * bg                    # :70-71
    return rgb, disp, acc, w, depth


def sample_pdf(bins, weights, num_samples, u_rand=None):
    """utils/nerf_util.py:76-117.  bins [R,M]; weights [R,M-1]; u_rand None => det=True (linspace),
    else the [R,num_samples] uniform draws the reference takes from torch.rand (:95).
    -> (samples [R,num_samples], inds [R,num_samples] int64 searchsorted(right=True) result)."""
    wts = (weights + F32(1e-5)).astype(F32)
    pdf = (wts / np.sum(wts, axis=-1, keepdims=True, dtype=F32)).astype(F32)
    cdf = np.cumsum(pdf, axis=-1, dtype=F32)
    cdf = np.concatenate([np.zeros_like(cdf[:, :1]), cdf], axis=-1)
    R = cdf.shape[0]
    if u_rand is None:
        u = np.broadcast_to(np.linspace(0.0, 1.0, num_samples).astype(F32), (R, num_samples))  # :87-91
    else:
        s = 1 / num_samples
        u = (np.arange(num_samples).astype(F32) * F32(s))[None, :]                      # :93-94
        u = (u + u_rand.astype(F32) * F32(s - 1e-6)).astype(F32)                   # :95
    inds = np.empty((R, num_samples), dtype=np.int64)
    for r in range(R):
        inds[r] = np.searchsorted(cdf[r], u[r], side="right")                      # :102
    below = np.maximum(0, inds - 1)
    above = np.minimum(cdf.shape[-1] - 1, inds)
    c0 = np.take_along_axis(cdf, below, axis=1)
    c1 = np.take_along_axis(cdf, above, axis=1)
    b0 = np.take_along_axis(bins, below, axis=1)
    b1 = np.take_along_axis(bins, above, axis=1)
    denom = c1 - c0
    denom = np.where(denom < F32(1e-5), F32(1), denom).astype(F32)                 # :112-113
    t = ((u - c0) / denom).astype(F32)
    return (b0 + t * (b1 - b0)).astype(F32), inds                                  # :114-115


# ----------------------------------------------------------------------------------------------
# the whole pass: model/nerf_trainer.py:120-201 predict_and_render_radiance
# ----------------------------------------------------------------------------------------------
def coarse_z(near, far, num_coarse, t_rand=None):
    """model/nerf_trainer.py:129-139.  near/far [R]; t_rand [R,S] uniform draws or None (perturb off)."""
    t = np.linspace(0.0, 1.0, num_coarse).astype(F32)
    z = (near[:, None] * (F32(1.0) - t) + far[:, None] * t).astype(F32)
    if t_rand is not None:
        mids = (F32(0.5) * (z[:, 1:] + z[:, :-1])).astype(F32)
        upper = np.concatenate([mids, z[:, -1:]], axis=-1)
        lower = np.concatenate([z[:, :1], mids], axis=-1)
        z = (lower + (upper - lower) * t_rand.astype(F32)).astype(F32)
    return z


def _one_pass(ro, rd, z, inv_T, planes, wvol, w, boxes, bg, noise):
    ps, pt, ss, st = boxes
    R, S = z.shape
    pts = (ro[:, None, :] + rd[:, None, :] * z[:, :, None]).astype(F32).reshape(-1, 3)   # :141
    pts_c = skin_warp(pts, inv_T, wvol, ss, st)                                           # :146
    feat = plane_features(pts_c, planes, ps, pt)                                          # :149
    rf = mlp(pts_c, feat, w).reshape(R, S, 68)                                            # :150-151
    return composite(rf, z, rd, bg, noise)                                                # :157-163


def render_rays(ray_batch, background_prior, inv_head_T, planes, wvol, weights, boxes,
                num_coarse, num_fine=0, t_rand=None, noise_coarse=None, u_rand=None, noise_fine=None):
    """One call of Trainer.predict_and_render_radiance over the whole ray batch (the reference's
    4096-ray chunk loop, nerf_trainer.py:65-71, is pure bookkeeping: rays are independent).

    ray_batch [B,R,8] = o3 d3 near far; background_prior [B,R,3]; inv_head_T [B,4,3];
    planes [2,B,C,H,W]; wvol [1,2,D,H,W]; weights: dict of the 5 nn.Linear tensors;
    t_rand [B,R,Sc] / noise_* [B,R,S] (pre-scaled by std) / u_rand [B,R,num_fine] or None.
    Returns dict with the slots of nerf_trainer.py:194-201."""
    B, R = ray_batch.shape[:2]
    out = {k: [] for k in ("rgb_coarse", "depth_coarse", "acc_coarse", "weights_max", "rgb_fine",
                           "depth_fine", "acc_fine", "z_fine", "pdf_inds", "weights_coarse")}
    for b in range(B):
        ro, rd = ray_batch[b, :, :3], ray_batch[b, :, 3:6]
        near, far = ray_batch[b, :, 6], ray_batch[b, :, 7]
        bg = None if background_prior is None else background_prior[b]
        z = coarse_z(near, far, num_coarse, None if t_rand is None else t_rand[b])
        rgb, _, acc, wts, depth = _one_pass(ro, rd, z, inv_head_T[b], planes[:, b], wvol[0], weights, boxes, bg,
                                            None if noise_coarse is None else noise_coarse[b])
        out["rgb_coarse"].append(rgb), out["depth_coarse"].append(depth), out["acc_coarse"].append(acc)
        out["weights_coarse"].append(wts)
        if num_fine > 0:
            z_mid = (F32(0.5) * (z[:, 1:] + z[:, :-1])).astype(F32)                                 # :166
            zs, inds = sample_pdf(z_mid, wts[:, 1:-1], num_fine, None if u_rand is None else u_rand[b])  # :167
            zf = np.sort(np.concatenate([z[:, ::2], zs], axis=-1), axis=-1)                          # :170
            rgbf, _, accf, wf, depthf = _one_pass(ro, rd, zf, inv_head_T[b], planes[:, b], wvol[0], weights, boxes, bg,
                                                  None if noise_fine is None else noise_fine[b])
            out["rgb_fine"].append(rgbf), out["depth_fine"].append(depthf), out["acc_fine"].append(accf)
            out["z_fine"].append(zf), out["pdf_inds"].append(inds)
            out["weights_max"].append(wf.max(axis=-1))
        else:
            out["weights_max"].append(wts.max(axis=-1))
    return {k: (np.stack(v) if v else None) for k, v in out.items()}


# ----------------------------------------------------------------------------------------------
# ray generation: dataloader/data_util.py:28-56 get_rays + dataloader/dataloader.py:174-180
# ----------------------------------------------------------------------------------------------
def get_rays(H, W, intr, c2w, near_far=None):
    """intr = (fx, fy, cx, cy): focal lengths in pixels, principal point as a fraction of the image
    size (data_util.py:38-39); c2w [3,4] or [4,4].  Ray r <-> pixel (y = r // W, x = r % W)
    (dataloader.py:72).  -> o [H*W,3], d [H*W,3] (normalised)."""
    fx, fy, cx, cy = [F32(v) for v in intr]
    K = np.array([[fx, 0, cx * F32(W)], [0, fy, cy * F32(H)], [0, 0, 1]], dtype=F32)
    j, i = np.meshgrid(np.arange(H, dtype=F32), np.arange(W, dtype=F32), indexing="ij")
    pix = np.stack([i, j, np.ones_like(i)], axis=-1).reshape(-1, 3)
    dirs = (pix @ np.linalg.inv(K).T.astype(F32)).astype(F32)
    d = (dirs @ np.asarray(c2w, dtype=F32)[:3, :3].T).astype(F32)
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(F32)
    o = np.broadcast_to(np.asarray(c2w, dtype=F32)[:3, 3], d.shape).copy()
    return o, d
