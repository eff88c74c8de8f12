"""CPU restatement of the render hot path on torch CPU ops -- the same ATen operators the reference
itself executes (grid_sample, Linear, cumprod, searchsorted, sort), in its chunked form.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/render_oracle.py header).  This is what bench.py
times as `cpu_baseline` and under `--impl reference`: the reference tree cannot travel to the GPU
box, and its render path *is* a sequence of ATen calls, so this port costs what the reference costs
on the same host cores.  Pinned against the reference-minted goldens by tests/test_oracle_golden.py.
"""
import math

import torch
import torch.nn.functional as F

from .render_oracle import default_boxes  # noqa: F401  (same box constants)


def _grid3(vol, q):
    """utils/util.py:409-418 voxel_feature: vol [1,1,D,H,W], q [N,3] -> [N]."""
    return F.grid_sample(vol, q.view(1, 1, 1, -1, 3), mode="bilinear", padding_mode="border", align_corners=True).view(-1)


def _grid2(plane, q):
    """utils/util.py:395-406 sample_from_2dgrid: plane [1,C,H,W], q [N,2] -> [N,C]."""
    out = F.grid_sample(plane, q.view(1, 1, -1, 2), mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.view(plane.shape[1], -1).t()


def _pass(ro, rd, z, inv_T, planes, wvol, w, boxes, bg, noise):
    ps, pt, ss, st = boxes
    R, S = z.shape
    pts = (ro[:, None, :] + rd[:, None, :] * z[:, :, None]).reshape(-1, 3)               # nerf_trainer.py:141
    p1 = (pts + inv_T[3][None]) @ inv_T[:3, :3]                                          # Skinning_Field.py:83
    w0 = _grid3(wvol[:, 0:1], pts * ss + st)                                             # :85
    w1 = _grid3(wvol[:, 1:2], p1 * ss + st)
    den = w0 + w1 + 1e-8                                                                 # :87
    pc = (w0 / den)[:, None] * pts + (w1 / den)[:, None] * p1                            # :90,95
    q = pc * ps + pt                                                                     # util.py:232-236
    f0 = _grid2(planes[0:1], q[:, [0, 1]])                                               # util.py:378
    f1 = _grid2(planes[1:2], q[:, [2, 1]])                                               # util.py:381
    feat = torch.stack([f0, f1], dim=-1).reshape(pc.shape[0], -1)                        # util.py:388
    freqs = 2.0 ** torch.linspace(0.0, 7.0, 8, device=pc.device)                         # embedder.py:32-61
    ang = pc[:, None, :] * freqs[None, :, None]
    pe = torch.sin(torch.stack([ang, ang + math.pi / 2], dim=-2)).reshape(pc.shape[0], -1)
    x = torch.cat([feat, pe], dim=-1)                                                    # nerf_model.py:101-117
    x = F.relu(F.linear(x, w["layers_xyz.0.weight"], w["layers_xyz.0.bias"]))
    x = F.relu(F.linear(x, w["layers_xyz.1.weight"], w["layers_xyz.1.bias"]))
    alpha = F.linear(x, w["fc_alpha.weight"], w["fc_alpha.bias"])
    f = F.linear(x, w["fc_rgbFeat.weight"], w["fc_rgbFeat.bias"])
    rgb = F.linear(f, w["fc_rgb.weight"], w["fc_rgb.bias"])
    rf = torch.cat([rgb, f, alpha], dim=-1).view(R, S, 68)
    d = z[:, 1:] - z[:, :-1]                                                             # nerf_util.py:36-38
    d = torch.cat([d, d[:, -1:]], dim=-1) * rd.norm(p=2, dim=-1, keepdim=True)
    col = torch.cat([torch.sigmoid(rf[..., :3]), rf[..., 3:-1]], dim=-1)                 # :45-46
    a = rf[..., -1] if noise is None else rf[..., -1] + noise
    alpha = 1.0 - torch.exp(-F.relu(a) * d)                                              # :58-59
    T = torch.cumprod(1.0 - alpha + 1e-10, dim=-1)
    T = torch.cat([torch.ones_like(T[:, :1]), T[:, :-1]], dim=-1)                        # :19-23
    wts = alpha * T
    rgb_map = (wts[..., None] * col).sum(dim=-2)
    depth = (wts * z).sum(dim=-1)
    acc = wts.sum(dim=-1)
    if bg is not None:
        rgb_map = torch.cat([rgb_map[:, :3] + (1.0 - acc[:, None]) * bg, rgb_map[:, 3:]], dim=-1)   # :70-71
    return rgb_map, acc, wts, depth


def _sample_pdf(bins, weights, n, u_rand):
    """utils/nerf_util.py:76-117."""
    weights = weights + 1e-5
    cdf = torch.cumsum(weights / weights.sum(dim=-1, keepdim=True), dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], dim=-1)
    if u_rand is None:
        u = torch.linspace(0.0, 1.0, n, device=cdf.device).expand(cdf.shape[0], n)
    else:
        s = 1 / n
        u = (torch.arange(n, device=cdf.device) * s)[None] + u_rand * (s - 1e-6)
    u = u.contiguous()
    inds = torch.searchsorted(cdf.contiguous(), u, right=True)
    below, above = (inds - 1).clamp(min=0), inds.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = cdf.gather(1, below), cdf.gather(1, above)
    b0, b1 = bins.gather(1, below), bins.gather(1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return b0 + (u - c0) / denom * (b1 - b0)


def render_rays_grad(ray_batch, background_prior, inv_head_T, planes, wvol, weights, boxes, num_coarse, num_fine=0,
                     cotangents=None, device="cpu", **rand):
    """Gradients of  sum_k <cotangents[k], out[k]>  with respect to planes, wvol and the MLP tensors, by torch autograd over
    the ATen call sequence above -- what loss.backward() does in the reference (train_avatar.py:149).  z_samples is detached
    like the reference does (model/nerf_trainer.py:167).  -> (outputs dict, grads dict), numpy."""
    T = lambda a: torch.as_tensor(a, dtype=torch.float32).to(device)
    leaves = {"planes": T(planes).requires_grad_(True), "wvol": T(wvol).requires_grad_(True)}
    w = {k: T(v).requires_grad_(True) for k, v in weights.items()}
    with torch.enable_grad():
        out = render_rays.__wrapped__(ray_batch, background_prior, inv_head_T, leaves["planes"], leaves["wvol"], w, boxes,
                                      num_coarse, num_fine, device=device, to_numpy=False, **rand)
        loss = 0.0
        for k, c in cotangents.items():
            if c is not None and out.get(k) is not None:
                loss = loss + (out[k] * T(c).reshape(out[k].shape)).sum()
        loss.backward()
    grads = {k: v.grad.cpu().numpy() for k, v in leaves.items()}
    grads.update({k: v.grad.cpu().numpy() for k, v in w.items()})
    return {k: (None if v is None else v.detach().cpu().numpy()) for k, v in out.items()}, grads


@torch.no_grad()
def render_rays(ray_batch, background_prior, inv_head_T, planes, wvol, weights, boxes, num_coarse, num_fine=0,
                t_rand=None, noise_coarse=None, u_rand=None, noise_fine=None, chunk=4096, device="cpu", to_numpy=True):
    """Same contract as oracle.render_oracle.render_rays, rays processed in chunks of `chunk` like the reference
    (model/nerf_trainer.py:65-71; nerf.validation.chunksize).  device="cuda" runs the very same ATen call sequence on the
    GPU in fp32 -- bench.py times that as the stand-in for the reference's own GPU path (the reference tree is Python and
    does not exist on the GPU box)."""
    T = lambda a: None if a is None else (a if isinstance(a, torch.Tensor) and a.requires_grad else torch.as_tensor(a, dtype=torch.float32).to(device))
    ray_batch, background_prior, inv_head_T, planes, wvol = map(T, (ray_batch, background_prior, inv_head_T, planes, wvol))
    t_rand, noise_coarse, u_rand, noise_fine = map(T, (t_rand, noise_coarse, u_rand, noise_fine))
    w = {k: T(v) for k, v in weights.items()}
    boxes = tuple(T(b) for b in boxes)
    B, R = ray_batch.shape[:2]
    keys = ("rgb_coarse", "depth_coarse", "acc_coarse", "weights_max", "rgb_fine", "depth_fine", "acc_fine")
    out = {k: [] for k in keys}
    tvals = torch.linspace(0.0, 1.0, num_coarse, device=device)
    for b in range(B):
        per = {k: [] for k in keys}
        for r0 in range(0, R, chunk):
            sl = slice(r0, min(R, r0 + chunk))
            ro, rd = ray_batch[b, sl, :3], ray_batch[b, sl, 3:6]
            near, far = ray_batch[b, sl, 6:7], ray_batch[b, sl, 7:8]
            bg = None if background_prior is None else background_prior[b, sl]
            z = near * (1.0 - tvals) + far * tvals                                        # nerf_trainer.py:129-130
            if t_rand is not None:                                                        # :132-139
                mids = 0.5 * (z[:, 1:] + z[:, :-1])
                upper, lower = torch.cat([mids, z[:, -1:]], -1), torch.cat([z[:, :1], mids], -1)
                z = lower + (upper - lower) * t_rand[b, sl]
            sel = lambda a: None if a is None else a[b, sl]
            rgb, acc, wts, depth = _pass(ro, rd, z, inv_head_T[b], planes[:, b], wvol, w, boxes, bg, sel(noise_coarse))
            per["rgb_coarse"].append(rgb), per["depth_coarse"].append(depth), per["acc_coarse"].append(acc)
            if num_fine > 0:
                zs = _sample_pdf(0.5 * (z[:, 1:] + z[:, :-1]), wts[:, 1:-1], num_fine, sel(u_rand))       # :166-167
                zf, _ = torch.sort(torch.cat([z[:, ::2], zs.detach()], dim=-1), dim=-1)                            # :170
                rgbf, accf, wf, depthf = _pass(ro, rd, zf, inv_head_T[b], planes[:, b], wvol, w, boxes, bg, sel(noise_fine))
                per["rgb_fine"].append(rgbf), per["depth_fine"].append(depthf), per["acc_fine"].append(accf)
                per["weights_max"].append(wf.max(dim=-1)[0])
            else:
                per["weights_max"].append(wts.max(dim=-1)[0])
        for k in keys:
            if per[k]:
                out[k].append(torch.cat(per[k], dim=0))
    if not to_numpy:
        return {k: (torch.stack(v) if v else None) for k, v in out.items()}
    return {k: (torch.stack(v).cpu().numpy() if v else None) for k, v in out.items()}
