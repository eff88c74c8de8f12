"""Host side of the fused volumetric render: torch tensors in, one C-ABI call, torch tensors out.

`render_rays` is Trainer.predict_and_render_radiance (reference model/nerf_trainer.py:120-201) over the
whole ray batch in ONE kernel launch -- the reference's 4096-ray chunk loop (nerf_trainer.py:65-71)
only exists to bound the per-sample tensors it materialises in HBM; the fused kernel has none.
PyTorch is used for device memory and streams only.  No CPU fallback: CPU tensors raise.
"""
import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from . import _lib

RenderOut = namedtuple("RenderOut", "rgb_coarse depth_coarse acc_coarse weights_max rgb_fine depth_fine acc_fine z_fine pdf_inds",
                       defaults=(None,))

MLP_KEYS = ("layers_xyz.0.weight", "layers_xyz.0.bias", "layers_xyz.1.weight", "layers_xyz.1.bias",
            "fc_alpha.weight", "fc_alpha.bias", "fc_rgbFeat.weight", "fc_rgbFeat.bias", "fc_rgb.weight", "fc_rgb.bias")
_MLP_FIELDS = ("w0", "b0", "w1", "b1", "w_alpha", "b_alpha", "w_feat", "b_feat", "w_rgb", "b_rgb")
_MLP_SHAPES = ((128, 176), (128,), (128, 128), (128,), (1, 128), (1,), (64, 128), (64,), (3, 64), (3,))

_workspaces = {}  # (device index, stream) -> uint8 tensor, grown on demand


def box_warp_param(xb, yb, zb):
    """reference utils/util.py:179-186 get_box_warp_param -> (scales, trans)."""
    out_s, out_t = [], []
    for lo, hi in (xb, yb, zb):
        f = 2.0 / (hi - lo)
        out_s.append(f)
        out_t.append(-(f * (lo + hi) * 0.5))
    return tuple(out_s), tuple(out_t)


def default_boxes(xyz_bounding=((-1.5, 1.5), (-1.6, 1.4), (-1.6, 1.2))):
    """(plane_scale, plane_trans, skin_scale, skin_trans): plane box = models.coarse.XYZ_bounding
    (reference config/singleview_512_base.yml:51), skin box = same with Y[0] = 0.3*Y[1] (nerf_trainer.py:29-34)."""
    xb, yb, zb = [tuple(float(v) for v in b) for b in xyz_bounding]
    ps, pt = box_warp_param(xb, yb, zb)
    ss, st = box_warp_param(xb, (0.3 * yb[1], yb[1]), zb)
    return ps, pt, ss, st


def _f32c(t, name, shape=None):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.HavError("%s must be a CUDA tensor (havatar_b200 has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise _lib.HavError("%s must be float32, got %s" % (name, t.dtype))
    t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise _lib.HavError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return t


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def release_workspaces(stream=None):
    """Drop the cached forward / backward workspaces of `stream` (a torch.cuda.Stream; None = all of them).  Workspaces are
    cached per (device, stream) and live until released: owners of private streams call this when they are done."""
    for cache in (_workspaces, _bwd_workspaces):
        for key in list(cache):
            if stream is None or key == (stream.device.index, stream.cuda_stream):
                del cache[key]


def render_rays(ray_batch, background_prior, inv_head_T, planes, wvol, weights, num_coarse, num_fine=0,
                boxes=None, t_rand=None, noise_coarse=None, u_rand=None, noise_fine=None, precision="fp16",
                want_z_fine=False, reuse_packed=False, out=None, return_ctx=False, camera=None, img_hw=None, pixel_index=None,
                want_pdf_inds=False, check_range=False, cta_pairs=False):
    """ray_batch [B,R,8] (o3 d3 near far; extra trailing columns such as the reference's viewdirs are
    ignored) -- or ray_batch=None with camera [B,18] (fx fy cx cy | c2w [3,4] row-major | near far, see `camera_block`),
    img_hw=(H, W) and optionally pixel_index [B,R] int32 (y * W + x; default: all H*W pixels in row-major order): the rays are
    then generated inside the kernel (dataloader/data_util.py:28-56) and no per-ray tensor is uploaded.
    check_range=True (precision 'fp16'): raise HavError when a plane texel / MLP weight does not fit fp16 or a hidden activation
    saturates at 65504 (HAV_RENDER_CHECK_RANGE; costs one 4-byte device->host read) instead of returning clipped values --
    models with such magnitudes need precision='bf16'.  precision 'fp16x3' = split-precision tensor-core mode (fp32-class
    results, HAV_PREC_FP16X3); cta_pairs=True runs the 16-bit modes on the CTA-pair kernel (HAV_RENDER_CTA_PAIRS).  background_prior [B,R,3] or None, inv_head_T [B,4,3], planes [2,B,64,H,W], wvol [1,2,D,H,W],
    weights: mapping with the reference's model_coarse keys (MLP_KEYS).  Random draws are explicit inputs
    (SURVEY.md section 8a quirk v): t_rand [B,R,Sc], noise_* [B,R,S] already scaled by the noise std,
    u_rand [B,R,num_fine]; None switches that randomness off (u_rand None == sample_pdf det=True).
    reuse_packed=True skips re-packing weights/planes (HAV_RENDER_REUSE_PACKED: same weights, planes, precision
    and batch as the previous call on this stream).  out = a previous RenderOut to write into (no allocation).
    return_ctx=True additionally returns the call's argument block (a RenderCtx that keeps every tensor alive) for
    render_backward.  Returns RenderOut of [B,R,*] tensors (fine slots None when num_fine == 0)."""
    L = _lib.lib()
    if camera is not None:
        camera = _f32c(camera, "camera")
        if camera.dim() != 2 or camera.shape[1] != 18 or img_hw is None:
            raise _lib.HavError("camera must be [B,18] and img_hw=(H, W) must be given")
        dev, B = camera.device, int(camera.shape[0])
        if pixel_index is not None:
            if not pixel_index.is_cuda or pixel_index.dtype != torch.int32 or pixel_index.dim() != 2 or pixel_index.shape[0] != B:
                raise _lib.HavError("pixel_index must be a CUDA int32 tensor [B,R]")
            pixel_index = pixel_index.contiguous()
            R = int(pixel_index.shape[1])
        else:
            R = int(img_hw[0]) * int(img_hw[1])
        ray_batch = None
    else:
        if ray_batch is None or ray_batch.dim() != 3 or ray_batch.shape[-1] < 8:
            raise _lib.HavError("ray_batch must be [B,R,>=8]")
        if ray_batch.shape[-1] > 8:
            ray_batch = ray_batch[..., :8]
        ray_batch = _f32c(ray_batch, "ray_batch")
        dev = ray_batch.device
        B, R = int(ray_batch.shape[0]), int(ray_batch.shape[1])
    planes = _f32c(planes, "planes")
    if planes.dim() != 5 or planes.shape[0] != 2 or planes.shape[1] != B:
        raise _lib.HavError("planes must be [2,B,C,H,W] with B == ray_batch.shape[0]")
    wvol = _f32c(wvol, "wvol")
    if wvol.dim() != 5 or wvol.shape[0] != 1 or wvol.shape[1] != 2:
        raise _lib.HavError("wvol must be [1,2,D,H,W]")
    Sf = (num_coarse + 1) // 2 + num_fine if num_fine > 0 else 0
    a = _lib.RenderArgs()
    a.struct_bytes = C.sizeof(_lib.RenderArgs)
    a.precision = _lib.PRECISIONS[precision]
    check_range = bool(check_range) and precision == "fp16"
    a.flags = (1 if reuse_packed else 0) | (2 if check_range else 0) | (4 if cta_pairs and precision in ('fp16', 'bf16') else 0)
    a.batch, a.rays, a.num_coarse, a.num_fine = B, R, int(num_coarse), int(num_fine)
    a.plane_c, a.plane_h, a.plane_w = int(planes.shape[2]), int(planes.shape[3]), int(planes.shape[4])
    a.vol_d, a.vol_h, a.vol_w = int(wvol.shape[2]), int(wvol.shape[3]), int(wvol.shape[4])
    ps, pt, ss, st = boxes if boxes is not None else default_boxes()
    for i in range(3):
        a.plane_scale[i], a.plane_trans[i] = float(ps[i]), float(pt[i])
        a.skin_scale[i], a.skin_trans[i] = float(ss[i]), float(st[i])
    keep = [ray_batch, planes, wvol, camera, pixel_index]
    a.ray_batch, a.planes, a.wvol = _ptr(ray_batch), _ptr(planes), _ptr(wvol)
    if camera is not None:
        a.camera, a.pixel_index = _ptr(camera), _ptr(pixel_index)
        a.img_h, a.img_w = int(img_hw[0]), int(img_hw[1])
    bg = _f32c(background_prior, "background_prior", (B, R, 3))
    ihT = _f32c(inv_head_T, "inv_head_T", (B, 4, 3))
    a.background, a.inv_head_T = _ptr(bg), _ptr(ihT)
    keep += [bg, ihT]
    for key, field, shape in zip(MLP_KEYS, _MLP_FIELDS, _MLP_SHAPES):
        w = _f32c(weights[key], key, shape)
        keep.append(w)
        setattr(a, field, _ptr(w))
    for name, t, shape in (("t_rand", t_rand, (B, R, num_coarse)), ("noise_coarse", noise_coarse, (B, R, num_coarse)),
                           ("u_rand", u_rand, (B, R, num_fine)), ("noise_fine", noise_fine, (B, R, Sf))):
        t = _f32c(t, name, shape)
        keep.append(t)
        setattr(a, name, _ptr(t))
    new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    prev = out
    out = dict(rgb_coarse=new(B, R, 67), depth_coarse=new(B, R, 1), acc_coarse=new(B, R, 1), weights_max=new(B, R, 1),
               rgb_fine=None, depth_fine=None, acc_fine=None, z_fine=None, pdf_inds=None)
    if num_fine > 0:
        out.update(rgb_fine=new(B, R, 67), depth_fine=new(B, R, 1), acc_fine=new(B, R, 1))
        if want_z_fine:
            out["z_fine"] = new(B, R, Sf)
        if want_pdf_inds:
            out["pdf_inds"] = torch.empty((B, R, num_fine), dtype=torch.int32, device=dev)
    if prev is not None:
        for k in out:
            v = getattr(prev, k)
            if (v is None) != (out[k] is None) or (v is not None and (v.shape != out[k].shape or v.device != dev)):
                raise _lib.HavError("out.%s does not match this call" % k)
            out[k] = v
    for k, v in out.items():
        setattr(a, k, _ptr(v))
    status = None
    if check_range:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        a.range_status = _ptr(status)
        keep.append(status)
    with torch.cuda.device(dev):
        need = int(L.hav_render_workspace_bytes(C.byref(a)))
        if need == 0 and B * R > 0:
            # re-run the checks through the real entry point to get the specific error code
            _lib.check(L.hav_render_forward(C.byref(a), None) or -2, "hav_render_forward")
        ws = _workspace(dev, need)
        a.workspace, a.workspace_bytes = C.c_void_p(ws.data_ptr()), ws.numel()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.hav_render_forward(C.byref(a), C.c_void_p(stream)), "hav_render_forward")
    if status is not None:
        bits = int(status.item())
        if bits:
            what = [w for b, w in ((1, "a plane texel or MLP weight exceeds 65504"), (2, "a hidden activation saturated at 65504")) if bits & b]
            raise _lib.HavError("fp16 operand range exceeded (%s): render with precision='bf16'" % "; ".join(what))
    res = RenderOut(**out)
    if return_ctx:
        return res, RenderCtx(a, keep, res, dev)
    del keep
    return res


class RenderCtx:
    """Argument block of one hav_render_forward call + the input tensors its pointers refer to (kept alive) + its outputs
    (`out`; the autograd node drops this reference and keeps the outputs through save_for_backward instead)."""

    def __init__(self, args, keep, out, dev):
        self.args, self.keep, self.out, self.dev = args, keep, out, dev


_bwd_workspaces = {}


def render_backward(ctx, g_rgb_coarse=None, g_depth_coarse=None, g_acc_coarse=None, g_rgb_fine=None, g_depth_fine=None,
                    g_acc_fine=None, grad_scale=0.0):
    """Gradients of the render_rays call described by `ctx` (render_rays(..., want_z_fine=True, return_ctx=True)) given the
    upstream gradients of its outputs -- what loss.backward() computes through predict_and_render_radiance in the reference
    (model/nerf_trainer.py:120-201, train_avatar.py:149).  One hav_render_backward call.  Returns a dict: 'planes'
    [2,B,64,H,W], 'wvol' [1,2,D,H,W] and the ten MLP_KEYS tensors."""
    L = _lib.lib()
    a, dev = ctx.args, ctx.dev
    B, R = int(a.batch), int(a.rays)
    if a.precision == _lib.PREC_FP32:
        raise _lib.HavError("render_backward needs a 16-bit tensor-core forward (precision 'fp16' or 'bf16')")
    if a.num_fine > 0 and not a.z_fine:
        raise _lib.HavError("render_backward needs the forward's z_fine (render_rays(..., want_z_fine=True))")
    b = _lib.RenderBwdArgs()
    b.struct_bytes = C.sizeof(_lib.RenderBwdArgs)
    b.grad_scale = float(grad_scale)
    b.fwd = C.pointer(a)
    keep = []
    for name, g, shape in (("g_rgb_coarse", g_rgb_coarse, (B, R, 67)), ("g_depth_coarse", g_depth_coarse, (B, R, 1)),
                           ("g_acc_coarse", g_acc_coarse, (B, R, 1)), ("g_rgb_fine", g_rgb_fine, (B, R, 67)),
                           ("g_depth_fine", g_depth_fine, (B, R, 1)), ("g_acc_fine", g_acc_fine, (B, R, 1))):
        if g is not None:
            g = _f32c(g.reshape(shape) if g.numel() == B * R * shape[-1] else g, name, shape)
            keep.append(g)
        setattr(b, name, _ptr(g))
    new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    grads = {"planes": new(2, B, int(a.plane_c), int(a.plane_h), int(a.plane_w)),
             "wvol": new(1, 2, int(a.vol_d), int(a.vol_h), int(a.vol_w))}
    b.g_planes, b.g_wvol = _ptr(grads["planes"]), _ptr(grads["wvol"])
    for key, field, shape in zip(MLP_KEYS, _MLP_FIELDS, _MLP_SHAPES):
        grads[key] = new(*shape)
        setattr(b, "g_" + field, _ptr(grads[key]))
    with torch.cuda.device(dev):
        need = int(L.hav_render_backward_workspace_bytes(C.byref(b)))
        if need == 0 and B * R > 0:
            _lib.check(L.hav_render_backward(C.byref(b), None) or -2, "hav_render_backward")
        key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
        ws = _bwd_workspaces.get(key)
        if ws is None or ws.numel() < need:
            ws = None
            _bwd_workspaces.pop(key, None)
            ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=dev)
            _bwd_workspaces[key] = ws
        b.workspace, b.workspace_bytes = C.c_void_p(ws.data_ptr()), ws.numel()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.hav_render_backward(C.byref(b), C.c_void_p(stream)), "hav_render_backward")
    del keep
    return grads


class _RenderFunction(torch.autograd.Function):
    """autograd node: forward = hav_render_forward, backward = hav_render_backward."""

    @staticmethod
    def forward(ctx, opts, ray_batch, background_prior, inv_head_T, planes, wvol, *mlp):
        weights = dict(zip(MLP_KEYS, mlp))
        out, rctx = render_rays(ray_batch, background_prior, inv_head_T, planes, wvol, weights, want_z_fine=True,
                                return_ctx=True, **opts)
        ctx.rctx = rctx
        rctx.keep.append(out.z_fine)
        rctx.out = None
        # backward re-reads planes / wvol / the MLP tensors through raw pointers: remember their versions so that an in-place
        # update between forward and backward (an optimiser step, EMA, load_state_dict) raises like stock autograd would
        ctx.input_versions = [(t, t._version) for t in (planes, wvol) + tuple(mlp)]
        ctx.save_for_backward(*[t for t in out[:7] if t is not None])   # keeps the storages hav_render_backward reads alive
        ctx.mark_non_differentiable(out.weights_max)
        res = [out.rgb_coarse, out.depth_coarse, out.acc_coarse, out.weights_max]
        if out.rgb_fine is not None:
            res += [out.rgb_fine, out.depth_fine, out.acc_fine]
        return tuple(res)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *g):
        for t, v in ctx.input_versions:
            if t._version != v:
                raise RuntimeError("havatar_b200 render: an input needed for the backward pass (%s) was modified in place after "
                                   "the forward (version %d -> %d)" % (tuple(t.shape), v, t._version))
        cont = lambda t: None if t is None else t.contiguous()
        kw = dict(g_rgb_coarse=cont(g[0]), g_depth_coarse=cont(g[1]), g_acc_coarse=cont(g[2]))
        if len(g) > 4:
            kw.update(g_rgb_fine=cont(g[4]), g_depth_fine=cont(g[5]), g_acc_fine=cont(g[6]))
        saved = ctx.saved_tensors   # noqa: F841  (same storages as the pointers inside ctx.rctx.args)
        grads = render_backward(ctx.rctx, **kw)
        del saved
        return (None, None, None, None, grads["planes"], grads["wvol"]) + tuple(grads[k] for k in MLP_KEYS)


def render_rays_autograd(ray_batch, background_prior, inv_head_T, planes, wvol, weights, num_coarse, num_fine=0, boxes=None,
                         t_rand=None, noise_coarse=None, u_rand=None, noise_fine=None, precision="fp16", camera=None, img_hw=None,
                         pixel_index=None):
    """render_rays as a differentiable torch op: gradients flow to `planes`, `wvol` and the MLP tensors in `weights`
    (the learnable inputs of predict_and_render_radiance); rays, poses and random draws are constants.
    Returns RenderOut (z_fine slot None)."""
    if precision == "fp32":
        raise _lib.HavError("the differentiable render runs on the tensor-core path: precision 'fp16' or 'bf16'")
    opts = dict(num_coarse=num_coarse, num_fine=num_fine, boxes=boxes, t_rand=t_rand, noise_coarse=noise_coarse,
                u_rand=u_rand, noise_fine=noise_fine, precision=precision, camera=camera, img_hw=img_hw, pixel_index=pixel_index)
    res = _RenderFunction.apply(opts, ray_batch, background_prior, inv_head_T, planes, wvol, *[weights[k] for k in MLP_KEYS])
    res = tuple(res) + (None,) * (7 - len(res))
    return RenderOut(*res, None, None)


def sample_pdf(bins, weights, num_samples, u=None):
    """utils/nerf_util.py:76-117 on the device function the render kernels run (hav_sample_pdf): bins [N,M], weights [N,M-1],
    u [N,num_samples] uniform draws or None (det=True).  -> (samples [N,num_samples], inds [N,num_samples] int32)."""
    L = _lib.lib()
    bins, weights, u = _f32c(bins, "bins"), _f32c(weights, "weights"), _f32c(u, "u")
    n, m = int(bins.shape[0]), int(bins.shape[1])
    if tuple(weights.shape) != (n, m - 1):
        raise _lib.HavError("weights must be [N, M-1]")
    samples = torch.empty((n, num_samples), dtype=torch.float32, device=bins.device)
    inds = torch.empty((n, num_samples), dtype=torch.int32, device=bins.device)
    scratch = torch.empty((n, m + 1), dtype=torch.float32, device=bins.device)
    with torch.cuda.device(bins.device):
        _lib.check(L.hav_sample_pdf(_ptr(bins), _ptr(weights), _ptr(u), n, m, int(num_samples), _ptr(samples), _ptr(inds),
                                    _ptr(scratch), C.c_void_p(torch.cuda.current_stream(bins.device).cuda_stream)), "hav_sample_pdf")
    return samples, inds


def camera_block(intr, c2w, near, far, device="cuda"):
    """The [B,18] camera tensor of render_rays(camera=...): intr [B,4] or [4] (fx, fy, cx, cy), c2w [B,3,4] / [B,4,4] / [3,4],
    near / far scalars or [B] (dataloader/dataloader.py:174-177: |T_ori[:3,3]| + {near,far} * length).  72 bytes per frame
    instead of the 8.4 MB ray tensor of a 512 x 512 frame."""
    intr = np.asarray(intr, dtype=np.float32).reshape(-1, 4)
    c2w = np.asarray(c2w, dtype=np.float32)
    c2w = c2w.reshape((-1,) + c2w.shape[-2:])[:, :3, :4].reshape(-1, 12)
    B = max(intr.shape[0], c2w.shape[0])
    nf = np.stack([np.broadcast_to(np.asarray(near, dtype=np.float32).reshape(-1), (B,)),
                   np.broadcast_to(np.asarray(far, dtype=np.float32).reshape(-1), (B,))], axis=1)
    cam = np.concatenate([np.broadcast_to(intr, (B, 4)), np.broadcast_to(c2w, (B, 12)), nf], axis=1).astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(cam)).to(device)


def get_rays(height, width, intr, c2w, near, far, device="cuda"):
    """Ray generation on the device (reference dataloader/data_util.py:28-56 + dataloader.py:174-180):
    intr = (fx, fy, cx, cy), c2w [3,4] (or [4,4]) array-like on the host.  -> ray_batch [H*W, 8]."""
    L = _lib.lib()
    dev = torch.device(device)
    out = torch.empty((height * width, 8), dtype=torch.float32, device=dev)
    intr_c = (C.c_float * 4)(*[float(v) for v in np.asarray(intr, dtype=np.float32).reshape(-1)[:4]])
    c2w_c = (C.c_float * 12)(*[float(v) for v in np.asarray(c2w, dtype=np.float32)[:3, :4].reshape(-1)])
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.hav_get_rays(_ptr(out), int(height), int(width), intr_c, c2w_c, float(near), float(far),
                                  C.c_void_p(stream)), "hav_get_rays")
    return out


class HostRenderer:
    """Host-buffer front end of render_rays: the per-frame inputs arrive in (pinned) host memory the way the
    reference's dataloader hands them over (train_avatar.py:108-120 `.to(device)` per batch) and the rendered maps
    are returned in pinned host memory (avatarHD_reenactment.py:168 `.cpu()`).  Model constants (MLP weights,
    skinning-weight volume) stay resident on the device.  Device staging and pinned result buffers are allocated
    once and reused, so a call is: H2D copies -> one hav_render_forward -> D2H copies -> stream sync."""

    def __init__(self, weights, wvol, num_coarse, num_fine=0, boxes=None, precision="fp16", device="cuda"):
        self.dev = torch.device(device)
        self.weights = {k: torch.as_tensor(v).to(self.dev, torch.float32).contiguous() for k, v in weights.items()}
        self.wvol = torch.as_tensor(wvol).to(self.dev, torch.float32).contiguous()
        self.num_coarse, self.num_fine, self.boxes, self.precision = num_coarse, num_fine, boxes, precision
        self._dev_in, self._dev_out, self._host_out = {}, None, None
        self.h2d_bytes = self.d2h_bytes = 0

    def _stage(self, name, host):
        buf = self._dev_in.get(name)
        if buf is None or buf.shape != host.shape:
            buf = torch.empty(host.shape, dtype=torch.float32, device=self.dev)
            self._dev_in[name] = buf
        buf.copy_(host, non_blocking=True)
        self.h2d_bytes += host.numel() * 4
        return buf

    def __call__(self, ray_batch, background_prior, inv_head_T, planes, **rand):
        """All arguments are float32 CPU tensors (pinned for async copies).  Returns a dict of pinned CPU tensors."""
        self.h2d_bytes = self.d2h_bytes = 0
        args = [self._stage(n, t) for n, t in (("ray_batch", ray_batch), ("background_prior", background_prior),
                                               ("inv_head_T", inv_head_T), ("planes", planes))]
        rnd = {k: self._stage(k, v) for k, v in rand.items() if v is not None}
        self._dev_out = render_rays(*args, self.wvol, self.weights, self.num_coarse, self.num_fine, boxes=self.boxes,
                                    precision=self.precision, out=self._dev_out, **rnd)
        if self._host_out is None or self._host_out["rgb_coarse"].shape != self._dev_out.rgb_coarse.shape:
            self._host_out = {k: torch.empty(v.shape, dtype=torch.float32, pin_memory=True)
                              for k, v in self._dev_out._asdict().items() if v is not None}
        for k, h in self._host_out.items():
            h.copy_(getattr(self._dev_out, k), non_blocking=True)
            self.d2h_bytes += h.numel() * 4
        torch.cuda.current_stream(self.dev).synchronize()
        return self._host_out


class PipelinedHostRenderer:
    """HostRenderer with the three legs of a step on three CUDA streams (H2D copy engine, compute, D2H copy engine)
    and two buffer sets, so that consecutive frames overlap: while frame i renders, frame i+1's inputs upload and
    frame i-1's maps download.  Every frame still pays its own host->device and device->host copies.

        r = PipelinedHostRenderer(weights, wvol, 64)
        for frame in frames:
            done = r.submit(**frame)          # returns the result of the frame submitted one call earlier (or None)
        last = r.drain()
    Results are dicts of pinned CPU tensors that stay valid until two more frames have been submitted."""

    DEPTH = 2

    def __init__(self, weights, wvol, num_coarse, num_fine=0, boxes=None, precision="fp16", device="cuda", maps="all"):
        """maps="all": every rendered map comes back (67-channel rgb + feature map, depth, acc, weights_max of each pass);
        maps="image": what the reference's validation loop reads back (train_avatar.py:182-218: rgb[..., :3], depth, acc) --
        the 64 feature channels are consumed on the device by the StyleUNet (avatarHD_reenactment.py:160-166)."""
        if maps not in ("all", "image"):
            raise ValueError("maps must be 'all' or 'image'")
        self.maps = maps
        self.dev = torch.device(device)
        self.weights = {k: torch.as_tensor(v).to(self.dev, torch.float32).contiguous() for k, v in weights.items()}
        self.wvol = torch.as_tensor(wvol).to(self.dev, torch.float32).contiguous()
        self.num_coarse, self.num_fine, self.boxes, self.precision = num_coarse, num_fine, boxes, precision
        self.s_h2d, self.s_comp, self.s_d2h = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.slots = [dict(inp={}, out=None, host=None, ev_in=torch.cuda.Event(), ev_comp=torch.cuda.Event(),
                           ev_out=torch.cuda.Event(), busy=False) for _ in range(self.DEPTH)]
        self.i = 0
        self.h2d_bytes = self.d2h_bytes = 0

    def submit(self, ray_batch=None, background_prior=None, inv_head_T=None, planes=None, camera=None, img_hw=None, **rand):
        """Host tensors of one frame.  Either ray_batch [B,R,8], or camera [B,18] + img_hw=(H, W): the rays are then generated
        inside the render kernel and the 32 B/ray ray tensor is never built nor uploaded."""
        sl = self.slots[self.i % self.DEPTH]
        prev = self.slots[(self.i - 1) % self.DEPTH] if self.i > 0 else None
        self.h2d_bytes = self.d2h_bytes = 0
        host_in = dict(background_prior=background_prior, inv_head_T=inv_head_T, planes=planes)
        host_in["camera" if camera is not None else "ray_batch"] = camera if camera is not None else ray_batch
        host_in.update({k: v for k, v in rand.items() if v is not None})
        with torch.cuda.stream(self.s_h2d):
            if sl["busy"]:
                self.s_h2d.wait_event(sl["ev_comp"])          # the render that last read these staging buffers
            for k, h in host_in.items():
                buf = sl["inp"].get(k)
                if buf is None or buf.shape != h.shape:
                    buf = sl["inp"][k] = torch.empty(h.shape, dtype=torch.float32, device=self.dev)
                buf.copy_(h, non_blocking=True)
                self.h2d_bytes += h.numel() * 4
            sl["ev_in"].record(self.s_h2d)
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(sl["ev_in"])
            if sl["busy"]:
                self.s_comp.wait_event(sl["ev_out"])          # the download that last read these output buffers
            a = sl["inp"]
            cam = dict(camera=a["camera"], img_hw=img_hw) if camera is not None else {}
            sl["out"] = render_rays(a.get("ray_batch"), a["background_prior"], a["inv_head_T"], a["planes"], self.wvol,
                                    self.weights, self.num_coarse, self.num_fine, boxes=self.boxes, precision=self.precision,
                                    out=sl["out"], **cam, **{k: a[k] for k in rand if rand[k] is not None})
            send = {k: v for k, v in sl["out"]._asdict().items() if v is not None}
            if self.maps == "image":      # compact the colour channels on the device; the feature channels stay there
                send = {k: (v[..., :3].contiguous() if k.startswith("rgb") else v) for k, v in send.items() if k != "weights_max"}
            sl["send"] = send
            sl["ev_comp"].record(self.s_comp)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(sl["ev_comp"])
            if sl["host"] is None or any(sl["host"][k].shape != v.shape for k, v in send.items()):
                sl["host"] = {k: torch.empty(v.shape, dtype=torch.float32, pin_memory=True) for k, v in send.items()}
            for k, h in sl["host"].items():
                h.copy_(send[k], non_blocking=True)
                self.d2h_bytes += h.numel() * 4
            sl["ev_out"].record(self.s_d2h)
        sl["busy"] = True
        self.i += 1
        if prev is not None:
            prev["ev_out"].synchronize()
            return prev["host"]
        return None

    def drain(self):
        if self.i == 0:
            return None
        sl = self.slots[(self.i - 1) % self.DEPTH]
        sl["ev_out"].synchronize()
        return sl["host"]

    def close(self):
        """Release the render workspace cached for this renderer's private compute stream."""
        self.s_comp.synchronize()
        release_workspaces(self.s_comp)

    def __del__(self):
        try:
            release_workspaces(self.s_comp)
        except Exception:
            pass
