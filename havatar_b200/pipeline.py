"""The reference's HD inference frame on the B200 kernels (avatarHD_reenactment.py:149-170):

    condition renderings -> XY_gen / YZ_gen plane generators (model/nerf_model.py:58-86)
                         -> fused volumetric render of the full frame (model/nerf_trainer.py:94-118, mode 'validation')
                         -> SWGAN_unet upsampler on the 64 feature channels (avatarHD_reenactment.py:167)

AvatarHD owns the three networks + the radiance MLP weights and a skinning-weight volume; `frame()` is one frame,
`graphed()` returns a CUDA-graph replay of it for fixed shapes.  Inference only."""
import os

import torch

from . import render as hrender
from . import styleunet
from .graph import GraphedForward


_SIDE = {}
_FORK = {}      # device index -> (main, side) streams of the most recent two-stream plane generation


_NAMESPACE = [0]


class stream_namespace:
    """While active, aux_stream() hands out a separate family of side streams: two branches of the iteration that each fork
    their own sub-streams (train_step.StageTwoStep.dg_step) must not share them, or they would serialise on each other."""

    def __init__(self, ns):
        self.ns = int(ns)

    def __enter__(self):
        self.prev, _NAMESPACE[0] = _NAMESPACE[0], self.ns

    def __exit__(self, *exc):
        _NAMESPACE[0] = self.prev


def aux_stream(device, slot):
    """A persistent side stream of `device` (slot 0 = YZ plane generator, 1 = skinning-weight volume decoder, 2 = second
    discriminator pass, 3 = ToRGB pyramid, 4 = the G-step forward of the overlapped stage-two iteration, 5 = style plans)."""
    key = (device.index, slot + 16 * _NAMESPACE[0])
    st = _SIDE.get(key)
    if st is None:
        st = _SIDE[key] = torch.cuda.Stream(device)
    return st


def note_fork(device, *streams):
    cur = _FORK.get(device.index, ())
    _FORK[device.index] = tuple(dict.fromkeys(cur + streams))


def run_parallel(fn_main, fn_side, device, slot=2):
    """(fn_main(), fn_side()) with fn_side on a side stream: for independent sub-graphs whose kernels are too small to fill the
    GPU alone (e.g. the discriminator on the real and on the generated images).  Autograd replays each backward on the stream of
    its forward, so the two backward passes overlap as well."""
    device = torch.device(device)
    if device.type != "cuda":
        return fn_main(), fn_side()
    main = torch.cuda.current_stream(device)
    side = aux_stream(device, slot)
    note_fork(device, main, side)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        b = fn_side()
    a = fn_main()
    main.wait_stream(side)
    for t in (b if isinstance(b, (tuple, list)) else (b,)):
        if isinstance(t, torch.Tensor):
            t.record_stream(main)
    return a, b


def forked_streams(device):
    """Streams that may hold gradient work of the plane generators' backward on `device` (parallel.GradSync orders its
    collectives after all of them)."""
    return _FORK.get(device.index if device.index is not None else torch.cuda.current_device(), ())


def two_stream_planes(owner, lat, front, sides):
    """The two plane generators are independent (model/nerf_model.py:73-84 runs them back to back): run YZ_gen on a side stream
    while XY_gen runs on the current one.  Most of their layers (16^2 .. 64^2 maps) are latency-bound at batch 1, so the two
    streams fill each other's gaps; under CUDA-graph capture the fork / join becomes two parallel branches of the graph."""
    import os

    if os.environ.get("HAV_PLANES_ONE_STREAM"):          # A/B aid
        xy, _ = owner.XY_gen(lat, front.contiguous())
        yz, _ = owner.YZ_gen(lat, sides.contiguous())
        return torch.stack([xy, yz], dim=0)
    main = torch.cuda.current_stream(front.device)
    side = aux_stream(front.device, 0)
    sides = sides.contiguous()
    note_fork(front.device, main, side)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        yz, _ = owner.YZ_gen(lat, sides)
    xy, _ = owner.XY_gen(lat, front.contiguous())
    main.wait_stream(side)
    yz.record_stream(main)
    return torch.stack([xy, yz], dim=0)


class AvatarHD(torch.nn.Module):
    def __init__(self, mlp_weights, wvol, render_size=128, out_size=512, plane_res=128, cond_size=256, feat_dim=64, latent_dim=32,
                 num_coarse=64, num_fine=0, precision="fp16", boxes=None):
        super().__init__()
        style_in = latent_dim + 12                      # latent code | inv_head_T flattened (model/nerf_trainer.py:17-18,39)
        self.XY_gen = styleunet.StyleGAN_zxc(out_ch=feat_dim, out_size=plane_res, style_dim=style_in, middle_size=16, zero_latent=False,
                                             zero_noise=True, no_skip=True, n_mlp=4, inp_size=cond_size, inp_ch=7)
        self.YZ_gen = styleunet.StyleGAN_zxc(out_ch=feat_dim, out_size=plane_res, style_dim=style_in, middle_size=16, zero_latent=False,
                                             zero_noise=True, no_skip=True, n_mlp=4, inp_size=cond_size, inp_ch=13)
        self.upsampler = styleunet.SWGAN_unet(inp_size=render_size, inp_ch=feat_dim, out_ch=3, out_size=out_size, style_dim=64, n_mlp=4,
                                              middle_size=8)
        self.mlp = torch.nn.ParameterDict({k.replace(".", "__"): torch.nn.Parameter(torch.as_tensor(v).float(), requires_grad=False)
                                           for k, v in mlp_weights.items()})
        self.register_buffer("wvol", torch.as_tensor(wvol).float())
        self.render_size, self.num_coarse, self.num_fine, self.precision, self.boxes = render_size, num_coarse, num_fine, precision, boxes
        self._noise = None

    @torch.no_grad()
    def planes(self, latent_code, inv_head_T, front, left, right):
        """set_conditional_embedding (model/nerf_model.py:58-86): left view flipped along W, its mask channel dropped."""
        lat = [torch.cat([latent_code, inv_head_T.reshape(inv_head_T.shape[0], -1)], dim=-1)]
        left = left.flip(dims=[3])
        if left.shape[1] > 3:
            left = left[:, :-1]
        return two_stream_planes(self, lat, front, torch.cat([left, right], dim=1))

    @torch.no_grad()
    def frame(self, ray_batch, background, latent_code, inv_head_T, front, left, right, style):
        B, R = ray_batch.shape[:2]
        h = w = self.render_size
        planes = self.planes(latent_code, inv_head_T, front, left, right)
        weights = {k.replace("__", "."): v for k, v in self.mlp.items()}
        o = hrender.render_rays(ray_batch, background, inv_head_T, planes, self.wvol, weights, self.num_coarse, self.num_fine,
                                boxes=self.boxes, precision=self.precision)
        rgb = o.rgb_fine if self.num_fine > 0 else o.rgb_coarse
        maps = rgb.view(B, h, w, 67)                                                   # pixel r <-> ray r (nerf_trainer.py:111-113)
        if self._noise is None or self._noise[0].device != maps.device:
            self._noise = self.upsampler.make_noise(maps.device)
        if os.environ.get("HAV_HD_NCHW"):            # A/B aid: the reference's layout all the way (two permute copies, fp32 entry layers)
            render = maps.permute(0, 3, 1, 2).contiguous()
            return self.upsampler([style], render[:, 3:].contiguous(), noise=self._noise), render[:, :3]
        # the render's output IS channels-last: the 64 feature channels go to the upsampler as one fp16 [B,h,w,64] tensor (one
        # cast kernel), whose entry layers then run on the channels-last kernels (TMA-tiled blur, 16-byte staging loads)
        feats = maps[..., 3:].to(torch.float16).contiguous()
        image = self.upsampler([style], feats, noise=self._noise)                      # avatarHD_reenactment.py:167
        return image, maps[..., :3].permute(0, 3, 1, 2)

    def graphed(self, *example):
        return GraphedForward(lambda *a: self.frame(*a), *example)
