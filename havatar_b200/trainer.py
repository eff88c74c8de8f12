"""Drop-in for the reference's render orchestrator `model/nerf_trainer.py::Trainer` on the B200 kernels.

Same constructor (`Trainer(cfg, latent_codes_size=0, freeze_motion=True)`), same `forward(**data)` keys and return
tuples (model/nerf_trainer.py:94-118), same attributes the entry scripts touch (`latent_codes`, `model_coarse`,
`headpose_skin_net.{fix_canonical_W, canonical_Wvolume}`; avatarHD_reenactment.py:138-144), and the same state_dict
keys / shapes, so `load_state_dict(ckpt['nerf_render'])` works.  What differs is underneath:

  set_conditional_embedding  -> styleunet.StyleGAN_zxc on the tcgen05 convolution kernel
  nerf_forward's chunk loop + predict_and_render_radiance (:38-92, :120-201) -> ONE hav_render_forward over all rays
  the reference's torch.rand / torch.randn draws -> generated here in the reference's order and handed to the kernel

Under torch.no_grad() (inference / validation renders) everything runs on the fused inference kernels and outputs carry no
graph.  With gradients enabled (the training step, train_avatar.py:112-149) the render is the differentiable fused op
(hav_render_forward + hav_render_backward through render.render_rays_autograd): gradients reach the MLP, the bi-plane
features -> the plane generators and latent codes (styleunet_train.py), and the skinning-weight volume -> the VolumeDecoder.
The skinning-weight VolumeDecoder (model/network/voxel_encoder.py:150-179) is input independent; it is evaluated with
torch ONCE per forward (the reference re-evaluates it in every chunk and pass, Skinning_Field.py:79) and cached per weight
version when no gradient is needed."""
import math

import numpy as np
import torch
from torch import nn

from . import render as hrender
from . import styleunet


def _box_warp(xb, yb, zb):
    """utils/util.py:179-186."""
    out_s, out_t = [], []
    for lo, hi in (xb, yb, zb):
        f = 2.0 / (float(hi) - float(lo))
        out_s.append(f)
        out_t.append(-(f * (float(lo) + float(hi)) * 0.5))
    return out_s, out_t


class _BoxWarp(nn.Module):
    """UniformBoxWarp_new (utils/util.py:214-236): buffers scale_factor / trans_factor [1,3]."""

    def __init__(self, scales, trans):
        super().__init__()
        self.register_buffer("scale_factor", torch.tensor(scales, dtype=torch.float32).reshape(1, 3))
        self.register_buffer("trans_factor", torch.tensor(trans, dtype=torch.float32).reshape(1, 3))

    def forward(self, x):
        return x * self.scale_factor + self.trans_factor

    def inv_trans(self, x):
        """utils/util.py:221-229."""
        return (x - self.trans_factor.to(x.device)) * (1.0 / self.scale_factor.to(x.device))


def voxel_feature(xyz, volume, padding_mode="border"):
    """utils/util.py:409-418: trilinear fetch of [B,N,3] normalised points from a [B,C,D,H,W] volume -> [B,N,C] (plain torch:
    only the one-off skinning-volume pre-training and its visualisation use it; the per-sample fetch is in the render kernel)."""
    B, N, _ = xyz.shape
    feat = torch.nn.functional.grid_sample(volume, xyz.reshape(B, N, 1, 1, 3), mode="bilinear", padding_mode=padding_mode,
                                           align_corners=True)
    return feat[:, :, :, 0, 0].permute(0, 2, 1)


def make_volume_pts(steps=50, perturb=False, gridwarper=None):
    """utils/util.py:239-254: a steps^3 lattice over [-1,1]^3 (optionally jittered by up to one cell), mapped back to world
    space through the box warp."""
    c = torch.linspace(-1.0, 1.0, steps=steps, dtype=torch.float32)
    xv, yv, zv = torch.meshgrid(c, c.clone(), c.clone(), indexing="ij")
    pts = torch.stack([xv, yv, zv], dim=-1).reshape(-1, 3)
    if perturb:
        pts = pts + torch.rand_like(pts) * (2 / (steps - 1))
    return gridwarper.inv_trans(pts) if gridwarper is not None else pts


class _UpConv3D(nn.Module):
    """UpConv3DBlock(up_mode='upsample') (voxel_encoder.py:183-211): keys up.1.{weight,bias}."""

    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.Sequential(nn.Upsample(mode="trilinear", scale_factor=2, align_corners=False),
                                nn.Conv3d(cin, cout, kernel_size=3, padding=1, stride=1))
        self.norm = nn.InstanceNorm3d(cout, affine=False)

    def forward(self, x):
        return self.norm(self.up(x))


class VolumeDecoder(nn.Module):
    """voxel_encoder.py:150-179 (plain torch: evaluated once per weight update, not on the per-frame path)."""

    def __init__(self, num_in=1024, num_out=1, final_res=64):
        super().__init__()
        self.register_buffer("init_lc", torch.rand(1, num_in, 1, 1, 1))
        n, lg = int(math.log2(final_res)), int(math.log2(num_in))
        self.filters = nn.ModuleList([_UpConv3D(2 ** (lg - i), 2 ** (lg - i - 1)) for i in range(n)])
        self.final_conv = nn.Conv3d(2 ** (lg - n), num_out, bias=True, kernel_size=3, padding=1, stride=1)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.xavier_normal_(m.weight.data, gain=0.02)
                nn.init.constant_(m.bias.data, 0.0)

    def forward(self):
        x = self.init_lc
        for f in self.filters:
            x = torch.relu(f(x))
        x = torch.sigmoid(self.final_conv(x))
        return torch.cat([x, 1 - x], dim=1)


class SkinningField(nn.Module):
    """Deformation_Field_new (model/Skinning_Field.py:43-98) as a parameter / volume holder: the warp itself runs inside
    the fused render kernel."""

    def __init__(self, gridwarper):
        super().__init__()
        self.canonical_Wvolume = VolumeDecoder(num_in=1024, num_out=1, final_res=64)
        self.gridwarper = gridwarper
        self.register_buffer("identity_trans", torch.eye(4, dtype=torch.float32)[:, :-1])
        self.fix_canoW = False
        self.canonical_W = None
        self._key, self._vol, self.last_volume = None, None, None

    def fix_canonical_W(self):
        """Skinning_Field.py:57-62: freeze the volume and pin the head region to the head bone."""
        with torch.no_grad():
            w = self.canonical_Wvolume().detach()
            w[:, 1:, :, 0, :] = 1.0
            w[:, 1:, :1, : w.shape[-1] // 8, :] = 1.0
            self.canonical_W = torch.cat([1 - w[:, 1:], w[:, 1:]], dim=1).contiguous()
        self.fix_canoW = True

    def sample_volume(self, pts, padding_mode="border"):
        """Skinning_Field.py:65-68: bone-0 weight of world-space points [N,3] -> [N,1]."""
        vol = self.canonical_Wvolume()
        return voxel_feature(self.gridwarper(pts.unsqueeze(0)), vol[:, 0:1], padding_mode=padding_mode)[0]

    def pretrain_wc(self, num_iter=1, lr=1e-3, save_path=None, pose_space=False, vol_thr=None, progress=False):
        """Skinning_Field.py:101-125: fit the head-bone channel of the weight volume to the indicator of the head box `vol_thr`
        (binary cross-entropy on a jittered 20^3 lattice, Adam) -- what train_avatar.py:95 runs for 3000 iterations before a
        from-scratch stage-one training.  Plain torch: a one-off initialisation, not on the per-frame path."""
        if vol_thr is None:
            vol_thr = [[-0.5, 0.5], [-0.8, 0.5], [-0.3, 1.0]]
        opt = torch.optim.Adam([{"params": self.parameters()}], lr=lr)
        dev = self.identity_trans.device
        it = range(num_iter)
        if progress:
            from tqdm import tqdm
            it = tqdm(it)
        loss = None
        with torch.enable_grad():
            for _ in it:
                pts = make_volume_pts(steps=20, perturb=True, gridwarper=self.gridwarper).to(dev)
                inside = torch.ones(pts.shape[0], dtype=torch.bool, device=dev)
                for ax in range(3):
                    inside &= (pts[:, ax] > vol_thr[ax][0]) & (pts[:, ax] < vol_thr[ax][1])
                gt = inside.float().unsqueeze(-1)
                wc = self.canonical_Wvolume()
                pred = voxel_feature(self.gridwarper(pts.unsqueeze(0)), wc[:, 0:1] if pose_space else wc[:, 1:])
                loss = torch.nn.functional.binary_cross_entropy(torch.clamp(pred, 0.0, 1.0)[0], gt)
                loss.backward()
                opt.step()
                opt.zero_grad()
        self._key = None                                     # cached volume is stale
        styleunet.invalidate_caches()
        if save_path is not None:
            torch.save(self.canonical_Wvolume(), save_path)
        return None if loss is None else float(loss.detach())

    def visualize_motion_weight_vol(self, path):
        """Skinning_Field.py:127-132: a 20^3 point cloud coloured by the head-bone weight, written as a Wavefront .obj
        ('v x y z b g r' lines, utils/util.py:111-123)."""
        dev = self.identity_trans.device
        with torch.no_grad():
            pts = make_volume_pts(steps=20, perturb=False, gridwarper=self.gridwarper).to(dev)
            wc = self.canonical_Wvolume()
            w = voxel_feature(self.gridwarper(pts.unsqueeze(0)), wc[:, 1:])[0, :, 0].cpu().numpy()
            v = pts.cpu().numpy()
        with open(path, "w") as fp:
            for (x, y, z), c in zip(v, w):
                fp.write("v %f %f %f %f %f %f\n" % (x, y, z, c, c, c))

    def volume(self):
        """[1,2,D,H,W] skinning-weight volume for the kernel (Skinning_Field.py:79), cached per weight version."""
        if self.fix_canoW:
            return self.canonical_W
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.canonical_Wvolume.parameters()):
            self.last_volume = self.canonical_Wvolume()         # training: part of the graph (Skinning_Field.py:79)
            return self.last_volume
        key = tuple((p.data_ptr(), p._version) for p in self.canonical_Wvolume.parameters()) + (styleunet._EPOCH[0],)
        if key != self._key:
            with torch.no_grad():
                self._vol = self.canonical_Wvolume().detach().contiguous()
            self._key = key
        return self._vol


class PlaneNeRF(nn.Module):
    """ConditionalTriplaneNeRFModel_multiRender_split_view (model/nerf_model.py:10-117), default enc_mode='split'."""

    def __init__(self, XYZ_bounding, latent_code_dim, feat_dim=64, plane_res=128):
        super().__init__()
        self.XY_gen = styleunet.StyleGAN_zxc(out_ch=feat_dim, out_size=plane_res, style_dim=latent_code_dim, middle_size=16,
                                             zero_latent=False, zero_noise=True, no_skip=True, n_mlp=4, inp_size=256, inp_ch=7)
        self.YZ_gen = styleunet.StyleGAN_zxc(out_ch=feat_dim, out_size=plane_res, style_dim=latent_code_dim, middle_size=16,
                                             zero_latent=False, zero_noise=True, no_skip=True, n_mlp=4, inp_size=256, inp_ch=13)
        s, t = _box_warp(*XYZ_bounding)
        self.gridwarper = _BoxWarp(s, t)
        self.layers_xyz = nn.ModuleList([nn.Linear(2 * feat_dim + 48, 128), nn.Linear(128, 128)])
        self.fc_alpha = nn.Linear(128, 1)
        self.fc_rgbFeat = nn.Linear(128, 64)
        self.fc_rgb = nn.Linear(64, 3)
        self.triPlane_embeddings = None

    def set_conditional_embedding(self, front_render_cond, left_render_cond, right_render_cond, latents, cond_c):
        """model/nerf_model.py:58-86."""
        lat = [torch.cat([latents, cond_c.reshape(latents.shape[0], -1)], -1)]
        left = left_render_cond.flip(dims=[3])
        if left.shape[1] > 3:
            left = left[:, :-1]
        # the two generators on two streams (autograd replays each backward on its forward stream, so training overlaps them
        # too; parallel.GradSync orders its collectives after both streams)
        if front_render_cond.is_cuda:
            from .pipeline import two_stream_planes

            self.triPlane_embeddings = two_stream_planes(self, lat, front_render_cond, torch.cat([left, right_render_cond], dim=1))
            return
        xy, _ = self.XY_gen(lat, front_render_cond.contiguous())
        yz, _ = self.YZ_gen(lat, torch.cat([left, right_render_cond], dim=1).contiguous())
        self.triPlane_embeddings = torch.stack([xy, yz], dim=0)

    def mlp_weights(self):
        return {"layers_xyz.0.weight": self.layers_xyz[0].weight, "layers_xyz.0.bias": self.layers_xyz[0].bias,
                "layers_xyz.1.weight": self.layers_xyz[1].weight, "layers_xyz.1.bias": self.layers_xyz[1].bias,
                "fc_alpha.weight": self.fc_alpha.weight, "fc_alpha.bias": self.fc_alpha.bias,
                "fc_rgbFeat.weight": self.fc_rgbFeat.weight, "fc_rgbFeat.bias": self.fc_rgbFeat.bias,
                "fc_rgb.weight": self.fc_rgb.weight, "fc_rgb.bias": self.fc_rgb.bias}


class Trainer(nn.Module):
    def __init__(self, cfg, latent_codes_size=0, freeze_motion=True, precision="fp16"):
        """precision: 'fp16' (default), 'bf16', 'fp32', or 'auto' = fp16 with the range report on (render.render_rays
        check_range) and a permanent switch to bf16 the first time an operand leaves the fp16 range (inference renders only).
        freeze_motion: accepted and, like the reference, without effect -- nerf_trainer.py:35-36 sets `requires_grad` on the
        MODULE object, which freezes no parameter, so the skinning field trains (SURVEY.md section 8a quirk i)."""
        super().__init__()
        self.cfg = cfg
        ld = cfg.experiment.latent_code_dim
        self.latent_codes = nn.Parameter(torch.zeros(latent_codes_size, ld)) if latent_codes_size > 0 else None
        code_dim = ld + (12 if cfg.experiment.cond_pose else 0) + (52 if cfg.experiment.cond_expr else 0)
        if code_dim != ld + 12:
            raise NotImplementedError("only the shipped conditioning (cond_pose=True, cond_expr=False) is built")
        bounds = [list(b) for b in cfg.models.coarse.XYZ_bounding]
        self.model_coarse = PlaneNeRF(bounds, code_dim)
        self.render_size, self.gen_size = cfg.models.StyleUnet.inp_size, cfg.models.StyleUnet.out_size
        yb = [0.3 * float(bounds[1][1]), float(bounds[1][1])]                 # nerf_trainer.py:29-34
        s, t = _box_warp(bounds[0], yb, bounds[2])
        self.headpose_skin_net = SkinningField(_BoxWarp(s, t))
        self.precision = precision
        # True: draw sample_pdf's u on the device instead of the CPU generator (utils/nerf_util.py:93-96 draws on the CPU and
        # copies; a pageable host->device copy cannot be captured into a CUDA graph -- train_step.Graphed sets this)
        self.device_rng = False

    def _boxes(self):
        g, h = self.model_coarse.gridwarper, self.headpose_skin_net.gridwarper
        f = lambda b: [float(v) for v in b.detach().cpu().reshape(-1)]
        key = (g.scale_factor.data_ptr(), g.scale_factor._version, h.scale_factor._version)
        if getattr(self, "_box_key", None) != key:
            self._box_val, self._box_key = (f(g.scale_factor), f(g.trans_factor), f(h.scale_factor), f(h.trans_factor)), key
        return self._box_val

    def nerf_forward(self, **inputs):
        """model/nerf_trainer.py:38-92 without the chunk loop; returns the 7-tuple of predict_and_render_radiance."""
        opt = getattr(self.cfg.nerf, inputs["mode"])
        inv_head_T = inputs["inv_head_T"]
        # the skinning-weight VolumeDecoder (torch conv3d) is independent of the plane generators: evaluate it on its own stream
        # while they run (its backward then overlaps theirs as well)
        wvol, vol_stream = None, None
        if inv_head_T.is_cuda:
            from . import pipeline

            main = torch.cuda.current_stream(inv_head_T.device)
            vol_stream = pipeline.aux_stream(inv_head_T.device, 1)
            pipeline.note_fork(inv_head_T.device, main, vol_stream)
            vol_stream.wait_stream(main)
            with torch.cuda.stream(vol_stream):
                wvol = self.headpose_skin_net.volume()
        self.model_coarse.set_conditional_embedding(inputs["front_render_cond"], inputs["left_render_cond"],
                                                    inputs["right_render_cond"], inputs["latent_code"],
                                                    inv_head_T.reshape(inv_head_T.shape[0], -1))
        if vol_stream is not None:
            main.wait_stream(vol_stream)
            wvol.record_stream(main)
        else:
            wvol = self.headpose_skin_net.volume()
        ray_batch, bg = inputs.get("ray_batch"), inputs["background_prior"]
        cam = {}
        if ray_batch is None:       # rays generated inside the kernel from the 18-float camera block (havatar_b200/data.py)
            cam = dict(camera=inputs["camera"], img_hw=inputs["img_hw"], pixel_index=inputs.get("pixel_index"))
            B = cam["camera"].shape[0]
            R = cam["pixel_index"].shape[1] if cam["pixel_index"] is not None else int(cam["img_hw"][0]) * int(cam["img_hw"][1])
            dev = cam["camera"].device
        else:
            B, R = ray_batch.shape[:2]
            dev = ray_batch.device
        nc, nf = int(opt.num_coarse), int(opt.num_fine)
        rnd = {}
        given = inputs.get("randoms")       # tests: the reference's draws as explicit tensors (SURVEY.md section 8a quirk v)
        if given is not None:
            rnd = {k: v for k, v in given.items() if k in ("t_rand", "u_rand", "noise_coarse", "noise_fine") and v is not None}
        elif opt.perturb:                                                           # :132-139, utils/nerf_util.py:93-96
            rnd["t_rand"] = torch.rand(B, R, nc, dtype=torch.float32, device=dev)
            if nf > 0:
                if self.device_rng:
                    rnd["u_rand"] = torch.rand(B, R, nf, dtype=torch.float32, device=dev)
                else:
                    rnd["u_rand"] = torch.rand([B * R, nf], dtype=torch.float32).to(dev).view(B, R, nf)   # CPU generator, like the reference
        std = float(opt.radiance_field_noise_std)
        if given is None and std > 0.0:                                                             # utils/nerf_util.py:47-57
            rnd["noise_coarse"] = torch.randn(B, R, nc, device=dev) * std
            if nf > 0:
                rnd["noise_fine"] = torch.randn(B, R, (nc + 1) // 2 + nf, device=dev) * std
        args = (ray_batch, bg, inv_head_T, self.model_coarse.triPlane_embeddings, wvol, self.model_coarse.mlp_weights(), nc, nf)
        prec = "fp16" if self.precision == "auto" else self.precision
        if torch.is_grad_enabled():
            o = hrender.render_rays_autograd(*args, boxes=self._boxes(), precision=prec, **cam, **rnd)
        elif self.precision == "auto":
            try:
                o = hrender.render_rays(*args, boxes=self._boxes(), precision="fp16", check_range=True, **cam, **rnd)
            except hrender._lib.HavError as e:
                if "fp16 operand range" not in str(e):
                    raise
                import warnings
                warnings.warn("havatar_b200: %s -- switching this Trainer to bf16 operands" % e)
                self.precision = "bf16"
                o = hrender.render_rays(*args, boxes=self._boxes(), precision="bf16", **cam, **rnd)
        else:
            o = hrender.render_rays(*args, boxes=self._boxes(), precision=prec, **cam, **rnd)
        return o.rgb_coarse, o.depth_coarse, o.acc_coarse, o.weights_max, o.rgb_fine, o.depth_fine, o.acc_fine

    def forward(self, **data):
        """model/nerf_trainer.py:94-118 (including its quirk ii: slots 1 and 5 of the long tuple are both depth_fine)."""
        ray_batch = data.get("ray_batch")
        B = (ray_batch if ray_batch is not None else data["camera"]).shape[0]
        latent_code = self.latent_codes[data["fidx"]] if data["mode"] == "train" else self.latent_codes[0:1]
        latent_code_loss = torch.square(latent_code - self.latent_codes.mean(dim=0, keepdims=True).detach()).mean()
        rgb_c, _, acc_c, weights, rgb_f, depth_f, acc_f = self.nerf_forward(
            ray_batch=ray_batch, background_prior=data["background_prior"], latent_code=latent_code, inv_head_T=data["inv_head_T"],
            front_render_cond=data["front_render_cond"], left_render_cond=data["left_render_cond"],
            right_render_cond=data["right_render_cond"], mode=data["mode"], randoms=data.get("randoms"), camera=data.get("camera"),
            img_hw=data.get("img_hw"), pixel_index=data.get("pixel_index"))
        if data["render_full_img"]:
            render = rgb_f if rgb_f is not None else rgb_c
            mask = acc_f if acc_f is not None else acc_c
            s = self.render_size
            return (render.reshape(B, s, s, -1).permute(0, 3, 1, 2), mask.reshape(B, s, s, -1).permute(0, 3, 1, 2), latent_code_loss)
        return rgb_c, depth_f, acc_c, weights, rgb_f, depth_f, acc_f, latent_code_loss
