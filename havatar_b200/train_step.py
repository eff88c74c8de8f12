"""The two training steps of the reference as callable objects on the B200 kernels, with the data-parallel gradient exchange of
SURVEY.md section 8e (BASELINE.json configs[2] and configs[4]).

  StageOneStep  = one iteration of train_avatar.py:112-158 (Trainer.forward(mode='train') on B patches of 64x64 rays, 64 + 16
                  hierarchical samples, perturb + density noise; mse + mask BCE on both passes + latent regulariser + the
                  skinning-volume smoothness term :124-129; Adam).  LPIPS is not available offline, so the
                  patch term is the non-saturating GAN loss through a 64x64 `Discriminator` on the rendered patch (the
                  "GAN discriminator" BASELINE.json names for this config) and the discriminator takes its own logistic step.
  StageTwoStep  = one iteration of train_avatarHD.py:201-303: D step (render + generator without grad, logistic loss, Adam),
                  R1 every d_reg_every (double backward through our convolution / upfirdn2d / fused_leaky_relu autograd), G step (render with
                  grad -> low-res losses; generator -> non-saturating + L1; one backward; generator and render Adam steps),
                  EMA accumulate (utils/styleUnet_util.py:51-56).  LPIPS omitted for the same reason.

What runs underneath: the fused tcgen05 render forward + backward (hav_render_forward / hav_render_backward), our upfirdn2d /
fused_bias_act kernels with their 1st / 2nd order autograd, and -- for the convolutions' gradients -- see styleunet_train.py.
With torch.distributed initialised each rank steps on its own frames and `parallel.GradSync` all-reduces the gradients
(bucketed, overlapped with backward); weights are broadcast from rank 0 at construction; optimiser steps and the EMA are
replicated."""
from types import SimpleNamespace as NS

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import parallel, pipeline, styleunet, trainer


def default_cfg(num_coarse=64, num_fine=16, perturb=True, noise_std=0.1, inp_size=128, out_size=512, lr=5e-4):
    """The fields of config/singleview_512_base.yml the orchestrator and the steps read (lr: optimizer.lr, 5e-4 in
    singleview_512_base.yml:87, 1e-4 in singleview_512_HD_base.yml:116; scheduler: :90-94)."""
    mode = lambda p, s: NS(num_coarse=num_coarse, num_fine=num_fine, perturb=p, radiance_field_noise_std=s, chunksize=4096)
    return NS(experiment=NS(latent_code_dim=32, cond_pose=True, cond_expr=False, model_mode=None, mask_weight=0.01),
              models=NS(coarse=NS(XYZ_bounding=[[-1.5, 1.5], [-1.6, 1.4], [-1.6, 1.2]],
                                  Head_bounding=[[-1.2, 1.2], [-1.6, 1.0], [-1.6, 1.2]]),
                        StyleUnet=NS(inp_size=inp_size, out_size=out_size)),
              optimizer=NS(type="Adam", lr=lr), scheduler=NS(lr_decay=250, lr_decay_factor=0.1),
              nerf=NS(train=mode(perturb, noise_std), validation=mode(False, 0.0)))


def _cfg_get(cfg, path, default):
    cur = cfg
    for name in path.split("."):
        cur = getattr(cur, name, None)
        if cur is None:
            return default
    return cur


def d_logistic_loss(real_pred, fake_pred):                      # utils/styleUnet_util.py:65-69
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def g_nonsaturating_loss(fake_pred):                            # utils/styleUnet_util.py:82-85
    return F.softplus(-fake_pred).mean()


def d_r1_loss(real_pred, real_img):                             # utils/styleUnet_util.py:72-79
    from .op import conv2d_gradfix

    with conv2d_gradfix.no_weight_gradients():
        grad_real, = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def volume_smoothness(w):                                       # train_avatar.py:124-129
    core = w[1:-1, 1:-1, 1:-1]
    nb = [w[:-2, 1:-1, 1:-1], w[2:, 1:-1, 1:-1], w[1:-1, 2:, 1:-1], w[1:-1, :-2, 1:-1], w[1:-1, 1:-1, 2:], w[1:-1, 1:-1, :-2]]
    return torch.mean(sum(torch.abs(core - v) for v in nb) / 6.0)


def _distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


FLAT_ADAM = [True]      # bench.py's reference-formulation leg switches the steps back to torch.optim.Adam


class _Group:
    """Parameters + Adam + (when distributed) the bucketed gradient all-reduce of one optimiser.  On a CUDA device the
    parameters, gradients and both moments are flat buffers and the update is the one-pass hav_adam_flat kernel
    (parallel.FlatAdam; step count and learning rate on the device, so the whole update can live inside a CUDA graph)."""

    def __init__(self, modules, lr, betas=(0.9, 0.999)):
        self.params = [p for m in modules for p in m.parameters()]
        self.flat = self.params[0].is_cuda and FLAT_ADAM[0]
        if self.flat:
            self.opt = parallel.FlatAdam(self.params, lr, betas)
            self.sync = parallel.GradSync(self.params, layout=self.opt.layout) if _distributed() else None
        else:       # construction on a CPU device (checkpoint / shape inspection): the steps themselves need the CUDA kernels
            self.sync = parallel.GradSync(self.params) if _distributed() else None
            self.opt = torch.optim.Adam(self.params, lr=lr, betas=betas)

    def set_lr(self, lr):
        if self.flat:
            self.opt.set_lr(lr)
        else:
            for grp in self.opt.param_groups:
                grp["lr"] = float(lr)

    def step(self):
        """All-reduce (if distributed), Adam update, gradients cleared."""
        if self.sync is not None:
            self.sync.finish()
        self.opt.step()
        if not self.flat:
            self.zero_grad()

    def zero_grad(self):
        if self.flat:
            self.opt.zero_grad()
        elif self.sync is not None:
            self.sync.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)

    def requires_grad(self, flag):
        for p in self.params:
            p.requires_grad_(flag)


class StageOneStep:
    def __init__(self, n_frames=4, device="cuda", cfg=None, precision="fp16", patch=64, lr=None, with_discriminator=True, seed=0,
                 capturable=False, pretrain_wc_iters=0, overlap=True):
        """pretrain_wc_iters > 0: a run that does not resume from a checkpoint first fits the skinning-weight volume to the head
        box (train_avatar.py:93-95 uses 3000 iterations); load a checkpoint into `self.net` instead when resuming."""
        torch.manual_seed(seed)
        self.cfg = cfg or default_cfg()
        lr = float(_cfg_get(self.cfg, "optimizer.lr", 5e-4)) if lr is None else lr
        self.device = torch.device(device)
        self.net = trainer.Trainer(self.cfg, n_frames, precision=precision).to(self.device)
        self.net.device_rng = capturable
        if pretrain_wc_iters > 0:
            self.net.headpose_skin_net.pretrain_wc(num_iter=pretrain_wc_iters,
                                                   vol_thr=_cfg_get(self.cfg, "models.coarse.Head_bounding", None))
        self.disc = styleunet.Discriminator(patch, img_channel=3).to(self.device) if with_discriminator else None
        mods = [self.net] + ([self.disc] if self.disc is not None else [])
        if _distributed():
            parallel.broadcast_parameters(mods)
        self.g = _Group([self.net], lr)                                                                   # train_avatar.py:68-71
        self.d = _Group([self.disc], 2e-3 * 16 / 17, betas=(0.0, 0.99 ** (16 / 17))) if self.disc is not None else None
        self.patch, self.it, self.lr0 = patch, 0, lr
        # overlap: run the patch discriminator's own forward / backward on a side stream WHILE the render network's backward runs
        # (neither reads what the other writes; both optimiser steps wait for both)
        self.overlap = bool(overlap) and self.device.type == "cuda"

    def groups(self):
        return [g for g in (self.g, self.d) if g is not None]

    def pre_step(self):
        """Host-side bookkeeping of one iteration (outside any CUDA graph): the exponential learning-rate decay of
        train_avatar.py:154-158.  The reference sets lr(i) AFTER optimizer.step() of iteration i (0-based), so iteration i runs
        with lr(i - 1) and iteration 0 with the undecayed rate."""
        if self.it > 0:
            steps = float(_cfg_get(self.cfg, "scheduler.lr_decay", 250)) * 1000.0
            factor = float(_cfg_get(self.cfg, "scheduler.lr_decay_factor", 0.1))
            self.g.set_lr(max(self.lr0 * (factor ** ((self.it - 1) / steps)), 5e-5))
        self.it += 1

    def parts(self):
        return [("step", self.body, True)]

    def __call__(self, batch):
        """batch: dict with ray_batch [B,R,8], background_prior [B,R,3], target [B,R,3], mask [B,R,1], fidx [B], inv_head_T
        [B,4,3], {front,left,right}_render_cond [B,7,256,256] (device tensors).  Returns {'loss', 'd_loss'} (detached)."""
        self.pre_step()
        return self.body(batch)

    def body(self, batch):
        cfg, P = self.cfg, self.patch
        if self.d is not None:
            self.d.requires_grad(False)
        out = self.net(mode="train", fidx=batch["fidx"], render_full_img=False, ray_batch=batch["ray_batch"],
                       background_prior=batch["background_prior"], inv_head_T=batch["inv_head_T"],
                       front_render_cond=batch["front_render_cond"], left_render_cond=batch["left_render_cond"],
                       right_render_cond=batch["right_render_cond"], randoms=batch.get("randoms"))
        rgb_c, _, acc_c, _, rgb_f, _, acc_f, lat = out
        target, mask, mw = batch["target"], batch["mask"], cfg.experiment.mask_weight
        loss = F.mse_loss(rgb_c[..., :3], target) + mw * F.binary_cross_entropy(acc_c.clip(1e-3, 1.0 - 1e-3), mask)   # :131-132
        if rgb_f is not None:                                                                                       # :134-136
            loss = loss + F.mse_loss(rgb_f[..., :3], target) + mw * F.binary_cross_entropy(acc_f.clip(1e-3, 1.0 - 1e-3), mask)
        # :124-129 re-evaluates the VolumeDecoder for this term; the forward's own evaluation is the same tensor, reuse it
        skin = self.net.headpose_skin_net
        vol = skin.last_volume if skin.last_volume is not None and not skin.fix_canoW else skin.canonical_Wvolume()
        sw = volume_smoothness(vol[0, 1])
        skin.last_volume = None
        loss = loss + lat + 1e-4 * sw                                                                               # :146
        fake = None
        if self.disc is not None:
            rgb = rgb_f if rgb_f is not None else rgb_c
            B = rgb.shape[0]
            fake = rgb[..., :3].reshape(B, P, P, 3).permute(0, 3, 1, 2).contiguous()
            loss = loss + 0.05 * g_nonsaturating_loss(self.disc(fake))       # in the slot of the 0.05-weighted patch term (:144)
        d_loss = None

        def d_pass():
            self.d.requires_grad(True)
            real = batch["target"].reshape(-1, P, P, 3).permute(0, 3, 1, 2).contiguous()
            fake_d = fake.detach()
            real_pred, fake_pred = pipeline.run_parallel(lambda: self.disc(real), lambda: self.disc(fake_d), self.device)
            dl = d_logistic_loss(real_pred, fake_pred)
            dl.backward()
            return dl

        if self.disc is not None and self.overlap:
            main = torch.cuda.current_stream(self.device)
            side = pipeline.aux_stream(self.device, 4)
            pipeline.note_fork(self.device, main, side)
            side.wait_stream(main)
            fake.record_stream(side)
            with torch.cuda.stream(side), pipeline.stream_namespace(1):
                d_loss = d_pass()
            loss.backward()                                                                                         # :149
            main.wait_stream(side)
            d_loss.record_stream(main)
            self.g.step()                                                                                           # :151 (clears the gradients)
            self.d.step()
        else:
            loss.backward()
            self.g.step()
            if self.disc is not None:
                d_loss = d_pass()
                self.d.step()
        return {"loss": loss.detach(), "d_loss": None if d_loss is None else d_loss.detach()}


class StageTwoStep:
    def __init__(self, n_frames=8, device="cuda", cfg=None, precision="fp16", render_size=128, gen_size=512, d_reg_every=16,
                 r1=10.0, latent=64, n_mlp=4, seed=0, capturable=False, lr=1e-3, nerf_lr=None, overlap=True):
        """lr: su_args.lr as train_avatarHD.py:118 overrides it (1e-3); nerf_lr: cfg.optimizer.lr (1e-4 in
        config/singleview_512_HD_base.yml:116).  overlap: on iterations without the R1 pass, run the G step's forward (render +
        generator, which do not depend on the discriminator) on a side stream WHILE the D step runs (dg_step)."""
        torch.manual_seed(seed)
        self.cfg = cfg or default_cfg(inp_size=render_size, out_size=gen_size, lr=1e-4)
        nerf_lr = float(_cfg_get(self.cfg, "optimizer.lr", 1e-4)) if nerf_lr is None else nerf_lr
        self.device = torch.device(device)
        self.net = trainer.Trainer(self.cfg, n_frames, precision=precision).to(self.device)                 # train_avatarHD.py:109
        self.net.device_rng = capturable
        mk = lambda: styleunet.SWGAN_unet(inp_size=render_size, inp_ch=64, out_ch=3, out_size=gen_size, style_dim=latent,
                                          n_mlp=n_mlp).to(self.device)                                          # :110-111
        self.generator, self.g_ema = mk(), mk()
        self.g_ema.load_state_dict(self.generator.state_dict())
        self.g_ema.requires_grad_(False)
        self.disc = styleunet.Discriminator(gen_size, img_channel=3).to(self.device)                         # :112
        if _distributed():
            parallel.broadcast_parameters([self.net, self.generator, self.g_ema, self.disc])
        g_ratio, d_ratio = 4 / 5, d_reg_every / (d_reg_every + 1)                                            # :117-122
        self.nerf = _Group([self.net], nerf_lr)                                                              # :121
        self.g = _Group([self.generator], lr * g_ratio, betas=(0.0, 0.99 ** g_ratio))                        # :119
        self.d = _Group([self.disc], lr * d_ratio, betas=(0.0, 0.99 ** d_ratio))                             # :120
        self.render_size, self.gen_size, self.latent = render_size, gen_size, latent
        self.d_reg_every, self.r1, self.it = d_reg_every, r1, 0
        self.overlap = bool(overlap) and self.device.type == "cuda"
        self.accum = 0.5 ** (32 / (10 * 1000))                                                               # :162
        self.gan_w = torch.tensor(1e-3, dtype=torch.float32, device=self.device)

    def groups(self):
        return [self.nerf, self.g, self.d]

    def _noise(self, b, given=None):                                 # mixing_noise with mixing = 0 (:170-175)
        return [given] if given is not None else [torch.randn(b, self.latent, device=self.device)]

    def pre_step(self):
        """self.it counts iterations STARTED; the reference's 0-based index i = self.it - 1 gates R1 (i % d_reg_every == 0, so
        iteration 0 regularises, :209) and sets the GAN weight (:205-206)."""
        self.it += 1
        self.gan_w.fill_(min(1e-3 * 1.1 ** ((self.it - 1) // 500), 0.1))                                       # :205-206

    def parts(self):
        """The parts of THIS iteration (after pre_step): the overlapped D + G part unless the R1 pass sits between them."""
        if self.overlap and (self.it - 1) % self.d_reg_every != 0:
            return [("dg", self.dg_step, True)]
        return [("d", self.d_step, True), ("r1", self.r1_step, False), ("g", self.g_step, True)]

    def all_parts(self):
        """Every part any iteration can consist of (what train_step.Graphed captures)."""
        seq = [("d", self.d_step, True), ("r1", self.r1_step, False), ("g", self.g_step, True)]
        return seq + ([("dg", self.dg_step, True)] if self.overlap else [])

    def __call__(self, batch):
        """batch: the StageOneStep keys with full low-res frames (R = render_size^2) plus gt_hr_img [B,3,G,G] and gt_lr_mask
        [B,1,render,render]."""
        self.pre_step()
        out = {}
        for _, fn, _ in self.parts():
            out.update(fn(batch) or {})
        return out

    def _inp(self, batch, phase):
        """Optional explicit draws for parity tests (SURVEY.md section 8a quirk v): batch['randoms_d' / 'randoms_g'] = the
        render's random tensors, batch['z_d' / 'z_g'] = the mixing noise, batch['gen_noise_d' / 'gen_noise_g'] = the per-layer
        generator noise maps of the D-step / G-step forward."""
        return dict(mode="train", fidx=batch["fidx"], render_full_img=True, ray_batch=batch["ray_batch"],
                    background_prior=batch["background_prior"], inv_head_T=batch["inv_head_T"],
                    front_render_cond=batch["front_render_cond"], left_render_cond=batch["left_render_cond"],
                    right_render_cond=batch["right_render_cond"], randoms=batch.get("randoms_" + phase))

    def d_step(self, batch, toggle=True):                                                                       # :211-231
        gt_hr = batch["gt_hr_img"]
        B = gt_hr.shape[0]
        if toggle:
            self.nerf.requires_grad(False), self.g.requires_grad(False)
        self.d.requires_grad(True)
        with torch.no_grad():
            render, _, _ = self.net(**self._inp(batch, "d"))
            fake = self.generator(self._noise(B, batch.get("z_d")), render[:, 3:].contiguous(), noise=batch.get("gen_noise_d"))
        real_pred, fake_pred = pipeline.run_parallel(lambda: self.disc(gt_hr), lambda: self.disc(fake), self.device)
        d_loss = d_logistic_loss(real_pred, fake_pred) * self.gan_w
        d_loss.backward()
        self.d.step()
        return {"d_loss": d_loss.detach(), "d": (d_loss / self.gan_w).detach()}

    def r1_step(self, batch):                                                                                   # :233-240
        if (self.it - 1) % self.d_reg_every != 0:                                                            # :209 d_regularize
            return {"r1": None}
        self.nerf.requires_grad(False), self.g.requires_grad(False), self.d.requires_grad(True)
        real = batch["gt_hr_img"].detach().requires_grad_(True)
        pred = self.disc(real)
        r1_loss = d_r1_loss(pred, real) * self.gan_w
        (self.r1 / 2 * r1_loss * self.d_reg_every + 0 * pred[0]).sum().backward()
        self.d.step()
        return {"r1": (r1_loss / self.gan_w).detach()}                                                       # loss_dict["r1"], :240

    def _g_forward(self, batch):
        """The part of the G step that does not involve the discriminator (:243-262): low-res render with gradients, its
        losses, the generator forward."""
        gt_hr, rs, gs = batch["gt_hr_img"], self.render_size, self.gen_size
        B = gt_hr.shape[0]
        gt_lr = F.interpolate(F.interpolate(gt_hr, size=(rs, rs), mode="bilinear", align_corners=True), size=(gs, gs),
                              mode="bilinear", align_corners=True)                                              # :202-204
        render, mask, lat = self.net(**self._inp(batch, "g"))
        lr_img = F.interpolate(render[:, :3], size=(gs, gs), mode="bilinear", align_corners=True)
        rgb_loss = F.mse_loss(lr_img, gt_lr)
        mask_loss = self.cfg.experiment.mask_weight * F.binary_cross_entropy(mask.clip(1e-3, 1.0 - 1e-3), batch["gt_lr_mask"])
        fake = self.generator(self._noise(B, batch.get("z_g")), render[:, 3:].contiguous(), noise=batch.get("gen_noise_g"))
        hr_l1 = F.l1_loss(fake, gt_hr)
        return dict(fake=fake, partial=rgb_loss + lat + mask_loss + hr_l1, rgb_loss=rgb_loss, mask_loss=mask_loss, hr_l1=hr_l1)

    def _g_finish(self, gf):
        """:264-303 from the discriminator's verdict on: adversarial term, one backward, generator and render Adam steps, EMA."""
        self.d.requires_grad(False)
        g_ns = g_nonsaturating_loss(self.disc(gf["fake"]))
        g_loss = gf["partial"] + g_ns * self.gan_w
        g_loss.backward()
        self.g.step()
        self.nerf.step()
        with torch.no_grad():                                                                                   # :303
            pe, pg = list(self.g_ema.parameters()), list(self.generator.parameters())
            torch._foreach_mul_(pe, self.accum)
            torch._foreach_add_(pe, pg, alpha=1 - self.accum)
        return {"g_loss": g_loss.detach(), "rgb_loss": gf["rgb_loss"].detach(), "mask_loss": gf["mask_loss"].detach(),
                "g_nonsat": g_ns.detach(), "hr_l1": gf["hr_l1"].detach()}

    def g_step(self, batch):                                                                                    # :243-303
        self.nerf.requires_grad(True), self.g.requires_grad(True)
        return self._g_finish(self._g_forward(batch))

    def dg_step(self, batch):
        """D step and G step of an iteration without the R1 pass, with the G step's forward overlapped: the low-res render and
        the generator forward of the G step read only weights the D step does not touch, so they run on a side stream (with
        their own family of sub-streams) while the D step -- no-grad render + generator, discriminator forward / backward, Adam
        -- runs on the main one; the streams join before the updated discriminator judges the G step's image.  Same arithmetic
        and update order as d_step(); g_step(), only the random draws are taken in a different order."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        side = pipeline.aux_stream(dev, 4)
        pipeline.note_fork(dev, main, side)
        self.nerf.requires_grad(True), self.g.requires_grad(True)
        side.wait_stream(main)
        with torch.cuda.stream(side), pipeline.stream_namespace(1):
            gf = self._g_forward(batch)
        out = self.d_step(batch, toggle=False)
        main.wait_stream(side)
        for v in gf.values():
            v.record_stream(main)
        out.update(self._g_finish(gf))
        out["r1"] = None
        return out


class Graphed:
    """A training step replayed as CUDA graphs: forward, backward, the gradient all-reduce, the Adam updates and the EMA of one
    iteration become one graph launch per graphable part (StageOneStep: one; StageTwoStep: the D step and the G step, with the
    every-16th R1 pass run eagerly between them), removing the ~1500 host-side kernel launches that bound the eager step.
    The step must be built with capturable=True (sample_pdf's uniform draws on the device instead of the CPU generator; the
    Adam state is device-side either way).
    `example` provides shapes; its tensors are cloned into the static input buffers every replay reads."""

    def __init__(self, step, example, warmup=3):
        self.step = step
        self.static = {k: v.clone() for k, v in example.items() if isinstance(v, torch.Tensor)}
        dev = step.device
        self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        self.graphs, self.outs = {}, {}
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):                     # allocator / cuDNN / Adam-state warm-up on the capture stream
                step(self.static)
            for grp in step.groups():                   # non-flat groups: gradients must be (re)allocated from the graphs' pool
                if not grp.flat and grp.sync is None:
                    grp.opt.zero_grad(set_to_none=True)
            step.pre_step()
            for name, fn, graphable in (step.all_parts() if hasattr(step, "all_parts") else step.parts()):
                # derived-tensor caches (packed weights, ...) must not cross a capture boundary: a hit on an eagerly built
                # entry would leave the graph without the kernel that refreshes it, and eager code must not reuse graph memory
                styleunet.invalidate_caches()
                if graphable:
                    from . import conv

                    conv.CAPTURE_GEN[0] += 1
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        self.outs[name] = fn(self.static)
                    self.graphs[name] = g
                    g.replay()                          # the captured iteration has not executed yet: run it once
                else:
                    fn(self.static)
            styleunet.invalidate_caches()
        torch.cuda.current_stream(dev).wait_stream(self.stream)

    def __call__(self, batch):
        for k, v in self.static.items():
            v.copy_(batch[k], non_blocking=True)
        self.step.pre_step()
        out = {}
        for name, fn, graphable in self.step.parts():
            if graphable:
                self.graphs[name].replay()
                out.update(self.outs[name] or {})
            else:
                styleunet.invalidate_caches()       # replays changed parameters behind their version counters
                out.update(fn(self.static) or {})
        styleunet.invalidate_caches()
        return out


def synthetic_batch(stage, batch, device, seed=0, patch=64, render_size=128, gen_size=512, frame_offset=0):
    """Seeded synthetic inputs in the dataloader's layouts (dataloader/dataloader.py:146-229): stage 1 =
    `batch` frames x one patch x patch window of rays of a 512x512 camera; stage 2 = `batch` full render_size^2 low-res frames +
    a gen_size^2 ground-truth image.  `frame_offset` = first global frame index of this rank (data-parallel sharding)."""
    import numpy as np

    from . import synth

    dev = torch.device(device)
    g = torch.Generator().manual_seed(1000 + seed)
    if stage == 1:
        sc = synth.scene(batch=batch, height=512, width=512, crop=(256 - patch // 2, 256 - patch // 2, patch, patch), seed=seed)
    else:
        sc = synth.scene(batch=batch, height=render_size, width=render_size, seed=seed)
    R = sc["ray_batch"].shape[1]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cond = lambda: torch.rand(batch, 7, 256, 256, generator=g).to(dev)
    out = {"ray_batch": t(sc["ray_batch"]), "background_prior": t(sc["background_prior"]), "inv_head_T": t(sc["inv_head_T"]),
           "front_render_cond": cond(), "left_render_cond": cond(), "right_render_cond": cond(),
           "fidx": (torch.arange(batch) + frame_offset).to(dev),
           "target": torch.rand(batch, R, 3, generator=g).to(dev),
           "mask": (torch.rand(batch, R, 1, generator=g) > 0.4).float().to(dev)}
    if stage == 2:
        out["gt_hr_img"] = (torch.rand(batch, 3, gen_size, gen_size, generator=g) * 2 - 1).to(dev)
        out["gt_lr_mask"] = out["mask"].reshape(batch, render_size, render_size, 1).permute(0, 3, 1, 2).contiguous()
    return out
