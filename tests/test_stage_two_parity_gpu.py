"""Stage-two training step (BASELINE.json configs[4]; train_avatarHD.py:201-303) against the UNMODIFIED reference modules:
tests/golden/stage_two_step.npz holds the losses of the D / R1 / G phases, gradients taken right after each backward and a
few weights after the optimiser steps, from oracle/gen_golden.py::gen_stage_two (reference Trainer + SWGAN_unet +
Discriminator + utils/styleUnet_util.py losses + torch.optim.Adam on CPU, iteration 0, LPIPS omitted, draws supplied).
Here havatar_b200.train_step.StageTwoStep runs the same iteration from the same weights on the B200 kernels."""
import os

import numpy as np
import pytest
import torch

from conftest import within
from havatar_b200 import train_step
from oracle.gen_golden import STAGE_TWO_CASE, STAGE_TWO_GRAD_KEYS, stage_two_inputs, stage_two_states, subsample

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_stage_two_iteration_matches_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "stage_two_step.npz"))
    c = STAGE_TWO_CASE
    cfg = train_step.default_cfg(num_coarse=c["num_coarse"], num_fine=c["num_fine"], inp_size=c["render_size"], out_size=c["gen_size"],
                                 lr=1e-4)
    step = train_step.StageTwoStep(n_frames=4, cfg=cfg, render_size=c["render_size"], gen_size=c["gen_size"], latent=c["latent"],
                                   n_mlp=c["n_mlp"], seed=0)
    shp = lambda m: {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd_n, sd_g, sd_d = stage_two_states(shp(step.net), shp(step.generator), shp(step.disc), c["seed"])
    with torch.no_grad():     # in place: the parameters are views into the optimisers' flat buffers
        for m, sd in ((step.net, sd_n), (step.generator, sd_g), (step.g_ema, sd_g), (step.disc, sd_d)):
            tgt = m.state_dict()
            for k, v in sd.items():
                tgt[k].copy_(torch.from_numpy(v))
    train_step.styleunet.invalidate_caches()
    inp = stage_two_inputs(c)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    step.net.model_coarse.XY_gen.zero_noise[0] = dev(inp["noise0"]["XY_gen"])
    step.net.model_coarse.YZ_gen.zero_noise[0] = dev(inp["noise0"]["YZ_gen"])
    sc = inp["scene"]
    batch = dict(ray_batch=dev(sc["ray_batch"]), background_prior=dev(sc["background_prior"]), inv_head_T=dev(sc["inv_head_T"]),
                 fidx=torch.tensor([1, 3]).cuda(), gt_hr_img=dev(inp["gt_hr"]), gt_lr_mask=dev(inp["gt_mask"]),
                 **{k: dev(v) for k, v in inp["conds"].items()})
    for ph in ("d", "g"):
        batch["randoms_" + ph] = {k: dev(inp["rnd"][ph][k]) for k in ("t_rand", "u_rand", "noise_coarse", "noise_fine")}
        batch["z_" + ph] = dev(inp["z"][ph])
        batch["gen_noise_" + ph] = [dev(n) for n in inp["gnoise"][ph]]

    # record gradients right before every optimiser step (FlatAdam clears them in the same pass)
    rec, calls = {}, []
    names = {"d": dict(step.disc.named_parameters()), "g": dict(step.generator.named_parameters()),
             "nerf": dict(step.net.named_parameters())}
    orig = train_step._Group.step

    def spy(self):
        tag = "nerf" if self is step.nerf else "g" if self is step.g else ("d" if "d" not in calls else "r1")
        calls.append(tag)
        src = names["d" if tag in ("d", "r1") else tag]
        for k in STAGE_TWO_GRAD_KEYS[tag]:
            rec["g_%s_%s" % (tag, k)] = subsample(src[k].grad.detach().float().cpu().numpy())
        return orig(self)

    train_step._Group.step = spy
    try:
        out = step(batch)
    finally:
        train_step._Group.step = orig
    torch.cuda.synchronize()
    assert calls == ["d", "r1", "g", "nerf"]                 # iteration 0 regularises (i % d_reg_every == 0, train_avatarHD.py:209)

    # ---- losses.  16-bit tensor-core operands through three deep networks: 2e-3 relative on O(1) losses (measured 9e-6 .. 6e-4,
    #      profiles/r02zz_tolerance_margins.txt; the limit was 2e-2 until then); the R1 penalty is a second-order quantity of ~4e-3
    #      absolute: 2.5e-2 relative (measured 7.6e-3)
    f = lambda k: float(out[k])
    within("stage two d_loss (relative)", abs(f("d") - float(g["d_loss"])) / abs(float(g["d_loss"])), 2e-3)
    within("stage two r1 (relative)", abs(f("r1") - float(g["r1"])) / abs(float(g["r1"])), 2.5e-2)
    for k in ("g_loss", "rgb_loss", "mask_loss", "g_nonsat", "hr_l1"):
        within("stage two %s (relative)" % k, abs(f(k) - float(g[k])) / abs(float(g[k])), 2e-3)
    # ---- gradients.  Metric = relative L2 error and cosine against the reference's gradient (a max-norm over a tensor whose
    #      entries cancel heavily measures the noise floor of the 16-bit operands, not the agreement of the two gradients).
    #      Inputs are not identical by then: the D step sees OUR generator's fake image (1e-2-class differences), the R1 pass
    #      runs after one Adam step (sign(g) * lr per weight), and R1 is a second-order quantity through two 16-bit passes.
    table, bad = [], {}
    for k, v in sorted(rec.items()):
        ref = g[k].astype(np.float64)
        l2 = float(np.linalg.norm(v - ref) / (np.linalg.norm(ref) + 1e-30))
        cos = float((v * ref).sum() / (np.linalg.norm(v) * np.linalg.norm(ref) + 1e-30))
        table.append("%-64s relL2 %.4f  cos %.5f  max-rel %.4f" % (k, l2, cos, _rel(v, g[k])))
        lim_l2, lim_cos = (0.25, 0.99) if k.startswith("g_r1_") else (0.12, 0.995)      # measured: 0.12 / 0.997 and 0.08 / 0.9994
        if l2 > lim_l2 or cos < lim_cos:
            bad[k] = (l2, cos)
    print("\n".join(table))
    assert not bad, (bad, table)
    # ---- weights after the whole iteration (Adam's first step moves every weight by ~lr in the gradient's direction)
    for key, mod, lr in (("w_disc_final_linear.1.weight", names["d"]["final_linear.1.weight"], 2 * 1e-3 * 16 / 17),
                         ("w_gen_to_rgbs.1.conv.weight", names["g"]["to_rgbs.1.conv.weight"], 1e-3 * 0.8),
                         ("w_nerf_model_coarse.fc_rgb.weight", names["nerf"]["model_coarse.fc_rgb.weight"], 1e-4)):
        got = subsample(mod.detach().float().cpu().numpy())
        # sign flips of near-zero gradient entries move a weight by up to 2 * lr: bounded, and rare
        assert np.abs(got - g[key]).max() <= 2.05 * lr + 1e-7, key
        assert (np.abs(got - g[key]) > 0.1 * lr).mean() < 0.15, key
    ema = dict(step.g_ema.named_parameters())["to_rgbs.1.conv.weight"]
    assert np.abs(subsample(ema.detach().cpu().numpy()) - g["w_ema_to_rgbs.1.conv.weight"]).max() < 1e-5
