"""The two training steps (havatar_b200/train_step.py; BASELINE.json configs[2] and configs[4]) run end to end on the GPU at
reduced sizes: losses finite, every sub-network's weights move, and the R1 double backward goes through our op autograd.
Gradient PARITY of the stage-one step against the unmodified reference is tests/test_trainer.py."""
import pytest
import torch

from havatar_b200 import train_step

pytestmark = pytest.mark.gpu


def _snap(mods):
    return [p.detach().clone() for m in mods for p in m.parameters()]


def _moved(before, mods):
    after = [p for m in mods for p in m.parameters()]
    return sum(int(not torch.equal(a, b)) for a, b in zip(before, after)), len(after)


def _unmoved_names(before, mods):
    named = [(n, p) for m in mods for n, p in m.named_parameters()]
    return {n for (n, p), b in zip(named, before) if torch.equal(p, b)}


# What the REFERENCE itself leaves without a gradient in a plane generator (listed by running the unmodified Trainer's training
# step on CPU): the first comb conv is never reached by StyleGAN_zxc.forward's indexing (styleUnet.py:858-866), and with
# zero_noise=True every noise map except the first is zero, so those NoiseInjection weights get an exactly-zero gradient.
GEN_UNUSED = {"comb_convs.0.0.weight", "comb_convs.0.1.bias"} | {"convs.%d.noise.weight" % i for i in range(6)}


def test_stage_one_step_runs_and_updates_every_subnetwork():
    cfg = train_step.default_cfg(num_coarse=32, num_fine=8)
    step = train_step.StageOneStep(n_frames=4, cfg=cfg, patch=64, seed=0)
    batch = train_step.synthetic_batch(1, 4, "cuda", seed=0, patch=64)
    net = step.net
    groups = {"mlp": [net.model_coarse.layers_xyz, net.model_coarse.fc_alpha, net.model_coarse.fc_rgbFeat, net.model_coarse.fc_rgb],
              "xy": [net.model_coarse.XY_gen], "yz": [net.model_coarse.YZ_gen], "skin": [net.headpose_skin_net], "disc": [step.disc]}
    before = {k: _snap(v) for k, v in groups.items()}
    lat0 = net.latent_codes.detach().clone()
    losses = [step(batch) for _ in range(2)]
    torch.cuda.synchronize()
    assert all(torch.isfinite(l["loss"]) and torch.isfinite(l["d_loss"]) for l in losses)
    for k, v in groups.items():
        unmoved = _unmoved_names(before[k], v)
        assert unmoved == (GEN_UNUSED if k in ("xy", "yz") else set()), (k, sorted(unmoved))
    assert not torch.equal(lat0, net.latent_codes)
    assert all(torch.isfinite(p).all() for p in net.parameters())


def test_stage_two_step_runs_with_r1():
    step = train_step.StageTwoStep(n_frames=2, render_size=32, gen_size=128, d_reg_every=2, seed=0)
    batch = train_step.synthetic_batch(2, 2, "cuda", seed=1, render_size=32, gen_size=128)
    mods = {"gen": [step.generator], "disc": [step.disc], "nerf": [step.net.model_coarse]}
    before = {k: _snap(v) for k, v in mods.items()}
    ema0 = _snap([step.g_ema])
    outs = [step(batch) for _ in range(2)]
    torch.cuda.synchronize()
    # the R1 gate uses the 0-based iteration index: iteration 0 regularises (train_avatarHD.py:209)
    assert outs[0]["r1"] is not None and torch.isfinite(outs[0]["r1"]) and outs[1]["r1"] is None
    assert all(torch.isfinite(o["g_loss"]) and torch.isfinite(o["d_loss"]) for o in outs)
    for k, v in mods.items():
        unmoved = _unmoved_names(before[k], v)
        allowed = {"%s.%s" % (g, n) for g in ("XY_gen", "YZ_gen") for n in GEN_UNUSED} if k == "nerf" else set()
        assert unmoved <= allowed, (k, sorted(unmoved - allowed))
    assert _moved(ema0, [step.g_ema])[0] > 0


def test_graphed_stage_one_step_matches_eager():
    """The whole iteration (forward, fused render backward, convolution backward, Adam) captured as one CUDA graph must walk
    the same trajectory as the eager step: deterministic configuration (no perturbation / density noise, no discriminator), ten
    iterations each from the same seed."""
    def make(capturable):
        cfg = train_step.default_cfg(num_coarse=32, num_fine=8, perturb=False, noise_std=0.0)
        return train_step.StageOneStep(n_frames=4, cfg=cfg, patch=64, seed=0, with_discriminator=False, capturable=capturable)

    batch = train_step.synthetic_batch(1, 4, "cuda", seed=0, patch=64)
    eager = make(False)
    le = [float(eager(batch)["loss"]) for _ in range(10)]
    step = make(True)
    run = train_step.Graphed(step, batch, warmup=3)        # 3 eager warm-up iterations + the captured one = iterations 1..4
    lg = [float(run(batch)["loss"]) for _ in range(6)]     # iterations 5..10
    torch.cuda.synchronize()
    assert step.it == eager.it == 10
    assert le[-1] < le[0]                                  # the batch is being fitted
    for a, b in zip(le[4:], lg):
        assert abs(a - b) < 2e-2 * abs(a), (le, lg)
    we, wg = eager.net.model_coarse.layers_xyz[1].weight, step.net.model_coarse.layers_xyz[1].weight
    assert float((we - wg).abs().max()) < 5e-2 * float(we.abs().max())
    assert all(torch.isfinite(p).all() for p in step.net.parameters())


def test_graphed_stage_two_step_with_eager_r1():
    step = train_step.StageTwoStep(n_frames=2, render_size=32, gen_size=128, d_reg_every=2, seed=0, capturable=True)
    batch = train_step.synthetic_batch(2, 2, "cuda", seed=1, render_size=32, gen_size=128)
    run = train_step.Graphed(step, batch)
    g0 = _snap([step.generator])
    outs = [run(batch) for _ in range(4)]
    torch.cuda.synchronize()
    assert all(torch.isfinite(o["g_loss"]) and torch.isfinite(o["d_loss"]) for o in outs)
    assert any(o["r1"] is not None for o in outs)
    moved, total = _moved(g0, [step.generator])
    assert moved >= 0.9 * total


def test_overlapped_stage_two_iteration_equals_the_sequential_one():
    """StageTwoStep.dg_step (G-step forward on a side stream while the D step runs) against d_step(); g_step() on an iteration
    without the R1 pass: same weights, same supplied random draws -> same losses and the same updated weights up to the
    summation-order noise of the atomics in the render backward."""
    import numpy as np

    from havatar_b200 import synth

    def run(overlap):
        step = train_step.StageTwoStep(n_frames=2, render_size=32, gen_size=128, d_reg_every=4, seed=0, overlap=overlap)
        batch = train_step.synthetic_batch(2, 2, "cuda", seed=1, render_size=32, gen_size=128)
        for ph in ("d", "g"):
            r = synth.randoms(2, 32 * 32, 64, 16, seed=5 if ph == "d" else 6)
            batch["randoms_" + ph] = {k: torch.from_numpy(r[k]).cuda() for k in ("t_rand", "u_rand", "noise_coarse", "noise_fine")}
            batch["z_" + ph] = torch.from_numpy(synth.named_normal("z" + ph, (2, 64), 1)).cuda()
            batch["gen_noise_" + ph] = [torch.from_numpy(synth.named_normal("n%s%d" % (ph, i), (1, 1, 2 ** r_, 2 ** r_), 1)).cuda()
                                        for i, r_ in enumerate(r_ for r_ in range(4, 7) for _ in range(2))]
        outs = [step(batch) for _ in range(2)]           # iteration 0 regularises (sequential path); iteration 1 is the one under test
        torch.cuda.synchronize()
        assert ("dg" in [n for n, _, _ in step.parts()]) == overlap
        return outs[1], step

    a, sa = run(False)
    b, sb = run(True)
    for k in ("d", "g_loss", "rgb_loss", "hr_l1", "g_nonsat"):
        assert abs(float(a[k]) - float(b[k])) < 2e-3 * abs(float(a[k])) + 1e-6, (k, float(a[k]), float(b[k]))
    for ma, mb in ((sa.generator, sb.generator), (sa.disc, sb.disc), (sa.net.model_coarse.layers_xyz, sb.net.model_coarse.layers_xyz)):
        # Adam's first steps move a weight by ~lr * sign(g): entries whose gradient sign flips under reordering noise differ by
        # up to 2 lr per step; everything else agrees closely (fraction taken over the whole sub-network)
        diff = torch.cat([(pa - pb).detach().abs().reshape(-1) for pa, pb in zip(ma.parameters(), mb.parameters())])
        assert float(diff.max()) <= 4.1 * 1e-3, float(diff.max())
        assert float((diff > 2e-4).float().mean()) < 0.05


def test_overlapped_stage_one_discriminator_pass_equals_the_sequential_one():
    """StageOneStep(overlap=True) runs the patch discriminator's own forward / backward on a side stream while the render
    network's backward runs: same losses and weights as the sequential order (deterministic configuration)."""
    def run(overlap):
        cfg = train_step.default_cfg(num_coarse=32, num_fine=8, perturb=False, noise_std=0.0)
        step = train_step.StageOneStep(n_frames=4, cfg=cfg, patch=64, seed=0, overlap=overlap)
        batch = train_step.synthetic_batch(1, 4, "cuda", seed=0, patch=64)
        outs = [step(batch) for _ in range(3)]
        torch.cuda.synchronize()
        return outs, step

    (a, sa), (b, sb) = run(False), run(True)
    for x, y in zip(a, b):
        assert abs(float(x["loss"]) - float(y["loss"])) < 2e-3 * abs(float(x["loss"]))
        assert abs(float(x["d_loss"]) - float(y["d_loss"])) < 2e-3 * abs(float(x["d_loss"]))
    for ma, mb in ((sa.disc, sb.disc), (sa.net.model_coarse.layers_xyz, sb.net.model_coarse.layers_xyz)):
        diff = torch.cat([(pa - pb).detach().abs().reshape(-1) for pa, pb in zip(ma.parameters(), mb.parameters())])
        assert float((diff > 5e-4).float().mean()) < 0.05
