"""Host-side drop-in surface: the `model.*` namespace shim the reference's entry scripts import
through, the attributes they touch, optimiser checkpoint round trips.  CPU only: nothing here launches a kernel."""
import sys

import numpy as np
import pytest
import torch

from havatar_b200 import compat, parallel, train_step, trainer


def test_model_namespace_shim_serves_the_entry_script_imports():
    """train_avatar.py:20, train_avatarHD.py:19-20,27, avatarHD_reenactment.py:6,14, utils/styleUnet_util.py:7."""
    names = compat.install()
    try:
        from model.nerf_trainer import Trainer
        from model.styleUnet import SWGAN_unet, Discriminator
        from model.op import conv2d_gradfix, FusedLeakyReLU, fused_leaky_relu, upfirdn2d   # noqa: F401
        import model.op.conv2d_gradfix as g
        import fused, upfirdn2d as ufd                                                    # noqa: E401  (model/op/fused_act.py:20)

        assert Trainer is trainer.Trainer and SWGAN_unet.__module__.startswith("havatar_b200")
        assert Discriminator.__module__.startswith("havatar_b200")
        assert hasattr(fused, "fused_bias_act") and hasattr(ufd, "upfirdn2d")
        assert g is conv2d_gradfix and not g.weight_gradients_disabled
        with g.no_weight_gradients():
            assert g.weight_gradients_disabled
        assert not g.weight_gradients_disabled
        assert set(names) >= {"model", "model.nerf_trainer", "model.styleUnet", "model.op.conv2d_gradfix"}
    finally:
        compat.uninstall()
    assert "model.nerf_trainer" not in sys.modules


def test_trainer_exposes_what_the_entry_scripts_touch():
    """train_avatar.py:95,98,124,311; avatarHD_reenactment.py:141-144; train_avatarHD.py:356."""
    net = trainer.Trainer(train_step.default_cfg(), 3)
    skin = net.headpose_skin_net
    for attr in ("pretrain_wc", "visualize_motion_weight_vol", "fix_canonical_W", "canonical_Wvolume", "sample_volume"):
        assert hasattr(skin, attr), attr
    assert net.latent_codes.shape == (3, 32) and hasattr(net, "model_coarse")
    assert skin.gridwarper.inv_trans(skin.gridwarper(torch.tensor([[0.3, 1.0, -0.2]]))).sub(torch.tensor([[0.3, 1.0, -0.2]])).abs().max() < 1e-6


def test_make_volume_pts_covers_the_skin_box():
    net = trainer.Trainer(train_step.default_cfg(), 1)
    pts = trainer.make_volume_pts(steps=5, perturb=False, gridwarper=net.headpose_skin_net.gridwarper)
    assert pts.shape == (125, 3)
    lo, hi = pts.min(0).values.numpy(), pts.max(0).values.numpy()
    assert np.allclose(lo, [-1.5, 0.42, -1.6], atol=1e-5) and np.allclose(hi, [1.5, 1.4, 1.2], atol=1e-5)   # nerf_trainer.py:29-34


def test_flat_adam_state_dict_round_trips_through_torch_adam_layout():
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7))]
    opt = parallel.FlatAdam(ps, 1e-3, betas=(0.5, 0.9))
    opt.m.normal_(), opt.v.uniform_(0.1, 1.0)
    opt.state[0] = 12.0
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == [0, 1]
    assert float(sd["state"][1]["step"]) == 12.0 and sd["state"][0]["exp_avg"].shape == (3, 5)
    # a torch.optim.Adam over the same shapes accepts it ...
    ref = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1e-3, betas=(0.5, 0.9))
    ref.load_state_dict(sd)
    # ... and its own state_dict loads back into a fresh FlatAdam
    opt2 = parallel.FlatAdam([torch.nn.Parameter(p.detach().clone()) for p in ps], 5e-2)
    opt2.load_state_dict(ref.state_dict())
    for (p, o) in zip(opt.layout.params, opt.layout.offsets):          # parameter slots (the 128-byte padding between them is unused)
        sl = slice(o, o + p.numel())
        assert torch.equal(opt2.m[sl], opt.m[sl]) and torch.equal(opt2.v[sl], opt.v[sl])
    assert float(opt2.state[0]) == 12.0 and abs(float(opt2.state[1]) - 1e-3) < 1e-9 and opt2.betas == (0.5, 0.9)
    assert abs(opt2.param_groups[0]["lr"] - 1e-3) < 1e-9


def test_stage_schedules_follow_the_reference_indexing():
    """train_avatar.py:154-158: lr(i) is set AFTER step i, so step i runs with lr(i-1); train_avatarHD.py:205-209: the GAN
    weight and the R1 gate use the 0-based iteration index (iteration 0 regularises)."""
    class G:
        def __init__(self):
            self.lrs = []

        def set_lr(self, lr):
            self.lrs.append(lr)

    s1 = train_step.StageOneStep.__new__(train_step.StageOneStep)
    s1.cfg, s1.it, s1.lr0, s1.g = train_step.default_cfg(), 0, 5e-4, G()
    for _ in range(3):
        s1.pre_step()
    assert s1.g.lrs == [max(5e-4 * 0.1 ** (i / 250000.0), 5e-5) for i in (0, 1)]      # iterations 1 and 2; iteration 0 keeps lr0
    s2 = train_step.StageTwoStep.__new__(train_step.StageTwoStep)
    s2.it, s2.d_reg_every, s2.gan_w = 0, 16, torch.zeros(())
    gates = []
    for _ in range(18):
        s2.pre_step()
        gates.append((s2.it - 1) % s2.d_reg_every == 0)
    assert [i for i, g in enumerate(gates) if g] == [0, 16]
    assert abs(float(s2.gan_w) - 1e-3) < 1e-9


def test_pretrain_wc_and_sample_volume_match_the_reference(golden_dir):
    """model/Skinning_Field.py:101-125 / :65-68: two Adam iterations of the head-box fit from the same weights and the same
    torch generator state as the unmodified reference (tests/golden/skin_pretrain.npz, oracle/gen_golden.py::gen_skin)."""
    import os

    from havatar_b200 import synth

    g = np.load(os.path.join(golden_dir, "skin_pretrain.npz"))
    cfg = train_step.default_cfg()
    net = trainer.Trainer(cfg, 2)
    sd = synth.trainer_state({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=11)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    skin = net.headpose_skin_net
    torch.manual_seed(77)
    skin.pretrain_wc(num_iter=2, vol_thr=cfg.models.coarse.Head_bounding)
    with torch.no_grad():
        vol = skin.canonical_Wvolume()
        smp = skin.sample_volume(torch.from_numpy(g["pts"]))
    assert np.abs(vol.numpy()[:, :, ::8, ::8, ::8] - g["vol"]).max() < 1e-5
    assert np.abs(smp.numpy() - g["sample"]).max() < 1e-5
    assert np.abs(dict(skin.named_parameters())["canonical_Wvolume.final_conv.bias"].detach().numpy() - g["b0"]).max() < 1e-6
