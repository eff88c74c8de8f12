"""Data formats and loader (havatar_b200/data.py) against the UNMODIFIED reference dataset class run on
the fixture dataset under tests/golden/dataset (tests/golden/dataset_items.npz, oracle/gen_golden.py::gen_dataset)."""
import os

import numpy as np
import pytest
import torch

from havatar_b200 import data


@pytest.fixture(scope="module")
def ds(golden_dir):
    return data.FrameDataset(os.path.join(golden_dir, "dataset", "sv_v31_all.json"), cond_res=32)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "dataset_items.npz"))


def test_item_order_and_metadata_match_the_reference(ds, gold):
    """(frame, view) list: view '8' skipped, sorted by fidx (dataloader.py:60-71); inv_head_T (:206,215-216) bit for bit."""
    assert len(ds) == int(gold["len"]) == 4
    for i in range(len(ds)):
        assert ds.fidx(i) == int(gold["item%d_fidx" % i])
        assert np.array_equal(ds.inv_head_T(i), gold["item%d_inv_head_T" % i])


def test_camera_parameters_reproduce_the_reference_rays(ds, gold):
    """The 18-float camera block carries what the reference's [H*W, 11] ray tensor is built from: origin = c2w[:, 3] and
    near / far bit for bit; directions through the oracle's get_rays to 1e-6 (closed-form K^-1 vs numpy's LU inverse)."""
    from oracle import render_oracle as ro

    for i in range(len(ds)):
        intr, c2w, near, far = ds.camera(i)
        rays = gold["item%d_mv_rays" % i]
        assert rays.shape == (256, 11)
        assert np.array_equal(np.broadcast_to(c2w[:, 3], (256, 3)), rays[:, :3])
        assert np.all(rays[:, 6] == near) and np.all(rays[:, 7] == far)
        _, d = ro.get_rays(ds.img_h, ds.img_w, intr, c2w)
        assert np.abs(d.reshape(-1, 3) - rays[:, 3:6]).max() < 1e-6
        assert np.all(rays[:, 8:11] == 1.0)                                   # white background (white_bg=True)


def test_condition_images_decode_to_the_reference_bytes(ds, gold, golden_dir):
    """uint8 decode (+ the INTER_LINEAR resize branch) == the reference's float tensors * 255."""
    for res, tag in ((32, "r32"), (24, "r24")):
        d2 = data.FrameDataset(os.path.join(golden_dir, "dataset", "sv_v31_all.json"), cond_res=res)
        for i in range(len(d2)):
            render, normal = d2.cond_uint8(i)
            for v, name in enumerate(data.VIEWS):
                ref = gold["item%d_%s_%s" % (i, name, tag)]                   # [res,res,7]
                assert np.array_equal((render[v].astype(np.float32) / 255.0), ref[..., :3])
                assert np.array_equal((normal[v].astype(np.float32) / 255.0), ref[..., 3:6])
                assert np.array_equal((normal[v].astype(np.int32).sum(-1) > 0).astype(np.float32), ref[..., 6])


def test_targets_match_the_reference(ds, gold):
    for i in range(len(ds)):
        rgb, m = ds.image_and_mask(i)
        assert np.array_equal(rgb.reshape(-1, 3), gold["item%d_gt" % i])


def test_checkpoint_layouts_and_partial_load():
    """train_avatar.py:296-307 / train_avatarHD.py:347-358 key sets; avatarHD_reenactment.py:138-146 loading."""
    from havatar_b200 import styleunet, train_step, trainer

    cfg = train_step.default_cfg(inp_size=32, out_size=128)
    net = trainer.Trainer(cfg, 3)
    gen = styleunet.SWGAN_unet(inp_size=32, inp_ch=64, out_ch=3, out_size=128, style_dim=64, n_mlp=4)
    disc = styleunet.Discriminator(128, img_channel=3)
    opts = [torch.optim.Adam(m.parameters(), lr=1e-3) for m in (net, gen, disc)]
    c1 = data.stage_one_checkpoint(5, net, opts[0], loss=0.1, psnr=20.0)
    assert set(c1) == {"iter", "optimizer_state_dict", "loss", "psnr", "trainer_state_dict"}
    c2 = data.stage_two_checkpoint(9, net, gen, disc, gen, *opts)
    assert set(c2) == {"iter", "nerf_optimizer", "g_optim", "d_optim", "nerf_render", "g", "d", "g_ema", "latent_codes"}
    with torch.no_grad():
        net.latent_codes.normal_()
        net.model_coarse.fc_rgb.weight.normal_()
    c2 = data.stage_two_checkpoint(9, net, gen, disc, gen, *opts)
    infer = trainer.Trainer(cfg, 0)                                            # avatarHD_reenactment.py:138
    up = styleunet.SWGAN_unet(inp_size=32, inp_ch=64, out_ch=3, out_size=128, style_dim=64, n_mlp=4)
    data.load_reenactment_checkpoint(c2, infer, up)
    assert torch.equal(infer.model_coarse.fc_rgb.weight, net.model_coarse.fc_rgb.weight)
    assert torch.equal(infer.latent_codes, net.latent_codes.data) and infer.headpose_skin_net.fix_canoW


@pytest.mark.gpu
def test_device_condition_tensors_and_cache(ds, gold):
    """hav_make_render_cond == make_render_cond_ (dataloader.py:218-229) bit for bit, channels first; second access is a hit."""
    loader = data.FrameLoader(ds)
    b = loader.batch([0, 1, 2, 3])
    torch.cuda.synchronize()
    for i in range(4):
        for name in data.VIEWS:
            got = b["%s_render_cond" % name][i].permute(1, 2, 0).cpu().numpy()
            assert np.array_equal(got, gold["item%d_%s_r32" % (i, name)]), (i, name)
    assert loader.cache.misses == 2 and loader.cache.hits == 2                # two instance directories, two views each
    loader.batch([0, 2])
    assert loader.cache.misses == 2


@pytest.mark.gpu
def test_loader_batch_renders_like_the_reference_ray_tensor(ds, gold):
    """A FrameLoader batch (camera block, rays generated in the kernel) renders what the reference's own ray tensor renders."""
    from havatar_b200 import render, synth

    loader = data.FrameLoader(ds)
    b = loader.batch([0, 3])
    sc = synth.scene(batch=2, crop=(0, 0, 16, 16), seed=81)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    w = {k: dev(v) for k, v in sc["weights"].items()}
    common = (b["inv_head_T"], dev(sc["planes"]), dev(sc["wvol"]), w, 16, 4)
    rays = dev(np.stack([gold["item0_mv_rays"][:, :8], gold["item3_mv_rays"][:, :8]]))
    a = render.render_rays(rays, b["background_prior"], *common, precision="fp32")
    c = render.render_rays(None, b["background_prior"], *common, precision="fp32", camera=b["camera"], img_hw=b["img_hw"])
    torch.cuda.synchronize()
    for k in ("rgb_coarse", "acc_coarse", "rgb_fine", "depth_fine"):
        assert float((getattr(a, k) - getattr(c, k)).abs().max()) < 2e-5, k    # directions agree to 1e-6, see the CPU test


@pytest.mark.gpu
def test_trainer_takes_a_loader_batch(ds, gold):
    """havatar_b200.trainer.Trainer.forward(mode='validation') on a FrameLoader batch (camera block) == the same call on the
    reference's ray tensor of that item."""
    from havatar_b200 import train_step, trainer

    cfg = train_step.default_cfg(num_coarse=16, num_fine=4, inp_size=16, out_size=64)
    torch.manual_seed(0)
    net = trainer.Trainer(cfg, 2).cuda().eval()
    loader = data.FrameLoader(ds)
    b = loader.batch([1])
    # the plane generators take 256 x 256 condition images; the fixture's are 32 x 32
    up = lambda t: torch.nn.functional.interpolate(t, size=(256, 256), mode="nearest")
    conds = {k: up(b[k]) for k in ("front_render_cond", "left_render_cond", "right_render_cond")}
    rays = torch.from_numpy(gold["item1_mv_rays"][None, :, :8].copy()).cuda()
    with torch.no_grad():
        r_cam, m_cam, _ = net(mode="validation", fidx=None, render_full_img=True, camera=b["camera"], img_hw=b["img_hw"],
                              background_prior=b["background_prior"], inv_head_T=b["inv_head_T"], **conds)
        r_ray, m_ray, _ = net(mode="validation", fidx=None, render_full_img=True, ray_batch=rays,
                              background_prior=b["background_prior"], inv_head_T=b["inv_head_T"], **conds)
    torch.cuda.synchronize()
    assert r_cam.shape == (1, 67, 16, 16) and torch.isfinite(r_cam).all()
    assert float((r_cam - r_ray).abs().max()) < 2e-3 and float((m_cam - m_ray).abs().max()) < 2e-3
