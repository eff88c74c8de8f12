"""HD inference frame (havatar_b200/pipeline.py): planes -> render -> upsampler wiring, CUDA-graph replay == eager."""
import numpy as np
import pytest
import torch

from havatar_b200 import pipeline, render, synth

pytestmark = pytest.mark.gpu


def test_hd_frame_wiring_and_graph_replay():
    torch.manual_seed(0)
    rs = 32
    sc = synth.scene(batch=1, height=rs, width=rs, seed=0)
    net = pipeline.AvatarHD(sc["weights"], sc["wvol"], render_size=rs, out_size=128, plane_res=32, cond_size=64, num_coarse=16).cuda()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    args = (dev(sc["ray_batch"]), dev(sc["background_prior"]), torch.zeros(1, 32, device="cuda"), dev(sc["inv_head_T"]),
            torch.rand(1, 7, 64, 64, device="cuda"), torch.rand(1, 7, 64, 64, device="cuda"), torch.rand(1, 7, 64, 64, device="cuda"),
            torch.randn(1, 64, device="cuda"))
    img, low = net.frame(*args)
    assert img.shape == (1, 3, 128, 128) and low.shape == (1, 3, rs, rs) and torch.isfinite(img).all()
    # the low-res rgb is exactly the fused render of the generated planes, pixel r <-> ray r (nerf_trainer.py:111-113)
    planes = net.planes(args[2], args[3], args[4], args[5], args[6])
    w = {k: dev(v) for k, v in sc["weights"].items()}
    o = render.render_rays(args[0], args[1], args[3], planes, dev(sc["wvol"]), w, 16, 0)
    assert torch.equal(low, o.rgb_coarse.view(1, rs, rs, 67).permute(0, 3, 1, 2)[:, :3])
    g = net.graphed(*args)
    img2, _ = g(*args)
    assert torch.equal(img2, img)
