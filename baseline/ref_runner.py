"""Drive the UNMODIFIED reference (installed under baseline/_ref by baseline/install_ref.py) on the benchmark workload.

    python baseline/ref_runner.py --device cuda [--reps 3] [--out x.npz]     # the reference's GPU path (the >= 10x denominator)
    python baseline/ref_runner.py --device cpu --chunks 8                     # the reference's CPU path on a bounded sample

Prints ONE JSON line.  Nothing of havatar_b200's kernels, models or engine is on this path: the only repo module imported is
havatar_b200/synth.py (numpy-only seeded inputs, so both arms see the same frame and the same weights).  What runs is the
reference's `Trainer.nerf_forward` (model/nerf_trainer.py:38-92: the 4096-ray chunk loop over `predict_and_render_radiance`,
:120-201) with its plane generators bypassed -- the benchmark's planes are inputs already resident on the device, exactly as in
our arm -- at BASELINE.json configs[1]: 512 x 512 rays, 64 samples, coarse only, perturb off; fp32, TF32 off (torch default for
matmul).  `--hd` adds the reference's SWGAN_unet forward (model/styleUnet.py:1323-1410) and the whole HD frame
(Trainer.forward 'validation' with the plane generators + SWGAN_unet, avatarHD_reenactment.py:149-170)."""
import argparse
import json
import os
import sys
import time
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
H = W = 512
S = 64
CHUNK = 4096


def import_reference(device):
    """sys.path + the environment shims the reference needs here: its two extension modules by bare
    name, a stub for the absent matplotlib (utils/training_util.py:6; never used on this path), and on CPU the two device
    defaults the reference hard-codes to 'cuda' (model/network/embedder.py:99, model/styleUnet.py:748-751)."""
    import torch

    if not os.path.isdir(os.path.join(REF, "havatar", "model")):
        raise RuntimeError("baseline/_ref is not installed: run `python baseline/install_ref.py` in the build container")
    sys.path.insert(0, os.path.join(REF, "havatar"))
    sys.path.insert(0, os.path.join(REF, "ext"))
    sys.path.append(ROOT)
    try:
        import fused, upfirdn2d  # noqa: F401,E401  the reference's own model/op extensions, built for sm_100a
        ext = "reference model/op extensions loaded"
    except Exception as exc:  # CPU-only host: the wrappers branch to their torch fallbacks and never call the extension
        if device != "cpu":
            raise
        for name in ("fused", "upfirdn2d"):
            sys.modules[name] = types.ModuleType(name)
        ext = "extension modules stubbed on a CPU-only host (%s)" % type(exc).__name__
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    warnings.filterwarnings("ignore", message="conv2d_gradfix not supported")
    warnings.filterwarnings("ignore", message=".*indexing argument.*")
    if device == "cpu":
        import model.network.embedder as emb

        orig = emb.get_embedder
        emb.get_embedder = lambda multires, i=0, input_dims=3, include_input=True, device="cpu": orig(
            multires, i=i, input_dims=input_dims, include_input=include_input, device="cpu")
        torch.Tensor.cuda = lambda self, *a, **k: self
    import yaml
    from utils.cfgnode import CfgNode

    with open(os.path.join(REF, "havatar", "config", "singleview_512_base.yml")) as f:
        cfg = CfgNode(yaml.load(f, Loader=yaml.FullLoader))
    return cfg, ext


def build_trainer(cfg, dev, sc, render_size=512):
    import torch
    from model.nerf_trainer import Trainer

    cfg.models.StyleUnet.inp_size = render_size
    for mode in (cfg.nerf.train, cfg.nerf.validation):
        mode.num_coarse, mode.num_fine, mode.perturb, mode.radiance_field_noise_std, mode.chunksize = S, 0, False, 0.0, CHUNK
    net = Trainer(cfg, 4).to(dev).eval()
    mc = net.model_coarse
    mc.load_state_dict({k: torch.from_numpy(v).to(dev) for k, v in sc["weights"].items()}, strict=False)
    hs = net.headpose_skin_net
    hs.fix_canoW = True
    hs.canonical_W = torch.from_numpy(sc["wvol"]).to(dev)
    return net


def render_only(net, dev, sc, rows=None):
    """nerf_forward with set_conditional_embedding bypassed (planes = the benchmark's input planes)."""
    import torch

    mc = net.model_coarse
    mc.triPlane_embeddings = torch.from_numpy(sc["planes"]).to(dev)
    mc.set_conditional_embedding = lambda *a, **k: None
    rays, bg = torch.from_numpy(sc["ray_batch"]).to(dev), torch.from_numpy(sc["background_prior"]).to(dev)
    if rows is not None:
        rays, bg = rays[:, rows[0] * W:rows[1] * W].contiguous(), bg[:, rows[0] * W:rows[1] * W].contiguous()
    inv = torch.from_numpy(sc["inv_head_T"]).to(dev)

    def run():
        with torch.no_grad():
            return net.nerf_forward(ray_batch=rays, background_prior=bg, latent_code=None, inv_head_T=inv, front_render_cond=None,
                                    left_render_cond=None, right_render_cond=None, mode="validation")
    return run, rays.shape[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--chunks", type=int, default=8, help="cpu: 4096-ray chunks of the frame per repetition")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--hd", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import numpy as np
    import torch

    cfg, ext = import_reference(a.device)
    from havatar_b200 import synth

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device(a.device)
    sc = synth.scene(batch=1, height=H, width=W, seed=a.seed)
    sc["weights"], sc["wvol"] = synth.mlp_weights(0), synth.skin_volume(2)
    res = {"impl": "reference", "kind": "reference", "device": a.device, "ext": ext, "torch": torch.__version__}
    net = build_trainer(cfg, dev, sc)
    if a.device == "cpu":
        threads = a.threads or (os.cpu_count() or 1)
        torch.set_num_threads(threads)
        nrows = CHUNK // W * a.chunks
        run, n = render_only(net, dev, sc, rows=(H // 2 - nrows // 2, H // 2 - nrows // 2 + nrows))
        for _ in range(a.warmup):
            run()
        secs = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            out = run()
            secs.append(time.perf_counter() - t0)
        assert bool(torch.isfinite(out[0]).all())
        res.update(rays_per_sec=n / (sum(secs) / len(secs)), seconds=secs, rays=n, cores=threads,
                   sample="%d x %d-ray chunks (rows %d..%d of the 512x512 frame), 64 samples, unmodified reference Trainer.nerf_forward "
                          "on torch-CPU, %d threads" % (a.chunks, CHUNK, H // 2 - nrows // 2, H // 2 - nrows // 2 + nrows - 1, threads))
        print(json.dumps(res))
        return
    run, n = render_only(net, dev, sc)
    for _ in range(max(a.warmup, 1)):
        out = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    rgb, depth, acc = out[0], out[1], out[2]
    res.update(rays_per_sec=n / (ms * 1e-3), ms_per_frame=ms, rays=n, acc_mean=float(acc.mean()),
               what="unmodified reference Trainer.nerf_forward (4096-ray chunk loop over predict_and_render_radiance) on CUDA, fp32, "
                    "TF32 off, planes resident, full 512x512x64 frame")
    if a.out:
        np.savez(a.out, rgb=rgb.reshape(-1, 67)[::61].cpu().numpy(), acc=acc.reshape(-1)[::61].cpu().numpy(),
                 depth=depth.reshape(-1)[::61].cpu().numpy())
    del out, rgb, depth, acc
    if a.hd:
        from model.styleUnet import SWGAN_unet

        hd = {}
        g = torch.Generator(device=dev).manual_seed(1)
        for rs_, out_ in ((128, 512), (512, 1024)):
            torch.cuda.empty_cache()
            up = SWGAN_unet(inp_size=rs_, inp_ch=64, out_ch=3, out_size=out_, style_dim=64, n_mlp=4).to(dev).eval()
            cond = torch.randn(1, 64, rs_, rs_, device=dev, generator=g)
            z = torch.randn(1, 64, device=dev, generator=g)
            with torch.no_grad():
                for _ in range(2):
                    img = up([z], cond)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.reps):
                    img = up([z], cond)
                e1.record()
                torch.cuda.synchronize()
            entry = {"swgan_unet_ms": e0.elapsed_time(e1) / a.reps, "finite": bool(torch.isfinite(img).all())}
            # the whole HD frame: plane generators + render at rs_ x rs_ x 64 + upsampler
            sc_h = synth.scene(batch=1, height=rs_, width=rs_, seed=a.seed)
            sc_h["weights"], sc_h["wvol"] = sc["weights"], sc["wvol"]
            net_h = build_trainer(cfg, dev, sc_h, render_size=rs_)
            data = dict(mode="validation", fidx=None, render_full_img=True, ray_batch=torch.from_numpy(sc_h["ray_batch"]).to(dev),
                        background_prior=torch.from_numpy(sc_h["background_prior"]).to(dev), inv_head_T=torch.from_numpy(sc_h["inv_head_T"]).to(dev),
                        front_render_cond=torch.rand(1, 7, 256, 256, device=dev, generator=g),
                        left_render_cond=torch.rand(1, 7, 256, 256, device=dev, generator=g),
                        right_render_cond=torch.rand(1, 7, 256, 256, device=dev, generator=g))

            def frame():
                with torch.no_grad():
                    render, _, _ = net_h(**data)
                    return up([z], render[:, 3:])
            for _ in range(2):
                frame()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.reps):
                img = frame()
            e1.record()
            torch.cuda.synchronize()
            entry.update(ms_per_frame=e0.elapsed_time(e1) / a.reps, frames_per_sec=1e3 * a.reps / e0.elapsed_time(e1))
            hd["%d_to_%d" % (rs_, out_)] = entry
            del up, net_h, data, img
        res["hd"] = hd
    print(json.dumps(res))


if __name__ == "__main__":
    main()
