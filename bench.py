#!/usr/bin/env python
"""bench.py -- rays/s of the fused volumetric render on B200 (BASELINE.json metric, configs[1]:
full 512x512 frame, 64 samples/ray, stage-one NeRF render, synthetic 3DMM inputs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one render of one 512x512 frame (262144 rays x 64 samples) per GPU.  N > 1 is launched by torchrun,
one rank per GPU, every rank renders its own frame (weak scaling, no data-path collective: rays are
independent -- DESIGN.md "multi-GPU").  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 512
S = 64
FLOP_PER_SAMPLE = 94848          # 2 * (176*128 + 128*128 + 128*1 + 128*64 + 64*3): model/nerf_model.py:46-51
BYTES_PER_RAY = 365              # algorithmic HBM bytes per ray at S = 64
METRIC = "rays_per_sec_512x512x64"
CPU_CHUNK_RAYS = 4096            # one reference chunk (nerf.validation.chunksize)


_STDOUT_FD = []


def emit(line):
    """Print the one JSON line on the real stdout (restoring it first if it was parked on stderr)."""
    sys.stdout.flush()
    if _STDOUT_FD:
        os.dup2(_STDOUT_FD.pop(), 1)
    print(json.dumps(line), flush=True)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_burst": d.get("bf16_tflops", 1590.0), "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows)}


def cpu_reference_rate(chunks, threads):
    """rays/s of the CPU port of the reference path (oracle/render_oracle_torch.py) on `chunks` 4096-ray chunks
    of the benchmark frame.  Returns (rays_per_s, seconds, sample description)."""
    import numpy as np
    import torch

    from havatar_b200 import synth
    from oracle import render_oracle as ro
    from oracle import render_oracle_torch as rt

    torch.set_num_threads(threads)
    rows = CPU_CHUNK_RAYS // W
    sc = synth.scene(batch=1, height=H, width=W, crop=(H // 2 - rows * chunks // 2, 0, rows * chunks, W), seed=0)
    args = (sc["ray_batch"], sc["background_prior"], sc["inv_head_T"], sc["planes"], sc["wvol"], sc["weights"],
            ro.default_boxes(), S, 0)
    warm = dict(sc, ray_batch=sc["ray_batch"][:, :CPU_CHUNK_RAYS], background_prior=sc["background_prior"][:, :CPU_CHUNK_RAYS])
    rt.render_rays(warm["ray_batch"], warm["background_prior"], *args[2:], chunk=CPU_CHUNK_RAYS)
    t0 = time.perf_counter()
    out = rt.render_rays(*args, chunk=CPU_CHUNK_RAYS)
    dt = time.perf_counter() - t0
    assert np.isfinite(out["rgb_coarse"]).all()
    n = rows * chunks * W
    return n / dt, dt, "%d x %d-ray chunks (rows %d..%d of the 512x512 frame), 64 samples, torch-CPU port, %d threads" % (
        chunks, CPU_CHUNK_RAYS, H // 2 - rows * chunks // 2, H // 2 + rows * chunks // 2 - 1, threads)


def _ref_installed():
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "havatar", "model"))


def _ref_child(argv, timeout):
    """baseline/ref_runner.py in a fresh process (the unmodified reference; none of our kernels in that process)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_runner.py")] + argv, capture_output=True, text=True,
                       timeout=timeout)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError("ref_runner failed (rc %d): %s" % (r.returncode, (r.stderr or r.stdout)[-400:]))
    return json.loads(lines[-1])


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores -- the UNMODIFIED reference
    installed under baseline/_ref (baseline/install_ref.py), Trainer.nerf_forward's chunk loop on torch-CPU with every host
    thread; each step is a bounded sample of the benchmark frame (8 x 4096-ray chunks = 1/8 of the 512x512x64 frame).  When
    baseline/_ref is absent the torch-CPU port under oracle/ stands in (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rates, secs, sample, kind = [], [], "", "port"
    sys.stdout.flush()                      # the reference's constructors print to stdout: park fd 1 on stderr until the JSON line
    _STDOUT_FD.append(os.dup(1))
    os.dup2(2, 1)
    if _ref_installed():
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import torch

        import ref_runner as rr
        from havatar_b200 import synth

        kind = "reference"
        chunks = 8 if args.warmup + args.steps <= 30 else 2
        cfg, _ = rr.import_reference("cpu")
        torch.set_num_threads(threads)
        sc = synth.scene(batch=1, height=H, width=W, seed=0)
        sc["weights"], sc["wvol"] = synth.mlp_weights(0), synth.skin_volume(2)
        net = rr.build_trainer(cfg, torch.device("cpu"), sc)
        nrows = CPU_CHUNK_RAYS // W * chunks
        r0 = H // 2 - nrows // 2
        run, n = rr.render_only(net, torch.device("cpu"), sc, rows=(r0, r0 + nrows))
        sample = "%d x %d-ray chunks (rows %d..%d of the 512x512 frame), 64 samples, unmodified reference Trainer.nerf_forward on torch-CPU, %d threads" % (
            chunks, CPU_CHUNK_RAYS, r0, r0 + nrows - 1, threads)
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            out = run()
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                rates.append(n / dt), secs.append(dt)
        assert bool(torch.isfinite(out[0]).all())
    else:
        chunks = 2 if args.warmup + args.steps <= 30 else 1
        for i in range(args.warmup + args.steps):
            r, dt, sample = cpu_reference_rate(chunks, threads)
            if i >= args.warmup:
                rates.append(r), secs.append(dt)
    value = sum(rates) / len(rates)
    line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "stage-one NeRF render, 512x512 frame x 64 samples/ray, coarse only (BASELINE.json configs[1])",
                       "step": "bounded sample: " + sample},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_train_legs(args, torch, dist, dev, rank, world, barrier, headline):
    """The two training iterations (BASELINE.json configs[2] and configs[4]) on every rank.  A watchdog guards the headline: if
    these legs have not finished by the deadline (a wedged collective, say), rank 0 prints the headline JSON line it already has
    with train = {"error": ...} and every rank exits, so the render numbers can never be lost to the secondary legs."""
    import threading

    def bail():
        if headline is not None:
            out = dict(headline)
            out.pop("train_note", None)
            out["train"] = {"error": "training legs did not finish within %d s; skipped" % args.train_deadline}
            emit(out)
        os._exit(0)

    dog = threading.Timer(args.train_deadline, bail)
    dog.daemon = True
    dog.start()
    # ---- the two training steps (BASELINE.json configs[2] and configs[4]): per-rank frames, gradients all-reduced over NCCL
    #      in buckets from inside backward (havatar_b200/parallel.py); device-timed, max over ranks
    train = {}
    if not args.no_train:
      try:
        from havatar_b200 import train_step

        torch.cuda.empty_cache()
        modes = ("eager", "graph") if args.train_mode == "both" else (args.train_mode,)
        # stage two runs its R1 pass every 16th iteration: warm up through the first one, then time one full period of 16
        for name, mk, mkb, frames, n_warm, n_t in (
                ("stage_one_b4_patch64", lambda c: train_step.StageOneStep(n_frames=4 * world, device=dev, capturable=c),
                 lambda: train_step.synthetic_batch(1, 4, dev, seed=rank, patch=64, frame_offset=4 * rank), 4, 3, max(3, min(args.steps, 10))),
                ("stage_two_b1_128_to_512", lambda c: train_step.StageTwoStep(n_frames=world, device=dev, capturable=c),
                 lambda: train_step.synthetic_batch(2, 1, dev, seed=rank, render_size=128, gen_size=512, frame_offset=rank), 1, 16, 16)):
            entry = {"frames_per_gpu": frames}
            for mode in modes:
                st, batch = mk(mode == "graph"), mkb()
                run = train_step.Graphed(st, batch) if mode == "graph" else st
                for _ in range(n_warm):
                    run(batch)
                barrier()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                for _ in range(n_t):
                    res_t = run(batch)
                b_.record()
                barrier()
                ms = torch.tensor([a_.elapsed_time(b_) / n_t], dtype=torch.float64, device=dev)
                if dist is not None:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                ok = all(bool(torch.isfinite(v)) for v in res_t.values() if v is not None)
                syncs = [g_.sync for g_ in st.groups() if g_.sync is not None]
                if dist is not None:      # replicas must still hold identical weights after the timed iterations
                    chk = torch.stack([p_.detach().double().sum() for g_ in st.groups() for p_ in g_.params[:8]])
                    lo_, hi_ = chk.clone(), chk.clone()
                    dist.all_reduce(lo_, op=dist.ReduceOp.MIN), dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
                    ok = ok and bool((hi_ - lo_).abs().max() <= 1e-9 * hi_.abs().max().clamp_min(1.0))
                entry[mode] = {"ms_per_step": float(ms), "frames_per_sec": world * frames * 1e3 / float(ms), "steps": n_t, "finite_and_in_sync": ok}
                entry["allreduce_bytes_per_step"] = sum(s_.bytes_per_step for s_ in syncs)
                entry["allreduce_buckets"] = sum(len(s_.buckets) for s_ in syncs)
                del st, batch, run
                torch.cuda.empty_cache()
            best = min((entry[m] for m in modes), key=lambda e: e["ms_per_step"])
            entry.update(ms_per_step=best["ms_per_step"], frames_per_sec=best["frames_per_sec"])
            # algorithmic tensor work of one iteration (forward = 1x, backward = 2x, the render backward's recompute = 1x more):
            # stage one: 4 frames x (plane generators 327.5 G x 3 + 4096 rays x 112 samples x 94 848 x 4); stage two: 1 frame x
            # (generators x 4: D-step forward + G-step forward/backward) + 128^2 rays x 112 samples x 94 848 x 5 + SWGAN_unet
            # 176.3 G x 4 + Discriminator(512) 71.5 G/image x (2 images x 3 + 1 image x 3)
            rs_ = 4096 * 112 * FLOP_PER_SAMPLE / 1e9
            gfl = 4 * (327.5 * 3 + rs_ * 4) if name.startswith("stage_one") else (327.5 * 4 + 4 * rs_ * 5 + 176.3 * 4 + 71.5 * 9)
            pk = _peaks()["bf16_sustained"]
            entry["roofline"] = {"bound": "tensor", "gflop_per_step": gfl, "achieved": gfl / best["ms_per_step"], "peak": pk, "unit": "TFLOP/s",
                                 "frac": gfl / best["ms_per_step"] / pk, "peak_source": "sustained bf16 (kernels timed inside a long step)"}
            train[name] = entry
        # the reference's own formulation of the stage-one iteration on this GPU, as far as it can be had here: render through
        # the ATen call sequence + torch autograd (oracle/render_oracle_torch.py, 2048-ray chunks like chunksize // B), cuDNN
        # convolutions through autograd, torch.optim.Adam, eager launches.  Rank 0 only, no gradient exchange.
        if rank == 0 and world == 1:
            try:
                train["stage_one_b4_patch64"]["reference_formulation"] = reference_formulation_stage_one(dev, train_step)
            except Exception as exc:
                train["stage_one_b4_patch64"]["reference_formulation"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
      except Exception as exc:
        train["error"] = "%s: %s" % (type(exc).__name__, exc)

    dog.cancel()
    return train


def reference_formulation_stage_one(dev, train_step):
    import torch

    from havatar_b200 import render as hrender
    from havatar_b200 import styleunet_train
    from oracle import render_oracle_torch as rt

    def aten_render(ray_batch, bg, inv_head_T, planes, wvol, weights, num_coarse, num_fine=0, boxes=None, t_rand=None,
                    noise_coarse=None, u_rand=None, noise_fine=None, precision=None):
        bx = tuple(torch.tensor(b, dtype=torch.float32, device=ray_batch.device) for b in boxes)
        o = rt.render_rays.__wrapped__(ray_batch, bg, inv_head_T, planes, wvol, weights, bx, num_coarse, num_fine, t_rand=t_rand,
                                       noise_coarse=noise_coarse, u_rand=u_rand, noise_fine=noise_fine, chunk=2048,
                                       device=ray_batch.device, to_numpy=False)
        un = lambda t: None if t is None else (t.unsqueeze(-1) if t.dim() == 2 else t)
        return hrender.RenderOut(o["rgb_coarse"], un(o["depth_coarse"]), un(o["acc_coarse"]), un(o["weights_max"]), o["rgb_fine"],
                                 un(o["depth_fine"]), un(o["acc_fine"]), None)

    saved, train_step.FLAT_ADAM[0] = hrender.render_rays_autograd, False
    hrender.render_rays_autograd = aten_render
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        st = train_step.StageOneStep(n_frames=4, device=dev)
        batch = train_step.synthetic_batch(1, 4, dev, seed=0, patch=64)
        with styleunet_train.library_convs():
            for _ in range(2):
                st(batch)
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(3):
                res = st(batch)
            b_.record()
            torch.cuda.synchronize()
        return {"ms_per_step": a_.elapsed_time(b_) / 3, "finite": bool(torch.isfinite(res["loss"])),
                "what": "same iteration with the reference's formulation: ATen render + torch autograd (fp32, 2048-ray chunks), cuDNN "
                        "convolutions (TF32 allowed, torch default), torch.optim.Adam, eager; our upfirdn2d / fused_bias_act kernels"}
    finally:
        hrender.render_rays_autograd = saved
        train_step.FLAT_ADAM[0] = True
        torch.cuda.empty_cache()


def run_ours(args):
    import numpy as np
    import torch

    from havatar_b200 import _lib, render, synth

    _lib.lib()  # fail loudly if the CUDA library is missing
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (havatar_b200 has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # NCCL writes its version banner to the process's stdout (fd 1) whatever NCCL_DEBUG_FILE says: park fd 1 on stderr
        # until the JSON line is due, so that stdout carries exactly one line
        sys.stdout.flush()
        _STDOUT_FD.append(os.dup(1))
        os.dup2(2, 1)

        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic frame of this rank (weak scaling: one frame per GPU per step, different head pose per rank)
    sc = synth.scene(batch=1, height=H, width=W, seed=rank)
    sc["weights"] = synth.mlp_weights(0)      # same model on every rank
    sc["wvol"] = synth.skin_volume(2)
    R = H * W
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host = {k: pin(sc[k]) for k in ("ray_batch", "background_prior", "inv_head_T", "planes")}
    d = {k: v.to(dev) for k, v in host.items()}
    # the same camera as 18 floats: the e2e legs upload this instead of the 8.4 MB ray tensor
    intr, c2w, near, far = synth.camera_params(H, W)
    cam_host = render.camera_block(intr, c2w, near, far, device="cpu").pin_memory()
    host_cam = dict(background_prior=host["background_prior"], inv_head_T=host["inv_head_T"], planes=host["planes"], camera=cam_host,
                    img_hw=(H, W))
    wts = {k: torch.from_numpy(v).to(dev) for k, v in sc["weights"].items()}
    wvol = torch.from_numpy(sc["wvol"]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step(out=None, reuse=False):
        return render.render_rays(d["ray_batch"], d["background_prior"], d["inv_head_T"], d["planes"], wvol, wts, S, 0,
                                  precision=args.precision, out=out, reuse_packed=reuse)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(max(args.warmup, 3)):
        out = step(out)
    barrier()

    # ---- timed region 1: device-resident inputs, K steps, per-step CUDA events, L2 flushed between steps.  Each iteration
    #      times the full step (pack weights + pack planes + render) and then the dominant kernel alone (weights / planes
    #      already packed: the roofline numerator), interleaved so both see the same clock / power conditions.
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0 = time.time()
    for (a, b), (c, e_) in zip(ev, kev):
        flush.zero_()
        a.record()
        out = step(out)
        b.record()
        flush.zero_()
        c.record()
        out = step(out, reuse=True)
        e_.record()
    barrier()
    t1 = time.time()
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    launches = args.steps * 4     # inside the timed step events: pack_mlp_fp32 + pack_mlp_16 + pack_planes + render_tc2 per step
    clocks = sampler.stop(t0, time.time()) if sampler is not None else None
    acc_mean = float(out.acc_coarse.mean())
    # "matched PSNR": the 16-bit tensor-core render against the fp32 CUDA-core (reference-exact) mode on the same frame
    psnr = None
    if args.precision != "fp32" and rank == 0:
        # on the 16 middle rows of the frame (8192 rays): the fp32 CUDA-core mode is a validation path, ~100x slower
        lo, hi = (H // 2 - 8) * W, (H // 2 + 8) * W
        sub = lambda t: t[:, lo:hi].contiguous()
        ref32 = render.render_rays(sub(d["ray_batch"]), sub(d["background_prior"]), d["inv_head_T"], d["planes"], wvol, wts, S, 0,
                                   precision="fp32")
        got = out.rgb_coarse[:, lo:hi]
        mse = float(((got[..., :3] - ref32.rgb_coarse[..., :3]) ** 2).mean())
        psnr = {"rgb_psnr_db_vs_fp32_mode": -10.0 * __import__("math").log10(max(mse, 1e-20)), "rays_compared": hi - lo,
                "max_abs_err_67ch": float((got - ref32.rgb_coarse).abs().max()),
                "max_abs_err_acc": float((out.acc_coarse[:, lo:hi] - ref32.acc_coarse).abs().max())}
        del ref32

    # ---- the reference-precision mode on the tensor cores (HAV_PREC_FP16X3: fp16 hi + lo split operands, fp32 everything else),
    #      same frame, same timing discipline (whole step: packing + render; L2 flushed between steps)
    x3 = None
    if rank == 0:
        try:
            o3 = None
            for _ in range(3):
                o3 = render.render_rays(d["ray_batch"], d["background_prior"], d["inv_head_T"], d["planes"], wvol, wts, S, 0,
                                        precision="fp16x3", out=o3)
            torch.cuda.synchronize()
            n3 = max(3, min(args.steps, 10))
            ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n3)]
            for a3, b3 in ev3:
                flush.zero_()
                a3.record()
                o3 = render.render_rays(d["ray_batch"], d["background_prior"], d["inv_head_T"], d["planes"], wvol, wts, S, 0,
                                        precision="fp16x3", out=o3)
                b3.record()
            torch.cuda.synchronize()
            ms3 = sum(a3.elapsed_time(b3) for a3, b3 in ev3) / n3
            lo, hi = (H // 2 - 8) * W, (H // 2 + 8) * W
            sub = lambda t: t[:, lo:hi].contiguous()
            r32 = render.render_rays(sub(d["ray_batch"]), sub(d["background_prior"]), d["inv_head_T"], d["planes"], wvol, wts, S, 0,
                                     precision="fp32")
            x3 = {"value": R / (ms3 * 1e-3), "unit": "rays/s", "ms_per_step": ms3, "dtype": "fp16 hi+lo split operands (3 MMAs per product), fp32 accumulate / planes / encoding / composite",
                  "kernel": "tc3::render_tc3_kernel<2> (CTA pairs, tcgen05 cta_group::2)",
                  "max_abs_err_67ch_vs_fp32_mode": float((o3.rgb_coarse[:, lo:hi] - r32.rgb_coarse).abs().max()),
                  "max_abs_err_acc_vs_fp32_mode": float((o3.acc_coarse[:, lo:hi] - r32.acc_coarse).abs().max()),
                  "tensor_tflops_3x_count": 3 * R * S * FLOP_PER_SAMPLE / (ms3 * 1e-3) / 1e12}
            x3_rgb = o3.rgb_coarse.reshape(-1, 67)[::61].cpu().numpy()
            del o3, r32
        except Exception as exc:
            x3 = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- timed region 2: end to end through the host-buffer API: every step uploads its inputs from pinned host
    #      memory and downloads rendered maps to pinned host memory; copies of neighbouring frames overlap the
    #      render (PipelinedHostRenderer: H2D / compute / D2H streams, two buffer sets).  Two variants: ALL maps (the
    #      67-channel colour + feature map of the pass, depth, acc, weights_max: 73 MB per frame, PCIe-bound) and the
    #      IMAGE maps the reference's callers read back on the host (rgb[..., :3], depth, acc: train_avatar.py:182-218;
    #      the 64 feature channels are consumed on the device by the StyleUNet, avatarHD_reenactment.py:160-166).
    #      The image variant is the headline `e2e`; the all-maps variant is reported next to it.
    def e2e_leg(maps, inputs):
        r = render.PipelinedHostRenderer(sc["weights"], sc["wvol"], S, 0, precision=args.precision, device=dev, maps=maps)
        for _ in range(3):
            r.submit(**inputs)
        r.drain()
        barrier()
        t_ = time.perf_counter()
        for _ in range(args.steps):
            r.submit(**inputs)
        out_ = r.drain()
        torch.cuda.synchronize()
        ms_ = (time.perf_counter() - t_) * 1e3 / args.steps
        # rays generated in the kernel differ from the uploaded ray tensor in the last bit of the directions only
        assert abs(float(out_["acc_coarse"].mean()) - acc_mean) < 1e-4
        res_ = (ms_, r.h2d_bytes, r.d2h_bytes, out_["rgb_coarse"].shape[-1])
        r.close()
        return res_

    e2e_all_ms, _, d2h_all, _ = e2e_leg("all", host_cam)
    e2e_ms, h2d, d2h, nch = e2e_leg("image", host_cam)
    assert nch == 3
    e2e_rays_ms, h2d_rays, _, _ = e2e_leg("image", host)          # the round-1 form: ray tensor uploaded per frame
    # the same, strictly serial (no overlap between frames), for reference
    hs = render.HostRenderer(sc["weights"], sc["wvol"], S, 0, precision=args.precision, device=dev)
    hs(**host)
    e0 = time.perf_counter()
    for _ in range(args.steps):
        hs(**host)
    e2e_serial_ms = (time.perf_counter() - e0) * 1e3 / args.steps

    # ---- secondary metric of BASELINE.json: HD frames/s (plane generators + full-frame render + StyleUNet upsampler,
    #      CUDA-graph replay; configs[3] shape 512^2 -> 1024^2 and the reference's default 128^2 -> 512^2)
    hd = {}
    if not args.no_hd:
      try:
        from havatar_b200 import pipeline

        for rs_, out_, bsz in ((512, 1024, 1), (128, 512, 1), (128, 512, 4)):
              sc_h = synth.scene(batch=bsz, height=rs_, width=rs_, seed=rank)
              net = pipeline.AvatarHD(sc["weights"], sc["wvol"], render_size=rs_, out_size=out_, precision=args.precision).to(dev)
              g = torch.Generator(device=dev).manual_seed(1)
              a_h = (torch.from_numpy(sc_h["ray_batch"]).to(dev), torch.from_numpy(sc_h["background_prior"]).to(dev),
                     torch.zeros(bsz, 32, device=dev), torch.from_numpy(sc_h["inv_head_T"]).to(dev),
                     torch.rand(bsz, 7, 256, 256, device=dev, generator=g), torch.rand(bsz, 7, 256, 256, device=dev, generator=g),
                     torch.rand(bsz, 7, 256, 256, device=dev, generator=g), torch.randn(bsz, 64, device=dev, generator=g))
              gf = net.graphed(*a_h)
              for _ in range(3):
                  gf(*a_h)
              barrier()
              hev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
              for a_, b_ in hev:
                  flush.zero_()
                  a_.record()
                  img, _ = gf(*a_h)
                  b_.record()
              barrier()
              hms = sum(a_.elapsed_time(b_) for a_, b_ in hev) / args.steps
              assert bool(torch.isfinite(img).all())
              # algorithmic FLOPs of a frame: both plane generators 327.5 G, render rs^2 x 64 samples x
              # 94 848, SWGAN_unet 176.3 G (128 -> 512) / 352.3 G (512 -> 1024); tensor roofline = measured burst bf16 peak
              gfl = bsz * (327.5 + rs_ * rs_ * S * FLOP_PER_SAMPLE / 1e9 + (352.3 if out_ == 1024 else 176.3))
              key = "%d_to_%d" % (rs_, out_) + ("" if bsz == 1 else "_batch%d" % bsz)
              hd[key] = {"ms_per_step": hms, "frames_per_step_per_gpu": bsz, "ms_per_frame": hms / bsz, "frames_per_sec": world * bsz * 1e3 / hms,
                         "gflop_per_step": gfl,
                         "roofline": {"bound": "tensor", "achieved": gfl / hms, "peak": _peaks()["bf16_burst"],
                                      "unit": "TFLOP/s", "frac": gfl / hms / _peaks()["bf16_burst"]}}
              del net, gf
      except Exception as exc:  # the secondary metric must never take the headline line down with it
        hd["error"] = "%s: %s" % (type(exc).__name__, exc)

    t = torch.tensor([step_ms, kern_ms, e2e_ms, e2e_all_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms, e2e_ms, e2e_all_ms = [float(v) for v in t.tolist()]
    if rank != 0:
        tr = run_train_legs(args, torch, dist, dev, rank, world, barrier, None)
        if dist is not None:
            if "error" in tr:
                os._exit(0)
            dist.destroy_process_group()
        return

    peaks = _peaks()
    achieved = R * S * FLOP_PER_SAMPLE / (kern_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "render_tc_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    threads = os.cpu_count() or 1
    cpu_kind = "port"
    try:
        if args.no_reference_gpu:
            raise RuntimeError("skipped (--no-reference-gpu)")
        if _ref_installed():      # the unmodified reference on this box's host cores (bounded sample: 1/8 of the frame x 3)
            rc_ = _ref_child(["--device", "cpu", "--chunks", "8", "--reps", "3", "--warmup", "1"], 600)
            cpu_rate, cpu_s, cpu_sample, cpu_kind = rc_["rays_per_sec"], sum(rc_["seconds"]), rc_["sample"], "reference"
        else:
            cpu_rate, cpu_s, cpu_sample = cpu_reference_rate(4, threads)
    except Exception as exc:
        cpu_rate, cpu_s, cpu_sample = None, None, "failed: %s" % exc
    # the reference's own GPU path: the UNMODIFIED reference (baseline/_ref: its Python modules + its model/op extensions built
    # for sm_100a by baseline/install_ref.py) on this GPU, in its own process -- Trainer.nerf_forward's 4096-ray chunk loop at
    # the same frame / weights / planes, fp32; plus its SWGAN_unet and whole HD frame.  This is the >= 10x denominator.
    ref_real = None
    if _ref_installed() and world == 1 and not args.no_reference_gpu:
        try:
            tmp = os.path.join("/tmp", "hav_ref_gpu_%d.npz" % os.getpid())
            ref_real = _ref_child(["--device", "cuda", "--reps", "3", "--hd", "--out", tmp], 900)
            z_ = np.load(tmp)
            mine = out.rgb_coarse.reshape(-1, 67)[::61].cpu().numpy()
            ref_real["max_abs_diff_vs_ours_67ch"] = float(np.abs(mine - z_["rgb"]).max())
            ref_real["max_abs_diff_vs_ours_acc"] = float(np.abs(out.acc_coarse.reshape(-1)[::61].cpu().numpy() - z_["acc"]).max())
            ref_real["rays_compared"] = int(z_["rgb"].shape[0])
            if x3 is not None and "error" not in x3:
                x3["max_abs_diff_vs_reference_gpu_67ch"] = float(np.abs(x3_rgb - z_["rgb"]).max())
                x3["speedup_vs_reference_gpu"] = x3["value"] / ref_real["rays_per_sec"]
            ref_real["value"], ref_real["unit"] = ref_real["rays_per_sec"], "rays/s"
            os.remove(tmp)
        except Exception as exc:
            ref_real = {"error": "%s: %s" % (type(exc).__name__, str(exc)[-300:])}
    # the reference's own GPU path, as far as it can be had on this box: the same ATen call sequence the reference executes
    # (oracle/render_oracle_torch.py: grid_sample, Linear, cumprod, ..., 4096-ray chunks, fp32, TF32 off), on this GPU
    ref_gpu = None
    try:
        from oracle import render_oracle as ro
        from oracle import render_oracle_torch as rt

        torch.backends.cuda.matmul.allow_tf32 = False
        gargs = (d["ray_batch"], d["background_prior"], d["inv_head_T"], d["planes"], wvol, wts, ro.default_boxes(), S, 0)
        for _ in range(2):
            rt.render_rays(*gargs, chunk=CPU_CHUNK_RAYS, device=dev, to_numpy=False)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            gout = rt.render_rays(*gargs, chunk=CPU_CHUNK_RAYS, device=dev, to_numpy=False)
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / 3
        ref_gpu = {"value": R / (gms * 1e-3), "unit": "rays/s", "ms_per_frame": gms, "kind": "port",
                   "what": "torch-CUDA port of the reference render path (same ATen ops, 4096-ray chunks, fp32, TF32 off), full 512x512x64 frame",
                   "max_abs_diff_vs_ours": float((gout["rgb_coarse"] - out.rgb_coarse).abs().max())}
        del gout
    except Exception as exc:
        ref_gpu = {"error": "%s: %s" % (type(exc).__name__, exc)}
    line = {
        "metric": METRIC, "value": world * R / (step_ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "stage-one NeRF render, 512x512 frame x 64 samples/ray, coarse only (BASELINE.json configs[1])",
                   "frames_per_gpu_per_step": 1, "rays_per_step": world * R, "samples_per_ray": S,
                   "mlp_arithmetic": "%s operands, fp32 accumulate (tcgen05/TMEM)" % args.precision if args.precision != "fp32" else "fp32 CUDA cores",
                   "l2": "256 MiB buffer written between timed steps (L2 flush); per-step CUDA events on the launch stream",
                   "acc_mean": acc_mean, "fidelity": psnr},
        "e2e": {"value": world * R / (e2e_ms * 1e-3), "unit": "rays/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "serial_ms_per_step": e2e_serial_ms,
                "maps": "image: rgb[..., :3], depth, acc read back per frame (train_avatar.py:182-218); feature channels stay on the device",
                "inputs": "per frame: 18-float camera block (rays generated inside the render kernel, dataloader/data_util.py:28-56), "
                          "background [R,3], inv_head_T, planes [2,1,64,128,128] -- all from pinned host memory",
                "with_ray_tensor_upload": {"value": world * R / (e2e_rays_ms * 1e-3), "ms_per_step": e2e_rays_ms, "h2d_bytes_per_step": h2d_rays,
                                           "note": "same call with the dataloader-built [R,8] ray tensor uploaded per frame (round-1 form)"},
                "all_maps": {"value": world * R / (e2e_all_ms * 1e-3), "ms_per_step": e2e_all_ms, "d2h_bytes_per_step": d2h_all,
                             "note": "same call downloading every map (67-channel colour + feature map, depth, acc, weights_max): PCIe / host-memory bound"},
                "api": "havatar_b200.render.PipelinedHostRenderer (pinned host in/out; H2D, hav_render_forward and D2H of consecutive frames overlap)"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_burst"], "traffic": traffic,
                     "kernel": "tc2::render_tc2_kernel", "kernel_ms": kern_ms, "flop_per_launch": R * S * FLOP_PER_SAMPLE,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone between L2 flushes), of %s" % peaks["source"],
                     "frac_of_sustained": achieved / peaks["bf16_sustained"],
                     "hbm_gbs_algorithmic": R * BYTES_PER_RAY / (kern_ms * 1e-3) / 1e9},
        "cpu_baseline": {"value": cpu_rate, "unit": "rays/s", "cores": threads, "kind": cpu_kind, "sample": cpu_sample,
                         "seconds": cpu_s},
        "reference_precision_mode": x3,
        "reference_gpu": ref_real,
        "reference_gpu_port": ref_gpu,
        "clocks": clocks,
        "train_note": dict(note="one optimiser iteration per step (eager = ~1500 host launches; graph = the iteration replayed as CUDA graphs), synthetic data, LPIPS omitted (weights unavailable offline): stage one = "
                                  "train_avatar.py:112-158 on 4 frames x 64x64-ray patches per GPU (64+16 samples, fused render fwd+bwd, patch "
                                  "discriminator); stage two = train_avatarHD.py:201-303 on 1 frame per GPU (D step + G step, 128^2 render -> 512^2)"),
        "hd": dict(hd, note="HD frames/s = XY/YZ plane generators (StyleGAN_zxc) + 512x512x64 or 128x128x64 render + SWGAN_unet, "
                            "one frame per GPU, CUDA-graph replay, random-init weights, per-rank values (not max-reduced)"),
    }
    note = line.pop("train_note")["note"]
    train = run_train_legs(args, torch, dist, dev, rank, world, barrier, line)
    line["train"] = dict(train, note=note)
    emit(line)
    if dist is not None:
        if "error" in train:      # a rank that failed inside a training leg leaves its peers in a collective: do not wait for them
            os._exit(0)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--train-mode", default="both", choices=["both", "eager", "graph"],
                    help="training-step leg: eager launches, whole-iteration CUDA graphs (train_step.Graphed), or both")
    ap.add_argument("--train-deadline", type=int, default=300, help="seconds after which the training legs are abandoned")
    ap.add_argument("--no-train", dest="no_train", action="store_true", help="skip the training-step measurement")
    ap.add_argument("--no-hd", dest="no_hd", action="store_true", help="skip the secondary HD frames/s measurement")
    ap.add_argument("--no-reference-gpu", dest="no_reference_gpu", action="store_true",
                    help="skip the reference legs that run in child processes (profiling runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
