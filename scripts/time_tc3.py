"""rays/s of the render kernels on the 512 x 512 x 64 frame: single-CTA 16-bit (v2), CTA-pair 16-bit (v3), split precision, with
the kernel-only timing of bench.py (weights / planes already packed, L2 flushed between launches).  GPU box only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from havatar_b200 import render, synth  # noqa: E402

sc = synth.scene(batch=1, height=512, width=512, seed=0)
d = {k: torch.from_numpy(np.ascontiguousarray(sc[k])).cuda() for k in ("ray_batch", "background_prior", "inv_head_T", "planes", "wvol")}
w = {k: torch.from_numpy(v).cuda() for k, v in sc["weights"].items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for name, kw in (("v2 fp16", dict(precision="fp16")), ("v3 fp16 (CTA pairs)", dict(precision="fp16", cta_pairs=True)),
                 ("v3 bf16 (CTA pairs)", dict(precision="bf16", cta_pairs=True)), ("v3 fp16x3 (split)", dict(precision="fp16x3"))):
    out = None
    for _ in range(3):
        out = render.render_rays(d["ray_batch"], d["background_prior"], d["inv_head_T"], d["planes"], d["wvol"], w, 64, 0, out=out, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = render.render_rays(d["ray_batch"], d["background_prior"], d["inv_head_T"], d["planes"], d["wvol"], w, 64, 0, out=out,
                                 reuse_packed=True, **kw)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    line = "%-22s %.3f ms  %.1f M rays/s  %.0f TFLOP/s (1x count)" % (name, ms, 262144 / ms / 1e3, 262144 * 64 * 94848 / ms / 1e9)
    if ref is None:
        ref = out.rgb_coarse.clone()
    else:
        line += "  max|d| vs v2 %.2e" % float((out.rgb_coarse - ref).abs().max())
    print(line, flush=True)
