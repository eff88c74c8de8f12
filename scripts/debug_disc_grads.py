"""Per-parameter gradient error of the Discriminator's first- and second-order (R1) backward: our tcgen05 convolution autograd
against torch's library convolutions (fp32, TF32 off) inside the same network formulation, same weights and inputs, and both
against the reference golden (tests/golden/stage_two_step.npz holds the reference's R1 gradients).  GPU box only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from havatar_b200 import styleunet, styleunet_train, train_step  # noqa: E402
from oracle.gen_golden import STAGE_TWO_CASE, stage_two_inputs, stage_two_states, subsample  # noqa: E402

torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
c = STAGE_TWO_CASE
disc = styleunet.Discriminator(c["gen_size"], img_channel=3).cuda()
shp = {k: tuple(v.shape) for k, v in disc.state_dict().items()}
from havatar_b200 import synth  # noqa: E402
sd = synth.styleunet_state(shp, c["seed"] + 2)
disc.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
inp = stage_two_inputs(c)
x = torch.from_numpy(inp["gt_hr"]).cuda()
names = [n for n, _ in disc.named_parameters()]


def grads(lib, r1):
    styleunet.invalidate_caches()
    disc.zero_grad(set_to_none=True)
    ctx = styleunet_train.library_convs() if lib else __import__("contextlib").nullcontext()
    with ctx:
        if r1:
            xi = x.clone().requires_grad_(True)
            pred = disc(xi)
            loss = train_step.d_r1_loss(pred, xi)
        else:
            loss = torch.nn.functional.softplus(-disc(x)).mean()
        loss.backward()
    return float(loss), {n: p.grad.detach().clone() for n, p in disc.named_parameters() if p.grad is not None}


for r1 in (False, True):
    lo, go = grads(False, r1)
    ll, gl = grads(True, r1)
    print("==== %s: loss ours %.6g library %.6g" % ("R1" if r1 else "first order", lo, ll))
    rows = []
    for n in names:
        if n in go and n in gl:
            a, b = go[n], gl[n]
            rows.append((float((a - b).abs().max() / (b.abs().max() + 1e-30)), n, tuple(a.shape), float(b.abs().max())))
    for e, n, s, m in sorted(rows, reverse=True)[:12]:
        print("  %.4f  %-34s %-22s max|g| %.3g" % (e, n, s, m))
    if r1:
        g = np.load(os.path.join(ROOT, "tests", "golden", "stage_two_step.npz"))
        print("  (reference golden is taken after the D step's Adam update, so only indicative here)")
    # where does the error of final_conv sit?  per input channel
    for key in ("final_conv.0.weight", "final_linear.0.weight"):
        if key in go:
            a, b = go[key], gl[key]
            d = (a - b).abs()
            if a.dim() == 4:
                per_ci = d.amax(dim=(0, 2, 3))
                top = torch.topk(per_ci, 5)
                print("  %s: worst input channels %s  errs %s  (cin = %d)" % (key, top.indices.tolist(), [float(v) for v in top.values], a.shape[1]))
            else:
                per_o = d.amax(dim=1)
                top = torch.topk(per_o, 5)
                print("  %s: worst rows %s errs %s" % (key, top.indices.tolist(), [float(v) for v in top.values]))
