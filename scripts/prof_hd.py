"""Where an HD inference frame spends its GPU time: per-kernel totals (torch.profiler, CUDA activities) of AvatarHD.frame, eager.
python scripts/prof_hd.py [render_size out_size]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from havatar_b200 import pipeline, synth  # noqa: E402

rs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
out = int(sys.argv[2]) if len(sys.argv) > 2 else 512
torch.manual_seed(0)
sc = synth.scene(batch=1, height=rs, width=rs, seed=0)
net = pipeline.AvatarHD(sc["weights"], sc["wvol"], render_size=rs, out_size=out).cuda()
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
args = (dev(sc["ray_batch"]), dev(sc["background_prior"]), torch.zeros(1, 32, device="cuda"), dev(sc["inv_head_T"]),
        torch.rand(1, 7, 256, 256, device="cuda"), torch.rand(1, 7, 256, 256, device="cuda"), torch.rand(1, 7, 256, 256, device="cuda"),
        torch.randn(1, 64, device="cuda"))
for _ in range(3):
    net.frame(*args)
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        net.frame(*args)
    torch.cuda.synchronize()
tot, cnt = collections.Counter(), collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name] += e.device_time
        cnt[e.name] += 1
total = sum(tot.values())
print("HD %d -> %d: %.3f ms of kernels per frame, %d launches per frame" % (rs, out, total / N / 1e3, sum(cnt.values()) // N))
for name, t in tot.most_common(40):
    print("%9.1f us %5.1f%% %5d  %s" % (t / N, 100.0 * t / total, cnt[name] // N, name[:150]))
