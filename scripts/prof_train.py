"""Where a training step spends its GPU time: per-kernel totals (torch.profiler, CUDA activities) of StageOneStep / StageTwoStep."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from havatar_b200 import train_step  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if stage == 1:
    st, batch = train_step.StageOneStep(n_frames=4, capturable=True), train_step.synthetic_batch(1, 4, "cuda", patch=64)
else:
    st, batch = train_step.StageTwoStep(n_frames=1, capturable=True), train_step.synthetic_batch(2, 1, "cuda", render_size=128, gen_size=512)
for _ in range(3):
    st(batch)
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        st(batch)
    torch.cuda.synchronize()
tot, cnt = collections.Counter(), collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name] += e.device_time
        cnt[e.name] += 1
total = sum(tot.values())
print("stage %d: %.2f ms of kernels per step, %d launches per step" % (stage, total / N / 1e3, sum(cnt.values()) // N))
for name, t in tot.most_common(40):
    print("%9.1f us %5.1f%% %5d  %s" % (t / N, 100.0 * t / total, cnt[name] // N, name[:150]))
