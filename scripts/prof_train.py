"""Where a training step spends its GPU time: torch.profiler kernel table of StageOneStep / StageTwoStep."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from havatar_b200 import train_step  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if stage == 1:
    st, batch = train_step.StageOneStep(n_frames=4), train_step.synthetic_batch(1, 4, "cuda", patch=64)
else:
    st, batch = train_step.StageTwoStep(n_frames=1), train_step.synthetic_batch(2, 1, "cuda", render_size=128, gen_size=512)
for _ in range(3):
    st(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        st(batch)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
