"""Graph-replayed training iterations under torchrun (one rank per GPU): ms per iteration, max over ranks.
torchrun --nproc-per-node N scripts/time_train_dist.py [stage ...]      knobs: HAV_GRAD_BUCKET_MB, NCCL_* environment"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from havatar_b200 import train_step  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    if os.environ.get("HAV_NCCL_HIGH_PRIO"):
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    else:
        dist.init_process_group("nccl", device_id=dev)
stages = [int(a) for a in sys.argv[1:]] or [1, 2]
for stage in stages:
    if stage == 1:
        st = train_step.StageOneStep(n_frames=4 * world, device=dev, capturable=True)
        batch = train_step.synthetic_batch(1, 4, dev, seed=rank, patch=64, frame_offset=4 * rank)
        n_warm, n_t = 3, 10
    else:
        st = train_step.StageTwoStep(n_frames=world, device=dev, capturable=True)
        batch = train_step.synthetic_batch(2, 1, dev, seed=rank, render_size=128, gen_size=512, frame_offset=rank)
        n_warm, n_t = 16, 16
    run = train_step.Graphed(st, batch)
    for _ in range(n_warm):
        run(batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_t + 1)]
    a.record()
    marks[0].record()
    for i in range(n_t):
        run(batch)
        marks[i + 1].record()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / n_t], dtype=torch.float64, device=dev)
    if rank == 0 and os.environ.get("HAV_PER_ITER"):
        print("  per iteration (rank 0):", " ".join("%.1f" % marks[i].elapsed_time(marks[i + 1]) for i in range(n_t)), flush=True)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    syncs = [g.sync for g in st.groups() if g.sync is not None]
    if rank == 0:
        print("  buckets issued from inside backward since construction: %s of %s collectives" % (
            [s_.issued_in_backward for s_ in syncs], [s_.collectives for s_ in syncs]), flush=True)
        print("stage %d, %d GPU(s): %.2f ms per iteration; %d buckets, %.0f MB exchanged (bucket %s MB, NCCL_MAX_NCHANNELS=%s, NCCL_ALGO=%s)" % (
            stage, world, float(ms), sum(len(s.buckets) for s in syncs), sum(s.bytes_per_step for s in syncs) / 1e6,
            os.environ.get("HAV_GRAD_BUCKET_MB", "64"), os.environ.get("NCCL_MAX_NCHANNELS", "-"), os.environ.get("NCCL_ALGO", "-")), flush=True)
    del run, st, batch
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
