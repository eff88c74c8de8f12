"""World-size-2 gloo run on CPU of the data-parallel gradient exchange (havatar_b200/parallel.py): frames
sharded over ranks, gradients all-reduced in buckets from inside backward.  The averaged gradients -- and the weights after two
Adam steps -- must equal a single process stepping on the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from havatar_b200 import parallel, shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _net(seed):
    torch.manual_seed(seed)
    # `unused` never receives a gradient: its bucket slot must still be exchanged (as zeros) without deadlocking
    net = nn.Sequential(nn.Linear(12, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 5))
    net.unused = nn.Parameter(torch.ones(7))
    net.codes = nn.Parameter(torch.randn(6, 12) * 0.1)          # row-sparse gradient, like Trainer.latent_codes
    return net


def _data():
    g = torch.Generator().manual_seed(11)
    return torch.randn(6, 9, 12, generator=g), torch.randn(6, 9, 5, generator=g)      # 6 "frames" x 9 rays


def _loss(net, x, y, fidx):
    # per-frame mean: ranks hold equal frame counts, so the mean over ranks of per-rank means is the global mean
    return ((net(x + net.codes[fidx][:, None, :]) - y) ** 2).mean()


def _worker(rank, world, port, bucket_bytes, use_layout, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _net(seed=100 + rank)                             # different init per rank: the broadcast must fix that
        parallel.broadcast_parameters(net)
        layout = parallel.FlatLayout(net.parameters()) if use_layout else None      # buckets = slices of one flat gradient buffer
        sync = parallel.GradSync(net.parameters(), bucket_bytes=bucket_bytes, layout=layout)
        opt = torch.optim.Adam(net.parameters(), lr=1e-2)
        x, y = _data()
        lo, hi = shard.frames_of_rank(6, world, rank)
        grads = None
        for step in range(2):
            _loss(net, x[lo:hi], y[lo:hi], torch.arange(lo, hi)).backward()
            issued_in_backward = sync.collectives
            sync.finish()
            if step == 0:
                grads = {k: p.grad.clone() for k, p in net.named_parameters()}
            opt.step()
            sync.zero_grad()
        if layout is not None:
            layout.check()
            assert all(p.data_ptr() >= layout.flat_p.data_ptr() for p in net.parameters())
        if rank == 0:
            ret["grads"] = {k: v.numpy() for k, v in grads.items()}
            ret["weights"] = {k: p.detach().numpy() for k, p in net.named_parameters()}
            ret["buckets"] = len(sync.buckets)
            ret["collectives"] = sync.collectives
            ret["issued_in_backward"] = issued_in_backward
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes,use_layout", [(64 << 20, False), (4096, False), (4096, True)])
def test_bucketed_allreduce_equals_full_batch(bucket_bytes, use_layout):
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, bucket_bytes, use_layout, ret), nprocs=world, join=True)
        got_g, got_w = dict(ret["grads"]), dict(ret["weights"])
        nb, nc, early = ret["buckets"], ret["collectives"], ret["issued_in_backward"]
    net = _net(seed=100)                                        # rank 0's init is what the broadcast spreads
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    x, y = _data()
    for step in range(2):
        _loss(net, x, y, torch.arange(6)).backward()
        if step == 0:
            ref_g = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in net.named_parameters()}
        opt.step()
        opt.zero_grad()
    for k, v in ref_g.items():
        assert abs(got_g[k] - v.numpy()).max() < 1e-6, k
    for k, p in net.named_parameters():
        assert abs(got_w[k] - p.detach().numpy()).max() < 1e-5, k
    assert nc == 2 * nb                                         # one collective per bucket per step, nothing else
    if bucket_bytes == 4096:
        assert nb >= 3 and early >= nb + 1                      # several buckets, and some went out from inside backward


def test_gradsync_single_process_is_a_noop_wrapper():
    net = _net(0)
    sync = parallel.GradSync(net.parameters(), bucket_bytes=4096)
    x, y = _data()
    _loss(net, x, y, torch.arange(6)).backward()
    sync.finish()
    ref = _net(0)
    _loss(ref, x, y, torch.arange(6)).backward()
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert torch.equal(p.grad, q.grad if q.grad is not None else torch.zeros_like(q)), k
    torch.optim.Adam(net.parameters()).zero_grad(set_to_none=True)       # breaks the bucket views ...
    _loss(net, x, y, torch.arange(6)).backward()
    with pytest.raises(RuntimeError, match="aliases"):                   # ... and finish() says so instead of exchanging stale data
        sync.finish()
