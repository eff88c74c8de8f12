"""hav_adam_flat (parallel.FlatAdam): one-pass Adam over flat parameter / gradient / moment buffers against torch.optim.Adam
(the optimiser of train_avatar.py:68-71 and train_avatarHD.py:117-122) on the same gradients."""
import pytest
import torch
from torch import nn

from havatar_b200 import parallel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("betas", [(0.9, 0.999), (0.0, 0.99 ** 0.8)])
def test_flat_adam_matches_torch_adam(betas):
    torch.manual_seed(0)
    shapes = [(64, 33, 3, 3), (7,), (128, 100), (1,), (5, 5, 5)]        # ragged sizes: every slot is padded to 32 elements
    ref = [nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ours = [nn.Parameter(p.detach().clone()) for p in ref]
    opt_ref = torch.optim.Adam(ref, lr=2e-3, betas=betas)
    opt = parallel.FlatAdam(ours, lr=2e-3, betas=betas)
    for step in range(5):
        if step == 3:
            opt.set_lr(5e-4)
            for g in opt_ref.param_groups:
                g["lr"] = 5e-4
        for a, b in zip(ref, ours):
            g = torch.randn_like(a) * (10.0 ** (step - 3))
            a.grad = g.clone()
            b.grad.add_(g)                     # autograd accumulates into the flat view in place
        opt_ref.step()
        opt.step()
        assert float(opt.layout.flat_g.abs().max()) == 0.0          # gradients cleared by the same pass
    torch.cuda.synchronize()
    opt.layout.check()
    for a, b in zip(ref, ours):
        assert b.data_ptr() >= opt.layout.flat_p.data_ptr()
        assert float((a - b).detach().abs().max()) <= 2e-6 * max(1.0, float(a.detach().abs().max()))
    assert float(opt.state[0]) == 5.0


def test_flat_adam_is_graph_capturable():
    torch.manual_seed(1)
    p = [nn.Parameter(torch.randn(1000, device="cuda"))]
    q = [nn.Parameter(p[0].detach().clone())]
    a, b = parallel.FlatAdam(p, lr=1e-2), parallel.FlatAdam(q, lr=1e-2)
    grad = torch.randn(1000, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            a.layout.flat_g[:1000].add_(grad)
            a.step()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
        b.layout.flat_g[:1000].add_(grad)
        b.step()
    torch.cuda.synchronize()
    assert float(a.state[0]) == 3.0
    assert torch.equal(p[0].detach(), q[0].detach())
