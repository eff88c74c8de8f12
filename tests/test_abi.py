"""The C-ABI library loads on a CPU-only host and exports every symbol include/*.h declares.
No compute calls here (no GPU); argument-validation paths that return before any CUDA call are exercised."""
import ctypes as C
import glob
import os
import re

import pytest

from havatar_b200 import _lib
from havatar_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _lib.lib()


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(hav_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported_and_bound(lib):
    decl = declared_symbols()
    assert len(decl) >= 7
    for name in decl:
        assert hasattr(lib, name), "libhavatar_b200.so does not export " + name
    assert decl == set(_lib.SIGNATURES), "ctypes binding table and header disagree"


def test_abi_version_and_error_strings(lib):
    assert lib.hav_abi_version() == _lib.HAV_ABI_VERSION
    assert lib.hav_error_string(0) == b"ok"
    assert b"NULL" in lib.hav_error_string(-1)


def test_struct_layout_guard(lib):
    a = _lib.RenderArgs()
    a.struct_bytes = C.sizeof(_lib.RenderArgs) - 8      # wrong size -> HAV_E_VALUE, workspace query says 0
    assert lib.hav_render_workspace_bytes(C.byref(a)) == 0
    assert lib.hav_render_forward(C.byref(a), None) == -5
    assert lib.hav_render_forward(None, None) == -1


def test_argument_errors_return_before_any_launch(lib):
    a = _lib.RenderArgs()
    a.struct_bytes = C.sizeof(_lib.RenderArgs)
    a.precision, a.batch, a.rays, a.num_coarse, a.num_fine = 0, 1, 16, 1, 0   # num_coarse < 2
    a.plane_c, a.plane_h, a.plane_w, a.vol_d, a.vol_h, a.vol_w = 64, 8, 8, 4, 4, 4
    assert lib.hav_render_forward(C.byref(a), None) == -2
    a.num_coarse = 8
    assert lib.hav_render_forward(C.byref(a), None) == -1                     # NULL tensors
    a.rays = 0
    assert lib.hav_render_forward(C.byref(a), None) == 0                      # empty ray batch is a no-op
    assert lib.hav_fused_bias_act(None, None, None, None, 0, 1, 1, 3, 0, 0.2, 1.0, None) == 0
    assert lib.hav_fused_bias_act(None, None, None, None, 4, 1, 1, 3, 0, 0.2, 1.0, None) == -1
    assert lib.hav_upfirdn2d(None, None, None, 1, 4, 4, 1, 30, 30, 1, 1, 1, 1, 0, 0, 0, 0, None) == -2


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhavatar_b200.so")
    with pytest.raises(_lib.HavError):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    import torch

    from havatar_b200 import op

    with pytest.raises(RuntimeError):
        op.fused_leaky_relu(torch.zeros(1, 2, 3, 3))
    with pytest.raises(RuntimeError):
        op.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))


def test_training_entry_points_validate_before_any_launch(lib):
    """hav_conv2d_wgrad / hav_rowscale_dot / hav_adam_flat: argument errors are reported without touching the device."""
    assert lib.hav_conv2d_wgrad(None, None) == -1
    a = _lib.ConvWgradArgs()
    a.struct_bytes = C.sizeof(_lib.ConvWgradArgs) - 4
    assert lib.hav_conv2d_wgrad(C.byref(a), None) == -5
    a.struct_bytes = C.sizeof(_lib.ConvWgradArgs)
    a.batch, a.cin, a.cout, a.in_h, a.in_w, a.ksize, a.up, a.down = 1, 8, 8, 0, 4, 3, 1, 1
    assert lib.hav_conv2d_wgrad(C.byref(a), None) == -2                       # empty image
    a.in_h, a.ksize = 4, 5
    assert lib.hav_conv2d_wgrad(C.byref(a), None) == -5                       # only 1x1 and 3x3
    a.ksize, a.up, a.down = 3, 2, 2
    assert lib.hav_conv2d_wgrad(C.byref(a), None) == -5                       # up and down together
    a.up, a.down = 1, 1
    assert lib.hav_conv2d_wgrad(C.byref(a), None) == -1                       # dw is NULL
    assert lib.hav_rowscale_dot(None, None, None, None, None, 4, 4, None) == -1
    assert lib.hav_adam_flat(None, None, None, None, 8, None, 0.9, 0.999, 1e-8, 1.0, 1, None) == -1
    buf = (C.c_float * 16)()
    p = C.cast(buf, C.c_void_p)
    assert lib.hav_adam_flat(p, p, p, p, 6, p, 0.9, 0.999, 1e-8, 1.0, 1, None) == -2     # n must be a multiple of 4


def test_flat_layout_views_on_cpu():
    """parallel.FlatLayout is pure tensor plumbing: parameters and gradients become views of two flat buffers."""
    import torch
    from torch import nn

    from havatar_b200 import parallel

    net = nn.Sequential(nn.Linear(5, 7), nn.Linear(7, 3))
    want = [p.detach().clone() for p in net.parameters()]
    lay = parallel.FlatLayout(net.parameters())
    assert lay.numel % 32 == 0 and all(o % 32 == 0 for o in lay.offsets)
    for p, w in zip(net.parameters(), want):
        assert torch.equal(p.detach(), w) and p.grad is not None and float(p.grad.abs().sum()) == 0.0
    net(torch.randn(4, 5)).sum().backward()
    lay.check()
    assert float(lay.flat_g.abs().sum()) > 0.0
    lay.flat_p.zero_()
    assert all(float(p.detach().abs().sum()) == 0.0 for p in net.parameters())


def test_bias_act_backward_splits_is_a_host_function(lib):
    """hav_bias_act_backward_splits needs no GPU: 1..64 partial sums per channel, about two CTAs per SM over all channels, never
    more than one per 2048 elements."""
    assert lib.hav_bias_act_backward_splits(0, 4, 16) == 0
    assert lib.hav_bias_act_backward_splits(4, 512, 16 * 16) == 1
    assert lib.hav_bias_act_backward_splits(1, 64, 512 * 512) == 5
    assert lib.hav_bias_act_backward_splits(1, 3, 1024 * 1024) == 64
    assert lib.hav_bias_act_backward_splits(2, 12, 100) == 1
    assert lib.hav_style_plan_run(None, None, None, 1, 1, 1, None, None, None, 1, 1, 0, 1e-8, None) == -1     # NULL tensors
    assert lib.hav_bias_act_backward(None, None, None, None, 1, 4, 16, 65, 0.2, 1.0, None) == -2               # splits > 64
