"""The UNMODIFIED reference code running on the B200 drop-ins (INTEGRATION.md sections 0 and 1).

baseline/_ref holds a byte copy of the reference's Python packages (baseline/install_ref.py; git-ignored, it travels to the GPU
box with the snapshot).  Two wirings are exercised:
  1. `havatar_b200.op.install_reference_modules()`: the reference's OWN model/styleUnet.py + model/op/*.py wrappers, with the two
     bare-name extension modules they import (`fused`, `upfirdn2d`) served by our kernels -> its SWGAN_unet must reproduce the
     reference golden;
  2. `havatar_b200.compat.install()`: the reference's OWN utils/styleUnet_util.py (losses, R1 penalty, EMA) driving OUR
     Discriminator / SWGAN_unet through the `model.*` names the entry scripts import.
Skipped when baseline/_ref is not installed."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "havatar")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")), reason="baseline/_ref not installed")]


def _purge():
    for name in [n for n in sys.modules if n == "model" or n.startswith("model.") or n == "utils" or n.startswith("utils.")
                 or n in ("fused", "upfirdn2d")]:
        del sys.modules[name]
    if REF in sys.path:
        sys.path.remove(REF)


@pytest.fixture
def clean_imports():
    _purge()
    yield
    from havatar_b200 import compat

    compat.uninstall()
    _purge()


def test_reference_styleunet_runs_on_our_op_kernels(golden_dir, clean_imports):
    import warnings

    from havatar_b200 import op
    from oracle.gen_golden import STYLEUNET_CASES, styleunet_inputs

    op.install_reference_modules()                         # sys.modules["fused"], ["upfirdn2d"] -> sm_100a kernels
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore", message="conv2d_gradfix not supported")
    ref = importlib.import_module("model.styleUnet")       # the reference's file, unmodified
    assert os.path.realpath(ref.__file__).startswith(os.path.realpath(REF))
    case = STYLEUNET_CASES["swgan_32_128"]
    net = ref.SWGAN_unet(**case["kw"])
    sd, style, cond, noise = styleunet_inputs(case, net)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    net = net.cuda()
    t = lambda a: torch.from_numpy(a).cuda()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            out = net([t(style)], t(cond), noise=[t(n) for n in noise])
    finally:
        torch.backends.cudnn.allow_tf32 = old
    g = np.load(os.path.join(golden_dir, "styleunet_swgan_32_128.npz"))["out"]
    # fp32 everywhere (cuDNN convolutions + our fp32 upfirdn2d / fused_bias_act kernels): 1e-4 of the output range
    assert np.abs(out.cpu().numpy() - g).max() < 1e-4 * np.abs(g).max()


def test_reference_training_utilities_drive_our_modules(clean_imports):
    from havatar_b200 import compat, train_step

    compat.install()
    sys.path.insert(0, REF)
    su = importlib.import_module("utils.styleUnet_util")   # the reference's file, unmodified; imports `model.op.conv2d_gradfix`
    assert os.path.realpath(su.__file__).startswith(os.path.realpath(REF))
    from model.styleUnet import SWGAN_unet, Discriminator   # the names train_avatarHD.py:20,27 import -> our drop-ins

    assert Discriminator.__module__.startswith("havatar_b200")
    torch.manual_seed(0)
    disc = Discriminator(64, 3).cuda()
    real = torch.rand(2, 3, 64, 64, device="cuda") * 2 - 1
    fake = torch.rand(2, 3, 64, 64, device="cuda") * 2 - 1
    su.requires_grad(disc, True)
    d_loss = su.d_logistic_loss(disc(real), disc(fake))
    assert torch.isfinite(d_loss) and abs(float(d_loss) - float(train_step.d_logistic_loss(disc(real), disc(fake)))) < 1e-5
    xr = real.clone().requires_grad_(True)
    pred = disc(xr)
    r1 = su.d_r1_loss(pred, xr)                            # utils/styleUnet_util.py:72-79: double backward through our kernels
    xr2 = real.clone().requires_grad_(True)
    r1_ours = train_step.d_r1_loss(disc(xr2), xr2)
    assert torch.isfinite(r1) and abs(float(r1) - float(r1_ours)) < 1e-3 * abs(float(r1_ours)) + 1e-9
    (10.0 / 2 * r1 * 16 + 0 * pred[0]).sum().backward()   # train_avatarHD.py:239: the 0 * pred term reaches the biases
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in disc.parameters())
    g = SWGAN_unet(inp_size=32, inp_ch=8, out_ch=3, out_size=128, style_dim=64, n_mlp=4).cuda()
    g_ema = SWGAN_unet(inp_size=32, inp_ch=8, out_ch=3, out_size=128, style_dim=64, n_mlp=4).cuda()
    su.accumulate(g_ema, g, 0)                             # train_avatarHD.py:115
    assert all(torch.equal(a, b) for a, b in zip(g_ema.parameters(), g.parameters()))
    noise = su.mixing_noise(2, 64, 0.0, "cuda")
    img = g(noise, torch.randn(2, 8, 32, 32, device="cuda"))
    assert img.shape == (2, 3, 128, 128)
    disc128 = Discriminator(128, 3).cuda()
    assert torch.isfinite(su.g_nonsaturating_loss(disc128(img)))
