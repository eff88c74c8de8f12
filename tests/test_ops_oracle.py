"""Pins oracle/ops_oracle.py against the reference's own CPU fallbacks of `model/op`
(tests/golden/ops.npz, minted by oracle/gen_golden.py::gen_ops).  CPU only."""
import os

import numpy as np

from oracle import ops_oracle as oo
from oracle.gen_golden import UFD_CASES, ufd_kernel


def test_upfirdn2d_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ops.npz"))
    for name, shape, taps, gain, up, down, pad in UFD_CASES:
        x, ref = g["ufd_%s_x" % name], g["ufd_%s_y" % name]
        got = oo.upfirdn2d(x, ufd_kernel(np, taps, gain), up[0], up[1], down[0], down[1], *pad)
        assert got.shape == ref.shape, name
        assert np.abs(got - ref).max() < 2e-6, name


def test_fused_act_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ops.npz"))
    assert np.array_equal(oo.fused_leaky_relu(g["act_x"], g["act_b"]), g["act_y"])
    assert np.array_equal(oo.fused_leaky_relu(g["act_x"]), g["act_y_nobias"])
    assert np.array_equal(oo.fused_leaky_relu(g["act2_x"], g["act2_b"]), g["act2_y"])
