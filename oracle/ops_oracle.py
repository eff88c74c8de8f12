"""CPU oracle for the reference's `model/op` operators (numpy, fp32).

TEST INFRASTRUCTURE ONLY (see oracle/render_oracle.py header).  Pinned against the reference's own
CPU fallbacks -- its executable specification of the two CUDA ops -- through
tests/golden/ops.npz minted by oracle/gen_golden.py.
"""
import numpy as np

F32 = np.float32


def fused_bias_act(x, bias, ref, act, grad, alpha, scale):
    """model/op/fused_bias_act_kernel.cu:18-65: y = f(x + b[(i/step_b) % size_b]) * scale with
    f selected by act*10+grad (30 lrelu, 31 lrelu-grad gated by sign(ref), 32/12 zero, 1x linear).
    x: [N,C,...]; bias: [C] or None; ref: like x or None."""
    x = np.asarray(x, dtype=F32)
    v = x
    if bias is not None:
        shape = [1, -1] + [1] * (x.ndim - 2)
        v = (x + np.asarray(bias, dtype=F32).reshape(shape)).astype(F32)
    mode = act * 10 + grad
    if mode == 30:
        y = np.where(v > 0, v, v * F32(alpha))
    elif mode == 31:
        y = np.where(np.asarray(ref) > 0, v, v * F32(alpha))
    elif mode in (12, 32):
        y = np.zeros_like(v)
    else:
        y = v
    return (y.astype(F32) * F32(scale)).astype(F32)


def fused_leaky_relu(x, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """model/op/fused_act.py:107-119 (CPU branch; it hard-codes slope 0.2 == the shipped setting)."""
    return fused_bias_act(x, bias, None, 3, 0, negative_slope, scale)


def upfirdn2d(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """model/op/upfirdn2d.py:172-213 upfirdn2d_native on NCHW input: zero-insert upsample, pad
    (negative pad crops), correlate with the flipped kernel, decimate."""
    x = np.asarray(x, dtype=F32)
    k = np.asarray(kernel, dtype=F32)
    n, c, in_h, in_w = x.shape
    kh, kw = k.shape
    up = np.zeros((n, c, in_h * up_y, in_w * up_x), dtype=F32)
    up[:, :, ::up_y, ::up_x] = x
    up = np.pad(up, ((0, 0), (0, 0), (max(pad_y0, 0), max(pad_y1, 0)), (max(pad_x0, 0), max(pad_x1, 0))))
    up = up[:, :, max(-pad_y0, 0): up.shape[2] - max(-pad_y1, 0), max(-pad_x0, 0): up.shape[3] - max(-pad_x1, 0)]
    fh, fw = up.shape[2] - kh + 1, up.shape[3] - kw + 1
    kf = k[::-1, ::-1]
    out = np.zeros((n, c, fh, fw), dtype=F32)
    for ky in range(kh):
        for kx in range(kw):
            out += up[:, :, ky:ky + fh, kx:kx + fw] * kf[ky, kx]
    return np.ascontiguousarray(out[:, :, ::down_y, ::down_x])


def make_kernel(k):
    """model/styleUnet.py:17-26 make_kernel: outer product of a 1-D tap list, normalised to sum 1."""
    k = np.asarray(k, dtype=F32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return (k / k.sum()).astype(F32)
