"""Host side of the tensor-core convolution (C ABI: hav_conv2d_forward, hav_conv_pack_weights, hav_modconv_demod,
hav_conv2d_wgrad, hav_rowscale_dot).

conv2d() is ModulatedConv2d.forward / EqualConv2d.forward of the reference (model/styleUnet.py:222-297, :108-118)
with the per-layer elementwise work of StyledConv / ToRGB / ConvLayer fused into the same launch (the inference call).
conv2d_autograd() is the differentiable primitive of the training steps: the same forward kernel, and a backward made of the
forward kernel again (data gradient), the tcgen05 weight-gradient kernel and one row kernel for the modulation gradients --
what the reference gets from cuDNN through model/op/conv2d_gradfix.py.  torch is used for device memory, streams and the
autograd graph.  No CPU fallback."""
import ctypes as C

import torch

from . import _lib


class PackedConvWeight:
    """16-bit weight image of one convolution in the kernel's shared-memory order (hav_conv_pack_weights)."""

    def __init__(self, data, cout, cin, ksize, precision, up=1):
        self.data, self.cout, self.cin, self.ksize, self.precision, self.up = data, cout, cin, ksize, precision, up


def _check(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
        raise _lib.HavError("%s must be a float32 CUDA tensor (havatar_b200 has no CPU path)" % name)
    return t.detach().contiguous()


# Bumped by every CUDA-graph capture this package starts (graph.GraphedForward, train_step.Graphed): a weight image packed inside
# one capture lives in that graph's memory pool and its event belongs to that capture, so it is reused only inside the same one.
CAPTURE_GEN = [0]


def pack_weights_cached(weight, scale=1.0, up=1, transpose_io=False, precision="fp16", flip=False, cache=None):
    """pack_weights memoised in `cache` (a dict owned by the module that owns `weight`) on (storage, version, cache epoch):
    within one training iteration the same parameter is packed once per (layout, precision) even when several passes use it
    (the D step and the G step both run the generator forward).  cache=None packs unconditionally."""
    if cache is None:
        return pack_weights(weight, scale, up=up, transpose_io=transpose_io, precision=precision, flip=flip)
    from . import styleunet          # the epoch that CUDA-graph replays bump (parameters change behind their version counters)

    key = (weight.data_ptr(), weight._version, styleunet._EPOCH[0], tuple(weight.shape), float(scale))
    slot = (int(up), bool(transpose_io), precision, bool(flip))
    hit = cache.get(slot)
    cur = torch.cuda.current_stream(weight.device)
    capturing = torch.cuda.is_current_stream_capturing()
    # an image packed inside a CUDA-graph capture lives in that graph's memory and may only be reused inside the same capture
    gen = CAPTURE_GEN[0] if capturing else None
    if hit is None or hit[0] != key or (hit[4] is not None and hit[4] != gen):
        packed = pack_weights(weight, scale, up=up, transpose_io=transpose_io, precision=precision, flip=flip)
        ev = torch.cuda.Event()
        ev.record(cur)
        hit = (key, packed, cur.cuda_stream, ev, gen)
        cache[slot] = hit
    elif hit[2] != cur.cuda_stream and (hit[4] is not None or not capturing):
        # packed on another stream (two passes of one network running side by side, pipeline.run_parallel): order this stream
        # after the packing kernel.  (An image packed eagerly and met again under capture needs no edge: every capture in this
        # package starts after a device synchronisation, and a capture may not depend on uncaptured work.)
        cur.wait_event(hit[3])
    return hit[1]


def pack_weights(weight, scale=1.0, up=1, transpose_io=False, precision="fp16", flip=False):
    """weight [Cout,Cin,k,k] (or [Cin,Cout,k,k] with transpose_io) -> PackedConvWeight holding scale * weight, laid out for
    conv2d(..., up=up) (the transposed convolution uses 64-channel tiles and four phase accumulators).  flip mirrors the taps
    (with transpose_io: the weight of a stride-1 layer's data-gradient convolution)."""
    L = _lib.lib()
    w = _check(weight, "weight")
    if w.dim() != 4 or w.shape[2] != w.shape[3]:
        raise _lib.HavError("weight must be [Cout,Cin,k,k]")
    cout, cin = (int(w.shape[1]), int(w.shape[0])) if transpose_io else (int(w.shape[0]), int(w.shape[1]))
    k = int(w.shape[2])
    nbytes = int(L.hav_conv_wpack_bytes(cout, cin, k, int(up)))
    if nbytes == 0:
        raise _lib.HavError("unsupported convolution shape (ksize must be 1 or 3)")
    buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    with torch.cuda.device(w.device):
        st = torch.cuda.current_stream(w.device).cuda_stream
        _lib.check(L.hav_conv_pack_weights(C.c_void_p(buf.data_ptr()), C.c_void_p(w.data_ptr()), cout, cin, k, float(scale),
                                           int(up), int(bool(transpose_io)) | (2 if flip else 0), _lib.PRECISIONS[precision], C.c_void_p(st)),
                   "hav_conv_pack_weights")
    return PackedConvWeight(buf, cout, cin, k, precision, int(up))


def modconv_demod(weight, style, scale, eps=1e-8):
    """rsqrt(sum((scale * weight * style)^2) + eps) -> [B,Cout]  (model/styleUnet.py:256-258)."""
    L = _lib.lib()
    w, s = _check(weight, "weight"), _check(style, "style")
    cout, cin, k = int(w.shape[0]), int(w.shape[1]), int(w.shape[2])
    if s.dim() != 2 or s.shape[1] != cin:
        raise _lib.HavError("style must be [B,Cin]")
    out = torch.empty((s.shape[0], cout), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        st = torch.cuda.current_stream(w.device).cuda_stream
        _lib.check(L.hav_modconv_demod(C.c_void_p(out.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(s.data_ptr()),
                                       int(s.shape[0]), cout, cin, k, float(scale), float(eps), C.c_void_p(st)), "hav_modconv_demod")
    return out


def is_cl(t):
    """Channels-last fp16 hand-over tensors ([B,H,W,C] torch.float16) are told apart from NCHW fp32 ones by dtype."""
    return t.dtype == torch.float16


def to_nchw(t):
    return t.permute(0, 3, 1, 2).float().contiguous() if is_cl(t) else t


def conv2d(x, packed, in_scale=None, out_scale=None, noise=None, noise_weight=0.0, bias=None, act=False, up=1, down=1, out_cl=False,
           residual=None):
    """out = act(conv(x * in_scale[:, :, None, None], W) * out_scale[:, :, None, None] + noise_weight * noise + bias) [+ residual].
    residual: a tensor of the output's shape and layout, added after the activation in the same launch.
    up=2: conv_transpose2d(stride 2, pad 0) (pack the weight with up=2); down=2: stride 2, pad 0; else pad k//2.
    x is NCHW float32, or channels-last float16 [B,H,W,C] (the internal hand-over layout); out_cl selects the output layout."""
    L = _lib.lib()
    in_cl = is_cl(x)
    if in_cl:
        if not x.is_cuda or x.dim() != 4:
            raise _lib.HavError("channels-last input must be a [B,H,W,C] float16 CUDA tensor")
        x = x.detach().contiguous()
        B, H, W, cin = [int(v) for v in x.shape]
    else:
        x = _check(x, "x")
        B, cin, H, W = [int(v) for v in x.shape]
    if (in_cl or out_cl) and packed.precision != "fp16":
        raise _lib.HavError("channels-last tensors need fp16-packed weights")
    if cin != packed.cin:
        raise _lib.HavError("x has %d channels, the packed weight expects %d" % (cin, packed.cin))
    k = packed.ksize
    if int(up) != packed.up:
        raise _lib.HavError("weights were packed for up=%d" % packed.up)
    if up == 2:
        Ho, Wo = 2 * H - 1 + k - 1, 2 * W - 1 + k - 1
    elif down == 2:
        Ho, Wo = (H - k) // 2 + 1, (W - k) // 2 + 1
    else:
        Ho, Wo = H, W
    a = _lib.ConvArgs()
    a.struct_bytes = C.sizeof(_lib.ConvArgs)
    a.precision = _lib.PRECISIONS[packed.precision]
    a.batch, a.cin, a.cout, a.in_h, a.in_w = B, cin, packed.cout, H, W
    a.ksize, a.up, a.down, a.act = k, int(up), int(down), int(bool(act))
    keep = [x, packed.data]
    a.x, a.wpack = C.c_void_p(x.data_ptr()), C.c_void_p(packed.data.data_ptr())
    for name, t, shape in (("in_scale", in_scale, (B, cin)), ("out_scale", out_scale, (B, packed.cout)), ("bias", bias, None)):
        if t is not None:
            t = _check(t, name)
            if shape is not None and tuple(t.shape) != shape:
                raise _lib.HavError("%s must be %s" % (name, shape))
            if name == "bias" and t.numel() != packed.cout:
                raise _lib.HavError("bias must have Cout elements")
            keep.append(t)
            setattr(a, name, C.c_void_p(t.data_ptr()))
    if noise is not None:
        nz = _check(noise, "noise")
        if nz.numel() == Ho * Wo:
            a.noise_per_sample = 0
        elif nz.numel() == B * Ho * Wo:
            a.noise_per_sample = 1
        else:
            raise _lib.HavError("noise must be [1 or B,1,%d,%d]" % (Ho, Wo))
        a.noise, a.noise_weight = C.c_void_p(nz.data_ptr()), float(noise_weight)
        keep.append(nz)
    a.in_layout, a.out_layout = int(in_cl), int(bool(out_cl))
    if out_cl:
        out = torch.empty((B, Ho, Wo, packed.cout), dtype=torch.float16, device=x.device)
    else:
        out = torch.empty((B, packed.cout, Ho, Wo), dtype=torch.float32, device=x.device)
    a.out = C.c_void_p(out.data_ptr())
    if residual is not None:
        if residual.shape != out.shape or residual.dtype != out.dtype or residual.device != out.device:
            raise _lib.HavError("residual must have the output's shape %s, dtype and device" % (tuple(out.shape),))
        residual = residual.detach().contiguous()
        keep.append(residual)
        a.residual = C.c_void_p(residual.data_ptr())
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(L.hav_conv2d_forward(C.byref(a), C.c_void_p(st)), "hav_conv2d_forward")
    del keep
    return out


def upfirdn2d_cl(x, kernel, up=1, down=1, pad=(0, 0), noise=None, noise_weight=0.0, bias=None, act=False):
    """upfirdn2d on a channels-last float16 tensor [B,H,W,C] with the StyledConv tail fused in:
    act(fir(x) + noise_weight * noise + bias) (C ABI: hav_upfirdn2d_cl)."""
    L = _lib.lib()
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float16 and x.dim() == 4):
        raise _lib.HavError("x must be a [B,H,W,C] float16 CUDA tensor")
    x = x.detach().contiguous()
    k = _check(kernel, "kernel")
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    B, H, W, Cc = [int(v) for v in x.shape]
    kh, kw = int(k.shape[0]), int(k.shape[1])
    Ho = (H * up + pad[2] + pad[3] - kh + down) // down
    Wo = (W * up + pad[0] + pad[1] - kw + down) // down
    out = torch.empty((B, Ho, Wo, Cc), dtype=torch.float16, device=x.device)
    nz = None if noise is None else _check(noise, "noise")
    per_sample = 0
    if nz is not None:
        if nz.numel() == B * Ho * Wo and B > 1:
            per_sample = 1
        elif nz.numel() != Ho * Wo:
            raise _lib.HavError("noise must be [1 or B,1,%d,%d]" % (Ho, Wo))
    bs = None if bias is None else _check(bias, "bias")
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(L.hav_upfirdn2d_cl(C.c_void_p(out.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(k.data_ptr()), B, H, W, Cc, kh, kw,
                                      int(up), int(down), int(pad[0]), int(pad[1]), int(pad[2]), int(pad[3]),
                                      C.c_void_p(nz.data_ptr()) if nz is not None else None, float(noise_weight), per_sample,
                                      C.c_void_p(bs.data_ptr()) if bs is not None else None, int(bool(act)), C.c_void_p(st)),
                   "hav_upfirdn2d_cl")
    return out


def conv_wgrad(g, x, ksize, in_scale=None, out_scale=None, wscale=1.0, up=1, down=1, out=None):
    """dW [Cout,Cin,k,k] of  y = out_scale * conv(in_scale * x, wscale * W)  given g = dL/dy (C ABI: hav_conv2d_wgrad).
    `out` accumulates into an existing gradient tensor."""
    L = _lib.lib()
    g, x = _check(g, "g"), _check(x, "x")
    B, cin, H, W = [int(v) for v in x.shape]
    cout = int(g.shape[1])
    if up == 2:
        Ho, Wo = 2 * H + 1, 2 * W + 1
    elif down == 2:
        Ho, Wo = (H - ksize) // 2 + 1, (W - ksize) // 2 + 1
    else:
        Ho, Wo = H, W
    if tuple(g.shape) != (B, cout, Ho, Wo):
        raise _lib.HavError("g must be %s, got %s" % ((B, cout, Ho, Wo), tuple(g.shape)))
    a = _lib.ConvWgradArgs()
    a.struct_bytes = C.sizeof(_lib.ConvWgradArgs)
    a.batch, a.cin, a.cout, a.in_h, a.in_w = B, cin, cout, H, W
    a.ksize, a.up, a.down, a.accumulate, a.wscale = int(ksize), int(up), int(down), int(out is not None), float(wscale)
    keep = [g, x]
    a.g, a.x = C.c_void_p(g.data_ptr()), C.c_void_p(x.data_ptr())
    for name, t, shape in (("in_scale", in_scale, (B, cin)), ("out_scale", out_scale, (B, cout))):
        if t is not None:
            t = _check(t, name)
            if tuple(t.shape) != shape:
                raise _lib.HavError("%s must be %s" % (name, shape))
            keep.append(t)
            setattr(a, name, C.c_void_p(t.data_ptr()))
    if out is None:
        out = torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=x.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == cout * cin * ksize * ksize):
        raise _lib.HavError("out must be a contiguous float32 CUDA tensor with Cout*Cin*k*k elements")
    a.dw = C.c_void_p(out.data_ptr())
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(L.hav_conv2d_wgrad(C.byref(a), C.c_void_p(st)), "hav_conv2d_wgrad")
    del keep
    return out


def rowscale_dot(a, x=None, scale=None, want_out=True, want_dot=True):
    """a, x [B,C,H,W]; scale [B,C].  Returns (a * scale[:, :, None, None] or None, (a * x).sum((2, 3)) or None) in one pass
    (C ABI: hav_rowscale_dot)."""
    L = _lib.lib()
    a = _check(a, "a")
    B, Cc = int(a.shape[0]), int(a.shape[1])
    n = a.numel() // max(B * Cc, 1)
    x = None if x is None else _check(x, "x")
    scale = None if scale is None else _check(scale, "scale")
    if x is not None and x.shape != a.shape:
        raise _lib.HavError("x must have the shape of a")
    out = torch.empty_like(a) if want_out else None
    dot = torch.empty((B, Cc), dtype=torch.float32, device=a.device) if want_dot else None
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.device(a.device):
        st = torch.cuda.current_stream(a.device).cuda_stream
        _lib.check(L.hav_rowscale_dot(p(out), p(dot), p(a), p(x), p(scale), B * Cc, n, C.c_void_p(st)), "hav_rowscale_dot")
    return out, dot


def _dgrad_raw(g, weight, wscale, up, down, H, W, in_scale=None):
    """Data gradient of conv(x, wscale * weight) = the forward kernel on g with the transposed weight image (mirrored taps for
    stride 1); the gradient of a stride-2 convolution is a transposed stride-2 convolution and vice versa.  in_scale multiplies
    g per (sample, output channel) while it is staged (the demodulation of a modulated layer)."""
    if up == 2:
        wp = pack_weights(weight, wscale, up=1, transpose_io=True, precision="bf16")
        return conv2d(g, wp, in_scale=in_scale, down=2)
    if down == 2:
        wp = pack_weights(weight, wscale, up=2, transpose_io=True, precision="bf16")
        dxs = conv2d(g, wp, in_scale=in_scale, up=2)
        if dxs.shape[2] != H or dxs.shape[3] != W:       # even input: its last row / column never reached an output
            dxs = torch.nn.functional.pad(dxs, (0, W - dxs.shape[3], 0, H - dxs.shape[2]))
        return dxs
    wp = pack_weights(weight, wscale, up=1, transpose_io=True, flip=True, precision="bf16")
    return conv2d(g, wp, in_scale=in_scale)


# model/op/conv2d_gradfix.py:12-20 no_weight_gradients(): while set, a backward that is itself being recorded (the R1 penalty's
# autograd.grad(..., create_graph=True), utils/styleUnet_util.py:72-79) skips the weight gradient it would otherwise compute and
# throw away.  Toggled through havatar_b200.op.conv2d_gradfix.no_weight_gradients().
WEIGHT_GRADIENTS_DISABLED = [False]


# The convolution is bilinear in (x, w), so its three kernels are closed under differentiation:
#     y  = F(x, w)          dF:  dx = D(gy, w)   dw = W(gy, x)
#     dx = D(g, w)          dD:  dg = F(h, w)    dw = W(g, h)
#     dw = W(g, x)          dW:  dg = F(x, V)    dx = D(g, V)
# Each is an autograd node whose backward applies the other two, which gives gradients of every order (the R1 penalty
# differentiates the discriminator's input gradient, utils/styleUnet_util.py:72-79) without any library convolution.
class _Fwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, wscale, up, down, precision):
        ctx.save_for_backward(x, weight)
        ctx.cfg = (wscale, up, down)
        return conv2d(x, pack_weights(weight, wscale, up=up, precision=precision), up=up, down=down)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        wscale, up, down = ctx.cfg
        dx = _Dgrad.apply(gy, weight, wscale, up, down, int(x.shape[2]), int(x.shape[3])) if ctx.needs_input_grad[0] else None
        dw = _Wgrad.apply(gy, x, wscale, up, down, int(weight.shape[-1])) \
            if ctx.needs_input_grad[1] and not WEIGHT_GRADIENTS_DISABLED[0] else None
        return dx, dw, None, None, None, None


class _Dgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, weight, wscale, up, down, H, W):
        ctx.save_for_backward(g, weight)
        ctx.cfg = (wscale, up, down)
        return _dgrad_raw(g.contiguous(), weight, wscale, up, down, H, W)

    @staticmethod
    def backward(ctx, h):
        g, weight = ctx.saved_tensors
        wscale, up, down = ctx.cfg
        dg = _Fwd.apply(h.contiguous(), weight, wscale, up, down, "bf16") if ctx.needs_input_grad[0] else None
        dw = _Wgrad.apply(g, h.contiguous(), wscale, up, down, int(weight.shape[-1])) if ctx.needs_input_grad[1] else None
        return dg, dw, None, None, None, None, None


class _Wgrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, x, wscale, up, down, k):
        ctx.save_for_backward(g, x)
        ctx.cfg = (wscale, up, down)
        return conv_wgrad(g.contiguous(), x.contiguous(), k, wscale=wscale, up=up, down=down)

    @staticmethod
    def backward(ctx, v):
        g, x = ctx.saved_tensors
        wscale, up, down = ctx.cfg
        v = v.contiguous()
        dg = _Fwd.apply(x, v, wscale, up, down, "bf16") if ctx.needs_input_grad[0] else None
        dx = _Dgrad.apply(g, v, wscale, up, down, int(x.shape[2]), int(x.shape[3])) if ctx.needs_input_grad[1] else None
        return dg, dx, None, None, None, None


class _ConvFunction(torch.autograd.Function):
    """y = out_scale[b,co] * conv(in_scale[b,ci] * x, wscale * weight): ModulatedConv2d in its shared-weight form
    (model/styleUnet.py:225-251) and, with both scales None, EqualConv2d (:108-118).  Forward: fp16 operands; backward: bf16
    operands (gradients have no fixed range), fp32 accumulation everywhere.  First-order backward is the fused path (scales
    applied while staging, one row kernel for the modulation gradients); when the backward itself is being recorded
    (create_graph=True) it is composed from the differentiable nodes _Dgrad / _Wgrad and torch elementwise ops instead."""

    @staticmethod
    def forward(ctx, x, weight, in_scale, out_scale, wscale, up, down, cache):
        k = int(weight.shape[-1])
        packed = pack_weights_cached(weight, wscale, up=up, precision="fp16", cache=cache)
        y = conv2d(x, packed, in_scale=in_scale, out_scale=out_scale, up=up, down=down)
        ctx.save_for_backward(x, weight, in_scale, out_scale, y if out_scale is not None else None)
        ctx.cfg = (float(wscale), int(up), int(down), k)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, in_scale, out_scale, y = ctx.saved_tensors
        wscale, up, down, k = ctx.cfg
        need = ctx.needs_input_grad
        H, W = int(x.shape[2]), int(x.shape[3])
        dx = dw = ds = dd = None
        if torch.is_grad_enabled():      # create_graph=True: every step below must itself be differentiable
            gs = g if out_scale is None else g * out_scale[:, :, None, None]
            if need[0] or (in_scale is not None and need[2]):
                dxs = _Dgrad.apply(gs, weight, wscale, up, down, H, W)
                dx = dxs if in_scale is None else dxs * in_scale[:, :, None, None]
                if in_scale is not None and need[2]:
                    ds = (dxs * x).sum(dim=(2, 3))
            if need[1] and not WEIGHT_GRADIENTS_DISABLED[0]:
                dw = _Wgrad.apply(gs, x if in_scale is None else x * in_scale[:, :, None, None], wscale, up, down, k)
            if out_scale is not None and need[3]:
                dd = (g * y).sum(dim=(2, 3)) / out_scale
            return dx, dw, ds, dd, None, None, None, None
        g = g.contiguous()
        if need[0] or (in_scale is not None and need[2]):
            dxs = _dgrad_raw(g, weight, wscale, up, down, H, W, in_scale=out_scale)
            if in_scale is not None:
                dx, ds = rowscale_dot(dxs, x, in_scale, want_out=need[0], want_dot=need[2])
            else:
                dx = dxs
        if need[1]:
            dw = conv_wgrad(g, x, k, in_scale=in_scale, out_scale=out_scale, wscale=wscale, up=up, down=down)
        if out_scale is not None and need[3]:
            _, gy = rowscale_dot(g, y, None, want_out=False)
            dd = gy / out_scale
        return dx, dw, ds, dd, None, None, None, None


def conv2d_autograd(x, weight, in_scale=None, out_scale=None, wscale=1.0, up=1, down=1, cache=None):
    """Differentiable  out_scale * conv(in_scale * x, wscale * weight)  on the tcgen05 kernels: gradients of any order w.r.t. x,
    weight [Cout,Cin,k,k], in_scale [B,Cin] and out_scale [B,Cout].  cache: the owning module's dict of packed weight images
    (pack_weights_cached), shared with its inference path."""
    return _ConvFunction.apply(x, weight, in_scale, out_scale, float(wscale), int(up), int(down), cache)
