"""StyleUNet on the B200 kernels: drop-in, checkpoint-compatible mirrors of the reference's model/styleUnet.py networks.

    SWGAN_unet     <- model/styleUnet.py:1190-1410  (HD upsampler, avatarHD_reenactment.py:139-167)
    StyleGAN_zxc   <- model/styleUnet.py:631-878    (plane generators XY_gen / YZ_gen, model/nerf_model.py:39-42, :58-86)

Same constructor arguments, same `forward` signatures and the same state_dict keys / shapes, so `load_state_dict` of a
reference checkpoint works unchanged.  With gradients enabled the forward is the differentiable formulation of
styleunet_train.py (training steps); under torch.no_grad() (inference) every convolution runs on the tcgen05 implicit-GEMM kernel
(havatar_b200/conv.py) in the shared-weight formulation of ModulatedConv2d's own non-fused branch (styleUnet.py:225-251),
with modulation, demodulation, noise, bias and leaky-relu fused into that launch where the layer has no blur in between;
blur / up / down / Haar go through the upfirdn2d kernel, the style MLP's activation through fused_bias_act.
torch is used for parameters, device memory, concatenation and the [B,64] style MLP matmuls.  No CPU fallback.
"""
import math
import random

import torch
from torch import nn

import ctypes as C

from . import _lib
from . import conv as hconv
from . import styleunet_train as T
from .op import fused_leaky_relu, upfirdn2d


def _fir(taps, gain=1.0):
    k = torch.tensor(taps, dtype=torch.float32)
    k = k[None, :] * k[:, None]
    return k / k.sum() * gain


class PixelNorm(nn.Module):
    def forward(self, x):
        return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)


class Blur(nn.Module):
    """styleUnet.py:70-86."""

    def __init__(self, taps, pad, upsample_factor=1):
        super().__init__()
        self.register_buffer("kernel", _fir(taps, float(upsample_factor ** 2)))
        self.pad = pad

    def forward(self, x, **tail):
        if hconv.is_cl(x):          # channels-last fp16 hand-over: FIR + (noise, bias, lrelu) in one launch
            return hconv.upfirdn2d_cl(x, self.kernel, pad=self.pad, **tail)
        return upfirdn2d(x, self.kernel, pad=self.pad)


class Upsample(nn.Module):
    """styleUnet.py:28-46."""

    def __init__(self, taps, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", _fir(taps, float(factor ** 2)))
        p = len(taps) - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, x):
        return upfirdn2d(x, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """styleUnet.py:49-67."""

    def __init__(self, taps, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", _fir(taps))
        p = len(taps) - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, x):
        if hconv.is_cl(x):          # channels-last fp16 hand-over (the condition pyramid of a channels-last condition image)
            return hconv.upfirdn2d_cl(x, self.kernel, up=1, down=self.factor, pad=self.pad)
        return upfirdn2d(x, self.kernel, up=1, down=self.factor, pad=self.pad)


def _haar():
    s = 1 / math.sqrt(2)
    lo, hi = torch.tensor([[s, s]]), torch.tensor([[-s, s]])
    return lo.T * lo, hi.T * lo, lo.T * hi, hi.T * hi      # ll, lh, hl, hh (styleUnet.py:371-381)


class HaarTransform(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        for n, k in zip(("ll", "lh", "hl", "hh"), _haar()):
            self.register_buffer(n, k)

    def forward(self, x):
        return torch.cat([upfirdn2d(x, k, down=2) for k in (self.ll, self.lh, self.hl, self.hh)], 1)


class InverseHaarTransform(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        ll, lh, hl, hh = _haar()
        for n, k in zip(("ll", "lh", "hl", "hh"), (ll, -lh, -hl, hh)):
            self.register_buffer(n, k)

    def forward(self, x):
        parts = x.chunk(4, 1)
        out = None
        for p, k in zip(parts, (self.ll, self.lh, self.hl, self.hh)):
            y = upfirdn2d(p.contiguous(), k, up=2, pad=(1, 0, 1, 0))
            out = y if out is None else out + y
        return out


# Derived tensors (packed weights, scaled linears, host copies of scalars) are cached per parameter version.  A CUDA-graph
# replay changes parameters without touching their version counters, so train_step.Graphed bumps this epoch after every
# replay and every cache key carries it.
_EPOCH = [0]


def invalidate_caches():
    _EPOCH[0] += 1


class _PackCache:
    """Packed 16-bit weight images of one conv parameter (one per layout / precision), rebuilt when the parameter is modified
    in place or replaced; shared by the module's inference path and its differentiable path (conv.pack_weights_cached)."""

    def __init__(self):
        self.store = {}

    def get(self, weight, scale, up):
        w = weight.detach()
        if w.dim() == 5:
            w = w[0]
        return hconv.pack_weights_cached(w.float(), scale, up=2 if up else 1, cache=self.store)


class EqualConv2d(nn.Module):
    """styleUnet.py:88-123; forward is driven by ConvLayer so that the activation fuses into the conv launch."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride, self.padding = stride, padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None
        self._cache = _PackCache()

    def run(self, x, bias=None, act=False, out_cl=False, residual=None):
        b = bias if bias is not None else self.bias
        return hconv.conv2d(x, self._cache.get(self.weight, self.scale, False), bias=b, act=act, down=self.stride, out_cl=out_cl,
                            residual=residual)

    def forward(self, x):
        return self.run(x)


class _ActBias(nn.Module):
    """FusedLeakyReLU's parameter holder (state_dict key `bias`, model/op/fused_act.py:90-104); applied inside the conv launch."""

    def __init__(self, channel, bias=True):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None

    def forward(self, x):
        return fused_leaky_relu(x, self.bias)


class ConvLayer(nn.Sequential):
    """styleUnet.py:326-368: [Blur] -> EqualConv2d -> [FusedLeakyReLU]; here blur, then ONE conv launch with bias + lrelu fused."""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=(1, 3, 3, 1), bias=True, activate=True):
        layers = []
        if downsample:
            p = (len(blur_kernel) - 2) + (kernel_size - 1)
            layers.append(Blur(list(blur_kernel), pad=((p + 1) // 2, p // 2)))
            stride, padding = 2, 0
        else:
            stride, padding = 1, kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=padding, stride=stride, bias=bias and not activate))
        if activate:
            layers.append(_ActBias(out_channel, bias=bias))
        super().__init__(*layers)
        self.activate = activate

    def forward(self, x, out_cl=False, residual=None):
        mods = list(self)
        if isinstance(mods[0], Blur):
            x = mods[0](x)
            mods = mods[1:]
        act_bias = mods[1].bias if self.activate else None
        return mods[0].run(x, bias=act_bias, act=self.activate, out_cl=out_cl, residual=residual)


class EqualLinear(nn.Module):
    """styleUnet.py:126-162."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul
        self._key, self._w, self._b = None, None, None

    def _scaled(self):
        """weight * scale and bias * lr_mul, recomputed only when a parameter changes (two launches saved per call)."""
        key = (self.weight.data_ptr(), self.weight._version, None if self.bias is None else (self.bias.data_ptr(), self.bias._version),
               _EPOCH[0])
        if key != self._key:
            self._w = (self.weight.detach() * self.scale).contiguous()
            self._b = None if self.bias is None else (self.bias.detach() * self.lr_mul).contiguous()
            self._key = key
        return self._w, self._b

    def forward(self, x):
        w, b = self._scaled()
        if self.activation:
            return fused_leaky_relu(torch.nn.functional.linear(x, w).contiguous(), b)
        return torch.nn.functional.linear(x, w, bias=b)


class ModulatedConv2d(nn.Module):
    """styleUnet.py:165-297.  run() adds the StyledConv / ToRGB epilogue to the same launch when no blur follows."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False, downsample=False,
                 blur_kernel=(1, 3, 3, 1), fused=True):
        super().__init__()
        if downsample:
            raise NotImplementedError("ModulatedConv2d(downsample=True) is never instantiated by the reference networks")
        self.eps = 1e-8
        self.kernel_size, self.in_channel, self.out_channel = kernel_size, in_channel, out_channel
        self.upsample, self.downsample = upsample, downsample
        if upsample:
            p = (len(blur_kernel) - 2) - (kernel_size - 1)
            self.blur = Blur(list(blur_kernel), pad=((p + 1) // 2 + 1, p // 2 + 1), upsample_factor=2)
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        self.fused = fused
        self._cache = _PackCache()

    def run(self, x, style, noise=None, noise_weight=0.0, bias=None, act=False, out_cl=False, sd=None, residual=None):
        if sd is not None:          # modulation / demodulation of the whole network computed up front (_StylePlan)
            s, d = sd
        else:
            s = self.modulation(style).contiguous()
            d = hconv.modconv_demod(self.weight.detach()[0], s, self.scale, self.eps) if self.demodulate else None
        packed = self._cache.get(self.weight, self.scale, self.upsample)
        if residual is not None and self.upsample:
            raise hconv._lib.HavError("a fused residual is only available on the non-upsampling convolution")
        if not self.upsample:
            return hconv.conv2d(x, packed, in_scale=s, out_scale=d, noise=noise, noise_weight=noise_weight, bias=bias, act=act,
                                out_cl=out_cl, residual=residual)
        # convT stride 2 (polyphase) -> 4x4 blur (:264-277); both in channels-last fp16, the StyledConv tail rides on the blur
        y = hconv.conv2d(x, packed, in_scale=s, out_scale=d, up=2, out_cl=True)
        y = self.blur(y, noise=noise, noise_weight=noise_weight, bias=bias, act=act)
        return y if out_cl else hconv.to_nchw(y)

    def forward(self, x, style):
        return self.run(x, style)


class NoiseInjection(nn.Module):
    """styleUnet.py:300-310 (parameter holder; the addition happens in the conv epilogue)."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))
        self._key, self._val, self._misses = None, 0.0, 0

    def value(self):
        key = (self.weight.data_ptr(), self.weight._version, _EPOCH[0])
        if key != self._key:                      # one device->host read per weight update, not per forward
            self._val, self._key = float(self.weight.detach().cpu()), key
        return self._val

    def scaled(self, noise):
        """(noise tensor, host weight) for the fused epilogue.  While the weight is static (inference) the host copy of the
        scalar is cached; while it changes every iteration (training: a device->host read per layer per step would serialise
        the host with the GPU, and cannot be captured into a CUDA graph) the product is formed on the device instead."""
        key = (self.weight.data_ptr(), self.weight._version, _EPOCH[0])
        if key == self._key:
            return noise, self._val
        if torch.cuda.is_current_stream_capturing() or self._misses >= 2:
            return self.weight.detach() * noise, 1.0
        self._misses += 1
        return noise, self.value()

    def forward(self, image, noise=None):
        if noise is None:
            b, _, h, w = image.shape
            noise = image.new_empty(b, 1, h, w).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, x):
        return self.input.repeat(x.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """styleUnet.py:565-599: modulated conv -> noise -> bias + leaky-relu, one launch (two + blur when upsampling)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=(1, 3, 3, 1), demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample, blur_kernel=blur_kernel,
                                    demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = _ActBias(out_channel)

    def forward(self, x, style, noise=None, out_cl=False, sd=None):
        if noise is None:     # fresh N(0,1) per call, like NoiseInjection.forward (:306-309)
            h, w = (x.shape[1], x.shape[2]) if hconv.is_cl(x) else (x.shape[2], x.shape[3])
            f = 2 if self.conv.upsample else 1
            noise = torch.empty(x.shape[0], 1, h * f, w * f, dtype=torch.float32, device=x.device).normal_()
        noise, nw = self.noise.scaled(noise)
        return self.conv.run(x, style, noise=noise, noise_weight=nw, bias=self.activate.bias, act=True, out_cl=out_cl, sd=sd)


class ToRGB(nn.Module):
    """styleUnet.py:602-628."""

    def __init__(self, in_channel, style_dim, out_channel=12, upsample=True, blur_kernel=(1, 3, 3, 1), use_wt=True):
        super().__init__()
        self.use_wt = use_wt
        if upsample:
            self.upsample = Upsample(list(blur_kernel))
            if use_wt:
                self.iwt = InverseHaarTransform(3)
                self.dwt = HaarTransform(3)
        self.out_channel = out_channel if use_wt else out_channel // 4
        self.conv = ModulatedConv2d(in_channel, self.out_channel, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, self.out_channel, 1, 1))

    def forward(self, x, style, skip=None, sd=None):
        if skip is not None:          # `out + skip` (:625-626) rides on the conv's epilogue
            skip = self.dwt(self.upsample(self.iwt(skip))) if self.use_wt else self.upsample(skip)
        return self.conv.run(x, style, bias=self.bias.view(-1), sd=sd, residual=skip)


class _StylePlan:
    """Every modulation vector and demodulation factor of one network's inference forward in two launches
    (hav_style_plan_run) instead of one small linear + one reduction per layer (styleUnet.py:237-258).  `entries` lists the
    network's ModulatedConv2d modules with the latent index each one reads, in call order.  The device table holds pointers,
    sizes and scales only; the tap-summed squared weights it points to are refreshed in place when a weight changes."""

    MAX_BATCH = 8

    def __init__(self, entries):
        self.entries = entries
        self._tables = {}        # batch -> (pointer key, device table, prefixes, offsets, pinned host copies)
        self._wsq = {}           # entry index -> [buffer, version key]

    def _wsq_of(self, i, m):
        w = m.weight.detach()[0]
        slot = self._wsq.get(i)
        if slot is None or slot[0].device != w.device:
            slot = self._wsq[i] = [torch.empty(w.shape[0], w.shape[1], dtype=torch.float32, device=w.device), None]
        key = (w.data_ptr(), m.weight._version, _EPOCH[0])
        if slot[1] != key:
            # refreshed in place (the table keeps pointing at it).  A training-step capture starts with a new cache epoch, so
            # the refresh is captured with it and replayed with the weights of each iteration
            with torch.cuda.device(w.device):
                st = torch.cuda.current_stream(w.device).cuda_stream
                _lib.check(_lib.lib().hav_conv_tap_squares(C.c_void_p(slot[0].data_ptr()), C.c_void_p(w.contiguous().data_ptr()),
                                                           int(w.shape[0]), int(w.shape[1]), int(w.shape[2]), C.c_void_p(st)),
                           "hav_conv_tap_squares")
            slot[1] = key
        return slot[0]

    def _table(self, B, device):
        wsq = [self._wsq_of(i, m) if m.demodulate else None for i, (m, _) in enumerate(self.entries)]
        key = (str(device),) + tuple((m.modulation.weight.data_ptr(), 0 if m.modulation.bias is None else m.modulation.bias.data_ptr(),
                                      0 if q is None else q.data_ptr()) for (m, _), q in zip(self.entries, wsq))
        hit = self._tables.get(B)
        if hit is not None and hit[0] == key:
            return hit
        n = len(self.entries)
        arr = (_lib.StyleLayer * n)()
        s_prefix, d_prefix, s_off, d_off, offs = [0], [0], 0, 0, []
        for i, ((m, li), q) in enumerate(zip(self.entries, wsq)):
            lin = m.modulation
            a = arr[i]
            a.mod_w, a.mod_b = lin.weight.data_ptr(), (lin.bias.data_ptr() if lin.bias is not None else None)
            a.wsq = q.data_ptr() if q is not None else None
            a.cin, a.cout, a.latent_index, a.s_off, a.d_off = m.in_channel, m.out_channel, int(li), s_off, d_off
            a.mod_scale, a.mod_lr_mul, a.conv_scale = lin.scale, lin.lr_mul, m.scale
            offs.append((s_off, d_off if q is not None else None))
            s_off += B * m.in_channel
            s_prefix.append(s_prefix[-1] + m.in_channel)
            if q is not None:
                d_off += B * m.out_channel
                d_prefix.append(d_prefix[-1] + m.out_channel)
            else:
                d_prefix.append(d_prefix[-1])
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().pin_memory()
        pref = torch.tensor([s_prefix, d_prefix], dtype=torch.int32).pin_memory()
        table = torch.empty(host.numel(), dtype=torch.uint8, device=device)
        pref_dev = torch.empty((2, n + 1), dtype=torch.int32, device=device)
        table.copy_(host, non_blocking=True)          # pinned sources: legal inside a CUDA-graph capture too
        pref_dev.copy_(pref, non_blocking=True)
        hit = self._tables[B] = (key, table, pref_dev, offs, (s_off, d_off, s_prefix[-1], d_prefix[-1]), (host, pref))
        return hit

    def run(self, latent):
        """latent [B, n_latent, D] -> [(s [B,Cin], d [B,Cout] or None)] per entry, or None when the batched path does not apply."""
        B = int(latent.shape[0])
        if not latent.is_cuda or B > self.MAX_BATCH or any(m.modulation.activation for m, _ in self.entries):
            return None
        lat = latent.detach().float().contiguous()
        _, table, pref, offs, (s_tot, d_tot, s_rows, d_rows), _ = self._table(B, lat.device)
        s_all = torch.empty(s_tot, dtype=torch.float32, device=lat.device)
        d_all = torch.empty(max(d_tot, 1), dtype=torch.float32, device=lat.device)
        with torch.cuda.device(lat.device):
            st = torch.cuda.current_stream(lat.device).cuda_stream
            _lib.check(_lib.lib().hav_style_plan_run(C.c_void_p(s_all.data_ptr()), C.c_void_p(d_all.data_ptr()), C.c_void_p(lat.data_ptr()), B,
                                                     int(lat.shape[1]), int(lat.shape[2]), C.c_void_p(table.data_ptr()),
                                                     C.c_void_p(pref[0].data_ptr()), C.c_void_p(pref[1].data_ptr()), len(self.entries),
                                                     s_rows, d_rows, 1e-8, C.c_void_p(st)), "hav_style_plan_run")
        out = []
        for (m, _), (so, do) in zip(self.entries, offs):
            s = s_all[so:so + B * m.in_channel].view(B, m.in_channel)
            d = None if do is None else d_all[do:do + B * m.out_channel].view(B, m.out_channel)
            out.append((s, d))
        return out


def _plan_beside(plan, latent, fn):
    """(plan.run(latent), fn()): the style plan's two launches run on a side stream beside fn (the condition encoder, which
    reads no style), and are joined before the first modulated convolution."""
    if not latent.is_cuda:
        return plan.run(latent), fn()
    from . import pipeline

    dev = latent.device
    main, side = torch.cuda.current_stream(dev), pipeline.aux_stream(dev, 5)
    pipeline.note_fork(dev, main, side)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        sd = plan.run(latent)
    out = fn()
    main.wait_stream(side)
    if sd is not None:
        seen = set()                     # all views share two buffers
        for pair in sd:
            for t_ in pair:
                if t_ is not None and t_.untyped_storage().data_ptr() not in seen:
                    seen.add(t_.untyped_storage().data_ptr())
                    t_.record_stream(main)
    return sd, out


class ConvBlock(nn.Module):
    def __init__(self, in_channel, out_channel, blur_kernel=(1, 3, 3, 1), downsample=True):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=downsample)

    def forward(self, x, out_cl=False):
        return self.conv2(self.conv1(x, out_cl=True), out_cl=out_cl)


class FromRGB(nn.Module):
    """styleUnet.py:439-467."""

    def __init__(self, out_channel, in_channel, downsample=True, blur_kernel=(1, 3, 3, 1), use_wt=True):
        super().__init__()
        self.use_wt = use_wt
        self.downsample = downsample
        if downsample:
            self.downsample = Downsample(list(blur_kernel))
            if use_wt:
                self.iwt = InverseHaarTransform(in_channel)
                self.dwt = HaarTransform(in_channel)
        self.in_channel = in_channel * 4 if use_wt else in_channel
        self.conv = ConvLayer(self.in_channel, out_channel, 1)

    def forward(self, x, skip=None, out_cl=False):
        if self.downsample:
            x = self.dwt(self.downsample(self.iwt(x))) if self.use_wt else self.downsample(x)
        cl = out_cl or (skip is not None and hconv.is_cl(skip))
        if skip is not None and hconv.is_cl(skip) != cl:
            return x, self.conv(x, out_cl=cl) + skip
        return x, self.conv(x, out_cl=cl, residual=skip)          # `out + skip` (:464-465) rides on the conv's epilogue


_CHANNELS = lambda m: {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * m, 128: 128 * m, 256: 64 * m, 512: 32 * m, 1024: 16 * m}


def _style_mlp(in_dim, dim, n_mlp, lr_mlp):
    layers = [PixelNorm(), EqualLinear(in_dim, dim, lr_mul=lr_mlp, activation="fused_lrelu")]
    layers += [EqualLinear(dim, dim, lr_mul=lr_mlp, activation="fused_lrelu") for _ in range(n_mlp - 1)]
    return nn.Sequential(*layers)


def _latents(styles, n_latent, inject_index):
    """styleUnet.py:1363-1377 / :823-838: one style -> repeated; two -> mixed at inject_index."""
    if len(styles) < 2:
        return styles[0].unsqueeze(1).repeat(1, n_latent, 1) if styles[0].ndim < 3 else styles[0]
    if inject_index is None:
        inject_index = random.randint(1, n_latent - 1)
    a = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
    b = styles[1].unsqueeze(1).repeat(1, n_latent - inject_index, 1)
    return torch.cat([a, b], 1)


class _CondEncoder:
    """The shared condition-image encoder of both networks (styleUnet.py:1379-1388, :847-856)."""

    @staticmethod
    def run(net, cond_img):
        cond_out = net.conv_in(cond_img, out_cl=True)          # feature maps travel channels-last fp16 between layers
        feats = [cond_out]
        for from_rgb, cond_conv in zip(net.from_rgbs, net.cond_convs):
            cond_img, cond_out = from_rgb(cond_img, cond_out, out_cl=True)
            cond_out = cond_conv(cond_out, out_cl=True)
            feats.append(cond_out)
        return feats


class _SkipFork:
    """The ToRGB pyramid of SWGAN_unet (styleUnet.py:1398-1406) depends on each level's feature map but nothing on the main
    path depends on it until the final inverse wavelet transform: its ~25 small launches per level (1x1 modulated convolution,
    Haar synthesis -> x2 upsample -> Haar analysis of the running skip, adds) run on a side stream beside the next level's
    convolutions.  `with fork.branch(x): ...` runs the body on the side stream after everything issued so far; `join(t)` hands
    the result back to the main stream."""

    def __init__(self, device):
        self.on = torch.device(device).type == "cuda"
        if self.on:
            from . import pipeline

            self.main = torch.cuda.current_stream(device)
            self.side = pipeline.aux_stream(torch.device(device), 3)
            pipeline.note_fork(torch.device(device), self.main, self.side)

    def branch(self, *consumed):
        import contextlib

        if not self.on:
            return contextlib.nullcontext()
        self.side.wait_stream(self.main)
        for t in consumed:
            t.record_stream(self.side)
        return torch.cuda.stream(self.side)

    def join(self, t):
        if self.on:
            self.main.wait_stream(self.side)
            t.record_stream(self.main)
        return t


class SWGAN_unet(nn.Module):
    """styleUnet.py:1190-1410."""

    def __init__(self, inp_size, inp_ch, out_ch, out_size, style_dim, n_mlp, middle_size=8, c_dim=0, channel_multiplier=2,
                 blur_kernel=(1, 3, 3, 1), lr_mlp=0.01):
        super().__init__()
        self.inp_size, self.style_dim = inp_size, style_dim
        self.middle_log_size = int(math.log(middle_size, 2))
        self.style = _style_mlp(style_dim + c_dim, style_dim, n_mlp, lr_mlp)
        self.channels = _CHANNELS(channel_multiplier)
        self.log_size = int(math.log(out_size, 2)) - 1
        in_channel = self.channels[inp_size // 2]
        self.from_rgbs, self.cond_convs, self.comb_convs = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.comb_convs.append(ConvLayer(in_channel * 2, in_channel, 3))
        self.conv_in = ConvLayer(inp_ch, in_channel, 3, downsample=True)
        for i in range(int(math.log(inp_size, 2)) - 2, self.middle_log_size - 1, -1):
            out_channel = self.channels[2 ** i]
            self.from_rgbs.append(FromRGB(in_channel, inp_ch, downsample=True, use_wt=False))
            self.cond_convs.append(ConvBlock(in_channel, out_channel, blur_kernel))
            self.comb_convs.append(ConvLayer(out_channel * (2 if i > self.middle_log_size else 1), out_channel, 3))
            in_channel = out_channel
        self.convs, self.to_rgbs, self.noises = nn.ModuleList(), nn.ModuleList(), nn.Module()
        in_channel = self.channels[middle_size]
        self.num_layers = (self.log_size - self.middle_log_size) * 2
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 8) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(self.middle_log_size + 1, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(in_channel=out_channel, style_dim=style_dim, out_channel=out_ch * 4))
            in_channel = out_channel
        self.iwt = InverseHaarTransform(3)
        self.n_latent = self.log_size * 2 - (self.middle_log_size * 2 - 1) + 1

    def make_noise(self, device, zero_noise=False):
        f = torch.zeros if zero_noise else torch.randn
        return [f(1, 1, 2 ** i, 2 ** i, device=device) for i in range(self.middle_log_size + 1, self.log_size + 1) for _ in range(2)]

    def get_latent(self, x):
        return self.style(x)

    def forward(self, *args, **kwargs):
        """Gradients enabled -> the differentiable formulation (styleunet_train.py); otherwise the fused inference kernels."""
        if torch.is_grad_enabled():
            return self._forward(True, *args, **kwargs)
        with torch.no_grad():
            return self._forward(False, *args, **kwargs)

    def _forward(self, ag, styles, condition_img, cond=None, return_latents=False, inject_index=None, truncation=1,
                 truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True):
        if not input_is_latent:
            mlp = (lambda v: T.style_mlp(self.style, v)) if ag else self.style
            styles = [mlp(s if cond is None else torch.cat([s, cond], dim=-1)) for s in styles]
        if noise is None:
            noise = [None] * self.num_layers if randomize_noise else [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        latent = _latents(styles, self.n_latent, inject_index)
        if ag:
            return T.swgan_unet_forward(self, latent, condition_img, noise)
        if getattr(self, "_style_plan", None) is None:       # (module, latent index) in call order
            ent = []
            for k, (c1, c2, tr) in enumerate(zip(self.convs[::2], self.convs[1::2], self.to_rgbs)):
                ent += [(c1.conv, 2 * k), (c2.conv, 2 * k + 1), (tr.conv, 2 * k + 2)]
            self._style_plan = _StylePlan(ent)
        # every modulation / demodulation of the network: two launches, beside the condition encoder
        sd, feats = _plan_beside(self._style_plan, latent, lambda: _CondEncoder.run(self, condition_img))
        i, skip, out = 0, None, None
        fork = _SkipFork(condition_img.device)
        for conv1, conv2, n1, n2, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[::2], noise[1::2], self.to_rgbs):
            if i == 0:
                out = self.comb_convs[-1](feats[-1], out_cl=True)
            elif i < 2 * len(self.comb_convs):
                out = self.comb_convs[-1 - (i // 2)](torch.cat([out, feats[-1 - (i // 2)]], dim=-1), out_cl=True)
            k3 = 3 * (i // 2)
            out = conv1(out, latent[:, i], noise=n1, out_cl=True, sd=None if sd is None else sd[k3])
            out = conv2(out, latent[:, i + 1], noise=n2, out_cl=True, sd=None if sd is None else sd[k3 + 1])
            with fork.branch(out):
                skip = to_rgb(out, latent[:, i + 2], skip, sd=None if sd is None else sd[k3 + 2])      # 12-channel wavelet skip stays NCHW fp32
            i += 2
        return self.iwt(fork.join(skip))


class Discriminator(nn.Module):
    """styleUnet.py:470-562 (wavelet discriminator of stage two, train_avatarHD.py:112).  Forward only: the GAN losses need
    first- and second-order gradients through it (utils/styleUnet_util.py:65-79), which need the convolution backward."""

    def __init__(self, size, img_channel=6, channel_multiplier=2, blur_kernel=(1, 3, 3, 1), c_dim=0):
        super().__init__()
        if c_dim > 0:
            raise NotImplementedError("pose-conditioned discriminator (c_dim > 0) is not used by the reference scripts")
        channels = _CHANNELS(channel_multiplier)
        self.dwt = HaarTransform(img_channel)
        self.from_rgbs, self.convs = nn.ModuleList(), nn.ModuleList()
        log_size = int(math.log(size, 2)) - 1
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            self.from_rgbs.append(FromRGB(in_channel, img_channel, downsample=i != log_size))
            self.convs.append(ConvBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.from_rgbs.append(FromRGB(channels[4], img_channel))
        self.stddev_group, self.stddev_feat = 4, 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(EqualLinear(channels[4] * 4 * 4, channels[4], activation="fused_lrelu"),
                                          EqualLinear(channels[4], 1))
        self.c_dim = c_dim

    def forward(self, x, flat_pose=None):
        if torch.is_grad_enabled():     # GAN losses / R1: first- and second-order autograd (utils/styleUnet_util.py:65-79)
            return T.discriminator_forward(self, x)
        with torch.no_grad():
            return self._forward_inference(x)

    def _forward_inference(self, x):
        x = self.dwt(x)
        out = None
        for from_rgb, block in zip(self.from_rgbs, self.convs):
            x, out = from_rgb(x, out, out_cl=True)
            out = block(out, out_cl=True)
        _, out = self.from_rgbs[-1](x, out, out_cl=True)
        out = hconv.to_nchw(out)                                  # 4x4 map: the minibatch statistics run in torch
        b, c, h, w = out.shape
        group = min(b, self.stddev_group)                                  # minibatch standard deviation (:539-545)
        sd = out.view(group, -1, self.stddev_feat, c // self.stddev_feat, h, w)
        sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8).mean([2, 3, 4], keepdims=True).squeeze(2)
        out = torch.cat([out, sd.repeat(group, 1, h, w)], 1).contiguous()
        out = self.final_conv(out)
        return self.final_linear(out.view(b, -1))


class StyleGAN_zxc(nn.Module):
    """styleUnet.py:631-878, the configuration the plane generators use (model/nerf_model.py:39-42): condition-image
    encoder (inp_size > 0) and no_skip=True (1x1 conv_out instead of the ToRGB pyramid)."""

    def __init__(self, out_ch, out_size, style_dim, mlp_dim=32, n_mlp=0, middle_size=8, inject_layers=(), zero_latent=False,
                 zero_noise=False, no_skip=False, channel_multiplier=2, blur_kernel=(1, 3, 3, 1), lr_mlp=0.01, n_latent=None,
                 inp_size=0, inp_ch=0, pass_kernel=False):
        super().__init__()
        if not (inp_size > 0 and no_skip):
            raise NotImplementedError("only the plane-generator configuration (inp_size > 0, no_skip=True) is built")
        self.no_skip, self.style_dim = no_skip, mlp_dim
        self.middle_log_size = int(math.log(middle_size, 2))
        self.cond_img_enc, self.n_mlp = True, n_mlp
        if n_mlp > 0:
            self.style = _style_mlp(style_dim, mlp_dim, n_mlp, lr_mlp)
        self.channels = _CHANNELS(channel_multiplier)
        self.log_size = int(math.log(out_size, 2))
        in_channel = self.channels[inp_size // 2]
        self.from_rgbs, self.cond_convs, self.comb_convs = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.comb_convs.append(ConvLayer(in_channel * 2, in_channel, 3))
        self.conv_in = ConvLayer(inp_ch, in_channel, 3, downsample=True)
        for i in range(int(math.log(inp_size, 2)) - 2, self.middle_log_size, -1):
            out_channel = self.channels[2 ** i]
            self.from_rgbs.append(FromRGB(in_channel, inp_ch, downsample=True, use_wt=False))
            self.cond_convs.append(ConvBlock(in_channel, out_channel, blur_kernel))
            self.comb_convs.append(ConvLayer(out_channel * 2, out_channel, 3))
            in_channel = out_channel
        self.convs, self.to_rgbs, self.noises = nn.ModuleList(), nn.ModuleList(), nn.Module()
        self.input = ConstantInput(self.channels[middle_size], size=middle_size)
        self.conv1 = StyledConv(self.channels[middle_size], self.channels[middle_size], 3, self.style_dim, blur_kernel=blur_kernel)
        self.conv_out = ConvLayer(self.channels[out_size], out_ch, 1)
        in_channel = self.channels[middle_size]
        self.num_layers = (self.log_size - self.middle_log_size) * 2 + 1
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 8) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(self.middle_log_size + 1, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, self.style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, self.style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(None)
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - (self.middle_log_size * 2 - 1) + 1 if n_latent is None else n_latent
        # zero_noise: every layer's noise is zero except the first, a fixed draw that is NOT in the state_dict (:746-751)
        self.zero_noise = self.make_noise(zero_noise=True) if zero_noise else None
        if zero_latent:
            self.register_buffer("zero_latents", torch.zeros(1, self.n_latent, self.style_dim))
        else:
            self.zero_latents = None

    def make_noise(self, zero_noise=False):
        f = torch.zeros if zero_noise else torch.randn
        noises = [torch.randn(1, 1, 2 ** self.middle_log_size, 2 ** self.middle_log_size)]
        for i in range(self.middle_log_size + 1, self.log_size + 1):
            noises += [f(1, 1, 2 ** i, 2 ** i), f(1, 1, 2 ** i, 2 ** i)]
        return noises

    def forward(self, *args, **kwargs):
        if torch.is_grad_enabled():
            return self._forward(True, *args, **kwargs)
        with torch.no_grad():
            return self._forward(False, *args, **kwargs)

    def _forward(self, ag, styles, cond_feats, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                 input_is_latent=False, noise=None, randomize_noise=True, **kwargs):
        dev = cond_feats.device
        batch = cond_feats.shape[0]
        if self.zero_latents is None:
            if not input_is_latent:
                styles = [T.style_mlp(self.style, s) if ag else self.style(s) for s in styles]
            if truncation < 1:
                styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
            latent = _latents(styles, self.n_latent, inject_index)
        else:
            latent = self.zero_latents.expand(batch, -1, -1)
        if self.zero_noise is None:
            if noise is None:
                noise = [None] * self.num_layers if randomize_noise else [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        else:
            if any(n.device != dev for n in self.zero_noise):
                self.zero_noise = [n.to(dev) for n in self.zero_noise]
            noise = self.zero_noise
        if ag:
            image = T.stylegan_zxc_forward(self, latent, cond_feats, noise)
            return (image, latent) if return_latents else (image, None)
        if getattr(self, "_style_plan", None) is None:       # (module, latent index) in call order
            self._style_plan = _StylePlan([(self.conv1.conv, 0)] + [(c.conv, k + 1) for k, c in enumerate(self.convs)])
        sd, feats = _plan_beside(self._style_plan, latent, lambda: _CondEncoder.run(self, cond_feats))
        out = self.conv1(self.input(latent), latent[:, 0], noise=noise[0], out_cl=True, sd=None if sd is None else sd[0])
        i = 1
        for conv1, conv2, n1, n2 in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2]):
            if 1 < i <= 2 * len(feats) + 1:
                out = self.comb_convs[-(i // 2)](torch.cat([out, feats[-(i // 2)]], dim=-1), out_cl=True)
            out = conv1(out, latent[:, i], noise=n1, out_cl=True, sd=None if sd is None else sd[i])
            out = conv2(out, latent[:, i + 1], noise=n2, out_cl=True, sd=None if sd is None else sd[i + 1])
            i += 2
        image = self.conv_out(out)                               # back to the reference layout (NCHW fp32)
        return (image, latent) if return_latents else (image, None)
