// HBM-bound element-wise / FIR operators behind the reference's `model/op` extension boundary,
// plus ray generation.  See include/havatar_b200.h for the contracts and reference citations.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/havatar_b200.h"
#include "render_common.cuh"

namespace hav {

constexpr int kSMs = 148;  // B200: grids are sized in multiples of the SM count

// ------------------------------------------------------------------------------------------------
// fused bias + activation  (model/op/fused_bias_act_kernel.cu:18-65)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bias_act_one(float x, float ref, int mode, float alpha, float scale) {
  float y;
  switch (mode) {
    case 30: y = (x > 0.0f) ? x : x * alpha; break;
    case 31: y = (ref > 0.0f) ? x : x * alpha; break;
    case 12:
    case 32: y = 0.0f; break;
    default: y = x; break;  // 10, 11 and anything else: linear (kernel.cu:41-49)
  }
  return y * scale;
}

// Vector path: 4 consecutive elements share one bias channel when step_b % 4 == 0.
__global__ void __launch_bounds__(256) bias_act_vec4_kernel(float4 *__restrict__ out, const float4 *__restrict__ x,
                                                            const float *__restrict__ bias,
                                                            const float4 *__restrict__ ref, int64_t n4, int64_t step_b4,
                                                            int64_t size_b, int mode, float alpha, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldcs(x + i);
    if (bias != nullptr) {
      float b = __ldg(bias + (i / step_b4) % size_b);
      v.x += b, v.y += b, v.z += b, v.w += b;
    }
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ref != nullptr) r = __ldcs(ref + i);
    float4 y;
    y.x = bias_act_one(v.x, r.x, mode, alpha, scale);
    y.y = bias_act_one(v.y, r.y, mode, alpha, scale);
    y.z = bias_act_one(v.z, r.z, mode, alpha, scale);
    y.w = bias_act_one(v.w, r.w, mode, alpha, scale);
    out[i] = y;
  }
}

__global__ void __launch_bounds__(256) bias_act_scalar_kernel(float *__restrict__ out, const float *__restrict__ x,
                                                              const float *__restrict__ bias,
                                                              const float *__restrict__ ref, int64_t n, int64_t step_b,
                                                              int64_t size_b, int mode, float alpha, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (bias != nullptr) v += __ldg(bias + (i / step_b) % size_b);
    float r = ref != nullptr ? ref[i] : 0.0f;
    out[i] = bias_act_one(v, r, mode, alpha, scale);
  }
}

// StyledConv's tail in one pass (model/styleUnet.py:596-598: NoiseInjection `image + weight * noise`, :300-310, then
// FusedLeakyReLU): y = lrelu(x + nw * noise[b or 0, pixel] + bias[c]) * scale for x [B,C,inner]; nw is read from device memory
// (it is a parameter).  One thread per 4 consecutive pixels of one (b, c) plane when inner % 4 == 0.
template <bool kVec>
__global__ void __launch_bounds__(256) noise_bias_act_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ bias,
                                                             const float *__restrict__ noise, const float *__restrict__ nw_ptr, long n_units,
                                                             long inner, int C, long noise_bstride, float alpha, float scale) {
  const float nw = __ldg(nw_ptr);
  constexpr int V = kVec ? 4 : 1;
  const long units_per_plane = inner / V;
  for (long u = (long)blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += (long)gridDim.x * blockDim.x) {
    const long plane = u / units_per_plane, i = (u - plane * units_per_plane) * V;
    const int c = (int)(plane % C);
    const long b = plane / C;
    const float bv = bias != nullptr ? __ldg(bias + c) : 0.0f;
    if (kVec) {
      const float4 xv = __ldcs(reinterpret_cast<const float4 *>(x + plane * inner + i));
      const float4 nv = __ldg(reinterpret_cast<const float4 *>(noise + b * noise_bstride + i));
      float4 y;
      y.x = xv.x + nw * nv.x + bv, y.y = xv.y + nw * nv.y + bv, y.z = xv.z + nw * nv.z + bv, y.w = xv.w + nw * nv.w + bv;
      y.x = (y.x > 0.0f ? y.x : y.x * alpha) * scale, y.y = (y.y > 0.0f ? y.y : y.y * alpha) * scale;
      y.z = (y.z > 0.0f ? y.z : y.z * alpha) * scale, y.w = (y.w > 0.0f ? y.w : y.w * alpha) * scale;
      *reinterpret_cast<float4 *>(out + plane * inner + i) = y;
    } else {
      float y = x[plane * inner + i] + nw * __ldg(noise + b * noise_bstride + i) + bv;
      out[plane * inner + i] = (y > 0.0f ? y : y * alpha) * scale;
    }
  }
}

// FusedLeakyReLUFunctionBackward in one pass (model/op/fused_act.py:23-47: the gated gradient, then grad_input.sum over every
// dim but the channel): grad_input = (ref > 0 ? g : g * alpha) * scale and, per (split, channel), the partial sum of grad_input
// over the split's share of the channel's B x inner elements.  The caller adds the `splits` partials (a [splits,C] tensor):
// fixed summation order, no atomics, and the second full read of grad_input by a separate reduction disappears.
// With noise != nullptr a second table npart[s, c] collects sum(grad_input * noise): the gradient of NoiseInjection's weight is
// its sum over (s, c)  (d/dw of x + w * noise, the reference gets it from two more full-tensor reductions).
template <bool kVec>
__global__ void __launch_bounds__(256) bias_act_backward_kernel(float *__restrict__ gin, float *__restrict__ partials,
                                                                const float *__restrict__ g, const float *__restrict__ ref, int B, int C,
                                                                long inner, long per, float alpha, float scale,
                                                                const float *__restrict__ noise = nullptr, long noise_bstride = 0,
                                                                float *__restrict__ npart = nullptr) {
  const int c = blockIdx.y, s = blockIdx.x;
  const long total = (long)B * inner, lo = (long)s * per, hi = lo + per < total ? lo + per : total;
  float acc = 0.0f, nacc = 0.0f;
  for (int b = (int)(lo / inner); b < B && (long)b * inner < hi; ++b) {
    const long seg0 = (long)b * inner, i0 = (lo > seg0 ? lo : seg0) - seg0, i1 = (hi < seg0 + inner ? hi : seg0 + inner) - seg0;
    const size_t base = ((size_t)b * C + c) * inner;
    if (kVec) {     // inner % 4 == 0 and per % 4 == 0: aligned float4 runs that never straddle a plane
      const float4 *g4 = reinterpret_cast<const float4 *>(g + base), *r4 = reinterpret_cast<const float4 *>(ref + base);
      float4 *o4 = reinterpret_cast<float4 *>(gin + base);
      for (long i = i0 / 4 + threadIdx.x; i < i1 / 4; i += blockDim.x) {
        const float4 gv = __ldcs(g4 + i), rv = __ldcs(r4 + i);
        float4 y;
        y.x = (rv.x > 0.0f ? gv.x : gv.x * alpha) * scale, y.y = (rv.y > 0.0f ? gv.y : gv.y * alpha) * scale;
        y.z = (rv.z > 0.0f ? gv.z : gv.z * alpha) * scale, y.w = (rv.w > 0.0f ? gv.w : gv.w * alpha) * scale;
        o4[i] = y;
        acc += (y.x + y.y) + (y.z + y.w);
        if (noise != nullptr) {
          const float4 nv = __ldg(reinterpret_cast<const float4 *>(noise + b * noise_bstride) + i);
          nacc += (y.x * nv.x + y.y * nv.y) + (y.z * nv.z + y.w * nv.w);
        }
      }
    } else {
      for (long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const float gv = g[base + i], y = (ref[base + i] > 0.0f ? gv : gv * alpha) * scale;
        gin[base + i] = y;
        acc += y;
        if (noise != nullptr) nacc += y * __ldg(noise + b * noise_bstride + i);
      }
    }
  }
  __shared__ float red[16];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o), nacc += __shfl_xor_sync(0xffffffffu, nacc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc, red[8 + (threadIdx.x >> 5)] = nacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f, tn = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[w], tn += red[8 + w];
    partials[(size_t)s * C + c] = t;
    if (npart != nullptr) npart[(size_t)s * C + c] = tn;
  }
}

// ------------------------------------------------------------------------------------------------
// upfirdn2d  (model/op/upfirdn2d_kernel.cu:49-207; executable spec model/op/upfirdn2d.py:172-213)
//   out[oy,ox] = sum_{ky,kx} U[oy*dy + ky - py0, ox*dx + kx - px0] * K[kh-1-ky, kw-1-kx]
//   U[Y,X] = x[Y/uy, X/ux] when Y%uy == 0 && X%ux == 0 and inside the input, else 0.
// One CTA computes a 32x64 output tile of one (major, minor) image from an input window staged in
// shared memory, so every input element is read from HBM/L2 once per tile.
// ------------------------------------------------------------------------------------------------
constexpr int kTileH = 32, kTileW = 64;
constexpr int kMaxTaps = 24;  // per dimension

struct UfdParams {
  int in_h, in_w, minor, out_h, out_w, kh, kw;
  int up_x, up_y, down_x, down_y, pad_x0, pad_y0;
  int win_h, win_w;  // smem window size (input coords)
};

__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__device__ __forceinline__ int ceil_div(int a, int b) { return -floor_div(-a, b); }

__global__ void __launch_bounds__(256) upfirdn2d_kernel(float *__restrict__ out, const float *__restrict__ x,
                                                        const float *__restrict__ kernel, UfdParams p, int tiles_x) {
  extern __shared__ float sm[];
  float *sk = sm;                          // flipped kernel [kh][kw]
  float *sw = sm + kMaxTaps * kMaxTaps;    // input window [win_h][win_w]
  const int img = blockIdx.y;              // major * minor + m
  const int major_i = img / p.minor, m = img % p.minor;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x % tiles_x;
  const int oy0 = tile_y * kTileH, ox0 = tile_x * kTileW;
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    int ky = i / p.kw, kx = i % p.kw;
    sk[i] = __ldg(kernel + (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx));
  }
  // first input row/col any output of this tile can touch: ceil((o0*d - pad) / up)
  const int iy0 = ceil_div(oy0 * p.down_y - p.pad_y0, p.up_y);
  const int ix0 = ceil_div(ox0 * p.down_x - p.pad_x0, p.up_x);
  const float *xin = x + (size_t)major_i * p.in_h * p.in_w * p.minor + m;
  for (int i = threadIdx.x; i < p.win_h * p.win_w; i += blockDim.x) {
    int wy = i / p.win_w, wx = i % p.win_w;
    int iy = iy0 + wy, ix = ix0 + wx;
    float v = 0.0f;
    if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v = __ldg(xin + ((size_t)iy * p.in_w + ix) * p.minor);
    sw[i] = v;
  }
  __syncthreads();
  float *oimg = out + (size_t)major_i * p.out_h * p.out_w * p.minor + m;
  for (int i = threadIdx.x; i < kTileH * kTileW; i += blockDim.x) {
    int oy = oy0 + i / kTileW, ox = ox0 + i % kTileW;
    if (oy >= p.out_h || ox >= p.out_w) continue;
    const int Y0 = oy * p.down_y - p.pad_y0, X0 = ox * p.down_x - p.pad_x0;
    // first tap with (Y0 + ky) % up == 0
    int ky_first = ((-Y0) % p.up_y + p.up_y) % p.up_y;
    int kx_first = ((-X0) % p.up_x + p.up_x) % p.up_x;
    float acc = 0.0f;
    for (int ky = ky_first; ky < p.kh; ky += p.up_y) {
      const int wy = (Y0 + ky) / p.up_y - iy0;  // exact: Y0 + ky is a multiple of up_y
      if (wy < 0 || wy >= p.win_h) continue;
      for (int kx = kx_first; kx < p.kw; kx += p.up_x) {
        const int wx = (X0 + kx) / p.up_x - ix0;
        if (wx < 0 || wx >= p.win_w) continue;
        acc = fmaf(sw[wy * p.win_w + wx], sk[ky * p.kw + kx], acc);
      }
    }
    oimg[((size_t)oy * p.out_w + ox) * p.minor] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Specialised upfirdn2d for the shapes the StyleUNet actually uses (minor == 1, square up/down):
//   <1,1,4,4> Blur, <1,2,4,4> Downsample, <2,1,4,4> Upsample, <1,2,2,2> Haar analysis, <2,1,2,2> Haar synthesis.
// Compile-time taps (fully unrolled), compile-time window pitch (no div/mod by runtime values).  up == down == 1: one thread = one
// output column x 8 rows with every window row held in registers across the output rows it overlaps; otherwise: a strip of 4
// horizontally adjacent outputs that shares its window reads, 128-bit stores when the row pitch allows.
// Same arithmetic as the generic kernel: out = sum over the taps that land on a real input sample.
// ------------------------------------------------------------------------------------------------
template <int UP, int DOWN, int KH, int KW>
struct UfdFast {
  static constexpr int kWinH = ((kTileH - 1) * DOWN + KH - 1) / UP + 2;
  static constexpr int kWinW = ((kTileW - 1) * DOWN + KW - 1) / UP + 2;
  static constexpr int kSmemFloats = KH * KW + kWinH * kWinW;
};

template <int UP, int DOWN, int KH, int KW>
__global__ void __launch_bounds__(256) upfirdn2d_fast_kernel(float *__restrict__ out, const float *__restrict__ x,
                                                             const float *__restrict__ kernel, UfdParams p, int tiles_x) {
  using F = UfdFast<UP, DOWN, KH, KW>;
  __shared__ float sk[KH * KW];
  __shared__ float sw[F::kWinH * F::kWinW];
  const int img = blockIdx.y;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
  const int oy0 = tile_y * kTileH, ox0 = tile_x * kTileW;
  if (threadIdx.x < KH * KW) {
    const int ky = threadIdx.x / KW, kx = threadIdx.x % KW;
    sk[threadIdx.x] = __ldg(kernel + (KH - 1 - ky) * KW + (KW - 1 - kx));
  }
  const int iy0 = ceil_div(oy0 * DOWN - p.pad_y0, UP), ix0 = ceil_div(ox0 * DOWN - p.pad_x0, UP);
  const float *xin = x + (size_t)img * p.in_h * p.in_w;
  if (UP == 1 && DOWN == 1) {
    // window rows are dealt to the 8 warps, a lane walks its row 32 columns at a time: coalesced, no div / mod
    for (int wy = threadIdx.x >> 5; wy < F::kWinH; wy += 8) {
      const int iy = iy0 + wy;
      const bool row_ok = iy >= 0 && iy < p.in_h;
      const float *xrow = xin + (size_t)(row_ok ? iy : 0) * p.in_w;
      for (int wx = threadIdx.x & 31; wx < F::kWinW; wx += 32) {
        const int ix = ix0 + wx;
        sw[wy * F::kWinW + wx] = (row_ok && ix >= 0 && ix < p.in_w) ? __ldg(xrow + ix) : 0.0f;
      }
    }
  } else {      // wide (stride-2) or narrow (up-sampling) windows: the flat loop keeps every thread busy -- measured faster there
    for (int i = threadIdx.x; i < F::kWinH * F::kWinW; i += 256) {
      const int wy = i / F::kWinW, wx = i - wy * F::kWinW;
      const int iy = iy0 + wy, ix = ix0 + wx;
      sw[i] = (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) ? __ldg(xin + (size_t)iy * p.in_w + ix) : 0.0f;
    }
  }
  __syncthreads();
  float kreg[KH * KW];
#pragma unroll
  for (int i = 0; i < KH * KW; ++i) kreg[i] = sk[i];
  float *oimg = out + (size_t)img * p.out_h * p.out_w;
  if (UP == 1 && DOWN == 1) {
    // one thread = one output column x 8 rows: lanes run along x (conflict-free window reads for DOWN == 1, full-line
    // stores), and a window row loaded once into registers feeds every output row of the strip it overlaps
    constexpr int kStrip = 8, NR = (kStrip - 1) * DOWN + KH;
    const int tx = threadIdx.x & (kTileW - 1), ty = threadIdx.x / kTileW;      // 64 columns x 4 strips = 256 threads
    const int ox = ox0 + tx, oyb = oy0 + ty * kStrip;
    if (ox < p.out_w && oyb < p.out_h) {
      const int wy0 = oyb * DOWN - p.pad_y0 - iy0, wx0 = ox * DOWN - p.pad_x0 - ix0;
      float acc[kStrip];
#pragma unroll
      for (int j = 0; j < kStrip; ++j) acc[j] = 0.0f;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        float row[KW];
#pragma unroll
        for (int c = 0; c < KW; ++c) row[c] = sw[(wy0 + r) * F::kWinW + wx0 + c];
#pragma unroll
        for (int j = 0; j < kStrip; ++j) {
          const int ky = r - j * DOWN;               // compile-time after unrolling
          if (ky >= 0 && ky < KH) {
#pragma unroll
            for (int kx = 0; kx < KW; ++kx) acc[j] = fmaf(row[kx], kreg[ky * KW + kx], acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kStrip; ++j)
        if (oyb + j < p.out_h) oimg[(size_t)(oyb + j) * p.out_w + ox] = acc[j];
    }
    return;
  }
  const bool vec_ok = (p.out_w & 3) == 0;
  // strips of 4 horizontally adjacent outputs: (kTileH * kTileW / 4) strips per tile, 2 per thread
  for (int sidx = threadIdx.x; sidx < kTileH * kTileW / 4; sidx += 256) {
    const int ty = sidx / (kTileW / 4), oy = oy0 + ty, ox = ox0 + (sidx - ty * (kTileW / 4)) * 4;
    if (oy >= p.out_h || ox >= p.out_w) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int Y0 = oy * DOWN - p.pad_y0;
    if (UP == 1) {      // DOWN == 2 (the 2x2 Haar analysis and the 4x4 downsampler): measured faster in this form
      const int wy0 = Y0 - iy0, wx0 = ox * DOWN - p.pad_x0 - ix0;
      constexpr int NX = 3 * DOWN + KW;   // window columns a strip touches
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
        float row[NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) row[c] = sw[(wy0 + ky) * F::kWinW + wx0 + c];
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int kx = 0; kx < KW; ++kx) acc[o] = fmaf(row[o * DOWN + kx], kreg[ky * KW + kx], acc[o]);
      }
    } else {
      // UP == 2, DOWN == 1: tap ky contributes when (Y0 + ky) is even; KH / 2 taps per dimension
      const int ky_first = Y0 & 1;                       // (-Y0) mod 2
#pragma unroll
      for (int a = 0; a < KH / 2; ++a) {
        const int ky = ky_first + 2 * a;
        const int wy = ((Y0 + ky) >> 1) - iy0;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int X0 = (ox + o) - p.pad_x0, kx_first = X0 & 1;
#pragma unroll
          for (int c = 0; c < KW / 2; ++c) {
            const int kx = kx_first + 2 * c;
            acc[o] = fmaf(sw[wy * F::kWinW + ((X0 + kx) >> 1) - ix0], kreg[ky * KW + kx], acc[o]);
          }
        }
      }
    }
    float *op = oimg + (size_t)oy * p.out_w + ox;
    if (vec_ok && ox + 3 < p.out_w) {
      *reinterpret_cast<float4 *>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (ox + o < p.out_w) op[o] = acc[o];
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Streaming 4x4 FIR (up == down == 1, minor == 1: the Blur after every transposed / before every stride-2 convolution,
// model/styleUnet.py:69-87).  No shared memory and no tile-wide synchronisation: a thread owns 4 adjacent output columns and
// marches down kBlurRows output rows, keeping the last three rows of horizontal sums in registers.  Each input row is read
// with three 16-byte loads from 16-byte aligned addresses (the row pitch is 2^k + 1 floats here, so the alignment of a row
// shifts by one element per row: the shift is uniform across the warp and selects one of four statically indexed code paths).
// The 4x4 tap matrix is factorised in the kernel: rank one ([1,3,3,1] x [1,3,3,1], every tap set the StyleUNet uses) takes
// the separable path (4 + 4 FMAs per output instead of 16); anything else the general path.
// ------------------------------------------------------------------------------------------------
constexpr int kBlurRows = 32;

template <int S, bool kSep>
__device__ __forceinline__ void blur_row(const float (&c)[12], const float (&kx)[4], const float (&k2)[16], float (&h)[4],
                                         float (&part)[4][4]) {
  if (kSep) {
#pragma unroll
    for (int o = 0; o < 4; ++o) h[o] = fmaf(c[S + o + 3], kx[3], fmaf(c[S + o + 2], kx[2], fmaf(c[S + o + 1], kx[1], c[S + o] * kx[0])));
  } else {
    // general taps: this input row is tap row i of the output row that is i rows above it; part[i] collects that output row
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][o] = fmaf(c[S + o + j], k2[i * 4 + j], part[i][o]);
  }
}

// kFast: every 16-byte chunk the strip touches lies inside the tensor's allocation (no pointer guards; rows above / below the
// image are a CTA-uniform branch, and only the first / last thread of a row masks its out-of-row columns).  False only at the
// two ends of the whole tensor.  Rows are processed four at a time so that the window rotates by renaming.
template <bool kSep, bool kFast>
__device__ __forceinline__ void blur_march(float *__restrict__ oimg, const float *__restrict__ ximg, const float *__restrict__ xlo,
                                           const float *__restrict__ xhi, const UfdParams &p, int ox, int oy0, int rows,
                                           const float (&kx)[4], const float (&ky)[4], const float (&k2)[16]) {
  // sliding state: separable -> the horizontal sums of the last three input rows; general -> partially summed output rows
  float H[4][4], part[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int o = 0; o < 4; ++o) H[i][o] = 0.0f, part[i][o] = 0.0f;
  const int ix0 = ox - p.pad_x0;                       // input column of tap 0 of output ox
  const bool edge = ix0 < 0 || ix0 + 7 > p.in_w;       // some of the 7 columns this thread reads lie outside the row
  const bool pair_ok = (p.out_w & 1) == 0 && ox + 3 < p.out_w;   // 8-byte aligned output pairs on every row
  const float *rowp_base = ximg + (long)(oy0 - p.pad_y0) * p.in_w + ix0;   // first element this thread needs on input row r = 0
  const int total = rows + 3;
  float *op = oimg + (size_t)oy0 * p.out_w + ox - (size_t)3 * p.out_w;   // output row r - 3

  // the three 16-byte chunks of input row r (zeros when the row is above / below the image), loaded one row AHEAD of their use
  auto load_row = [&](int r, float4 (&dst)[3], int &sh_out) {
    const float *rp = rowp_base + (long)r * p.in_w;
    const int sh = (int)(((uintptr_t)rp >> 2) & 3);
    sh_out = sh;
    const float4 *a0 = reinterpret_cast<const float4 *>(rp - sh);
    const int iy = oy0 + r - p.pad_y0;
#pragma unroll
    for (int q = 0; q < 3; ++q) dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < p.in_h) {          // uniform across the CTA: rows above / below the image are zero padding
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (kFast) {
          dst[q] = __ldg(a0 + q);
        } else {
          const float *ap = reinterpret_cast<const float *>(a0 + q);
          if (ap >= xlo && ap + 4 <= xhi) dst[q] = __ldg(a0 + q);             // inside the tensor's allocation (16-byte aligned)
          else {
            if (ap + 0 >= xlo && ap + 0 < xhi) dst[q].x = __ldg(ap + 0);
            if (ap + 1 >= xlo && ap + 1 < xhi) dst[q].y = __ldg(ap + 1);
            if (ap + 2 >= xlo && ap + 2 < xhi) dst[q].z = __ldg(ap + 2);
            if (ap + 3 >= xlo && ap + 3 < xhi) dst[q].w = __ldg(ap + 3);
          }
        }
      }
    }
  };
  float4 nxt[3];
  int sh_nxt;
  load_row(0, nxt, sh_nxt);
  auto one_row = [&](int r, float (&Hm3)[4], float (&Hm2)[4], float (&Hm1)[4], float (&Hnew)[4]) {
    float c[12];
    const int sh = sh_nxt;
#pragma unroll
    for (int q = 0; q < 3; ++q) c[4 * q] = nxt[q].x, c[4 * q + 1] = nxt[q].y, c[4 * q + 2] = nxt[q].z, c[4 * q + 3] = nxt[q].w;
    if (r + 1 < total) load_row(r + 1, nxt, sh_nxt);      // in flight while this row is filtered
    if (edge) {   // first / last thread of a row only: zero padding left / right of the row (a short divergent block)
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        const int ix = ix0 - sh + e;
        if (ix < 0 || ix >= p.in_w) c[e] = 0.0f;
      }
    }
    switch (sh) {        // warp-uniform (every lane's first element is 4 lanes * 16 bytes further along the same row)
      case 0: blur_row<0, kSep>(c, kx, k2, Hnew, part); break;
      case 1: blur_row<1, kSep>(c, kx, k2, Hnew, part); break;
      case 2: blur_row<2, kSep>(c, kx, k2, Hnew, part); break;
      default: blur_row<3, kSep>(c, kx, k2, Hnew, part); break;
    }
    float o4[4];
    if (kSep) {
#pragma unroll
      for (int o = 0; o < 4; ++o) o4[o] = fmaf(Hnew[o], ky[3], fmaf(Hm1[o], ky[2], fmaf(Hm2[o], ky[1], Hm3[o] * ky[0])));
    } else {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        o4[o] = part[3][o];                                   // the output row whose LAST tap row this input row is
        part[3][o] = part[2][o], part[2][o] = part[1][o], part[1][o] = part[0][o], part[0][o] = 0.0f;
      }
    }
    if (r >= 3) {
      if (pair_ok) {
        *reinterpret_cast<float2 *>(op) = make_float2(o4[0], o4[1]);
        *reinterpret_cast<float2 *>(op + 2) = make_float2(o4[2], o4[3]);
      } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (ox + o < p.out_w) op[o] = o4[o];
      }
    }
    op += p.out_w;
  };
  for (int r = 0; r < total; r += 4) {      // the window rotates through H[0..3] by renaming
    one_row(r, H[1], H[2], H[3], H[0]);
    if (r + 1 < total) one_row(r + 1, H[2], H[3], H[0], H[1]);
    if (r + 2 < total) one_row(r + 2, H[3], H[0], H[1], H[2]);
    if (r + 3 < total) one_row(r + 3, H[0], H[1], H[2], H[3]);
  }
}

__global__ void __launch_bounds__(128, 6) blur4x4_stream_kernel(float *__restrict__ out, const float *__restrict__ x,
                                                             const float *__restrict__ kernel, UfdParams p, int bands_x, long n_img) {
  // flipped taps (upfirdn2d correlates with the flipped kernel) and their rank-one factorisation k2[i][j] = ky[i] * kx[j]
  float k2[16], kx[4], ky[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) k2[i] = __ldg(kernel + 15 - i);
  int best = 0;
#pragma unroll
  for (int i = 1; i < 16; ++i)
    if (fabsf(k2[i]) > fabsf(k2[best])) best = i;
  const int r0 = best >> 2, c0 = best & 3;
  const float piv = k2[best];
  bool sep = piv != 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) kx[j] = k2[r0 * 4 + j];
#pragma unroll
  for (int i = 0; i < 4; ++i) ky[i] = sep ? k2[i * 4 + c0] / piv : 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) sep = sep && fabsf(k2[i] - ky[i >> 2] * kx[i & 3]) <= 1e-6f * fabsf(piv);
  const int band = blockIdx.x % bands_x, strip = blockIdx.x / bands_x;
  const int ox = (band * blockDim.x + threadIdx.x) * 4, oy0 = strip * kBlurRows;
  if (ox >= p.out_w) return;
  const float *xhi = x + (size_t)n_img * p.in_h * p.in_w;
  for (long img = blockIdx.y; img < n_img; img += gridDim.y) {
    float *oimg = out + (size_t)img * p.out_h * p.out_w;
    const float *ximg = x + (size_t)img * p.in_h * p.in_w;
    const int rows = min(kBlurRows, p.out_h - oy0);
    // fast path: all rows of the strip exist, the thread's 7 columns lie inside the row, and the aligned 16-byte chunks around
    // them stay inside the allocation (they may spill into the neighbouring rows of the same tensor, which is harmless)
    const int iy_first = oy0 - p.pad_y0, ix0 = ox - p.pad_x0;
    const int iy_lo = max(iy_first, 0), iy_hi = min(iy_first + rows + 2, p.in_h - 1);       // rows that are actually read
    const float *first = ximg + (long)iy_lo * p.in_w + ix0, *last = ximg + (long)iy_hi * p.in_w + ix0;
    const bool fast = first - 3 >= x && last + 12 <= xhi;
    if (sep) {
      if (fast) blur_march<true, true>(oimg, ximg, x, xhi, p, ox, oy0, rows, kx, ky, k2);
      else blur_march<true, false>(oimg, ximg, x, xhi, p, ox, oy0, rows, kx, ky, k2);
    } else {
      blur_march<false, false>(oimg, ximg, x, xhi, p, ox, oy0, rows, kx, ky, k2);
    }
  }
}

static cudaError_t launch_blur4x4_stream(float *out, const float *x, const float *kernel, const UfdParams &p, int64_t imgs,
                                         cudaStream_t st) {
  // a thread owns 4 output columns: as many warps per CTA as a row needs (at most 4), so that narrow images do not park idle
  // warps on the SM
  const int warps = (p.out_w + 127) / 128;
  const int threads = 32 * (warps < 4 ? warps : 4);
  const int bands_x = (p.out_w + threads * 4 - 1) / (threads * 4), strips = (p.out_h + kBlurRows - 1) / kBlurRows;
  const int gy = (int)(imgs < 65535 ? imgs : 65535);
  blur4x4_stream_kernel<<<dim3(bands_x * strips, gy), threads, 0, st>>>(out, x, kernel, p, bands_x, (long)imgs);
  return cudaGetLastError();
}

template <int UP, int DOWN, int KH, int KW>
static cudaError_t launch_ufd_fast(float *out, const float *x, const float *kernel, const UfdParams &p, int64_t imgs, cudaStream_t st) {
  const int tiles_x = (p.out_w + kTileW - 1) / kTileW, tiles_y = (p.out_h + kTileH - 1) / kTileH;
  for (int64_t i0 = 0; i0 < imgs; i0 += 65535) {
    const int ny = (int)((imgs - i0) < 65535 ? (imgs - i0) : 65535);
    upfirdn2d_fast_kernel<UP, DOWN, KH, KW><<<dim3(tiles_x * tiles_y, ny), 256, 0, st>>>(
        out + (size_t)i0 * p.out_h * p.out_w, x + (size_t)i0 * p.in_h * p.in_w, kernel, p, tiles_x);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// upfirdn2d on channels-last fp16 tensors [B,H,W,C] (the StyleUNet's internal hand-over layout), square up / down
// factors, fp32 accumulation, with the StyledConv tail fused in: out = act(fir(x) + noise_weight * noise + bias[c]).
// One thread = one output pixel x 8 channels (one 16-byte load per tap, one 16-byte store).
// ------------------------------------------------------------------------------------------------
struct UfdClParams {
  int B, in_h, in_w, C, out_h, out_w, kh, kw, up, down, pad_x0, pad_y0, act, noise_bstride;
  float noise_weight;
};

__global__ void __launch_bounds__(256) upfirdn2d_cl_kernel(uint16_t *__restrict__ out, const uint16_t *__restrict__ x,
                                                           const float *__restrict__ kernel, const float *__restrict__ noise,
                                                           const float *__restrict__ bias, UfdClParams p) {
  __shared__ float sk[kMaxTaps * kMaxTaps];
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    const int ky = i / p.kw, kx = i % p.kw;
    sk[i] = __ldg(kernel + (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx));   // correlate with the flipped kernel
  }
  __syncthreads();
  const int c8n = p.C >> 3;
  const long total = (long)p.B * p.out_h * p.out_w * c8n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int c8 = (int)(r % c8n); r /= c8n;
    const int ox = (int)(r % p.out_w); r /= p.out_w;
    const int oy = (int)(r % p.out_h);
    const int b = (int)(r / p.out_h);
    const int Y0 = oy * p.down - p.pad_y0, X0 = ox * p.down - p.pad_x0;
    const int ky0 = ((-Y0) % p.up + p.up) % p.up, kx0 = ((-X0) % p.up + p.up) % p.up;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ky = ky0; ky < p.kh; ky += p.up) {
      const int iy = (Y0 + ky) / p.up;
      if (iy < 0 || iy >= p.in_h) continue;
      for (int kx = kx0; kx < p.kw; kx += p.up) {
        const int ix = (X0 + kx) / p.up;
        if (ix < 0 || ix >= p.in_w) continue;
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(x + (((size_t)b * p.in_h + iy) * p.in_w + ix) * p.C) + c8);
        const float w = sk[ky * p.kw + kx];
        const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          acc[2 * e] = fmaf(f.x, w, acc[2 * e]), acc[2 * e + 1] = fmaf(f.y, w, acc[2 * e + 1]);
        }
      }
    }
    const float nz = noise != nullptr ? p.noise_weight * __ldg(noise + (size_t)b * p.noise_bstride + (size_t)oy * p.out_w + ox) : 0.0f;
    uint4 o;
    __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = acc[2 * e] + nz, v1 = acc[2 * e + 1] + nz;
      if (bias != nullptr) v0 += __ldg(bias + c8 * 8 + 2 * e), v1 += __ldg(bias + c8 * 8 + 2 * e + 1);
      if (p.act) {
        v0 = (v0 > 0.0f ? v0 : 0.2f * v0) * 1.41421356237309515f;
        v1 = (v1 > 0.0f ? v1 : 0.2f * v1) * 1.41421356237309515f;
      }
      oh[e] = __floats2half2_rn(v0, v1);
    }
    *(reinterpret_cast<uint4 *>(out + (((size_t)b * p.out_h + oy) * p.out_w + ox) * p.C) + c8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// ray generation (dataloader/data_util.py:28-56)
// ------------------------------------------------------------------------------------------------
struct RayGen {
  float cam[16], near, far;   // fx fy cx cy | c2w row-major [3,4]
  int H, W;
};
__global__ void __launch_bounds__(256) get_rays_kernel(float *__restrict__ rays, RayGen g) {
  const int n = g.H * g.W;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    float org[3], d[3];
    pixel_ray([&](int k) { return g.cam[k]; }, g.W, g.H, r, org, d);
    float4 *o = reinterpret_cast<float4 *>(rays + (size_t)r * 8);
    o[0] = make_float4(org[0], org[1], org[2], d[0]);
    o[1] = make_float4(d[1], d[2], g.near, g.far);
  }
}

// ------------------------------------------------------------------------------------------------
// condition renderings: dataloader/dataloader.py:218-229 make_render_cond_ on the device
// ------------------------------------------------------------------------------------------------
// render / normal: [n, px, 3] uint8 RGB as decoded from the ortho_*_256_baseGama.png pair; out: [n, 7, px] float32 =
// render / 255 | normal / 255 | (|normal| > 0) -- channels first, i.e. already permuted the way the entry scripts feed the
// plane generators (train_avatar.py:121-123 .permute(0, 3, 1, 2)).  One thread per pixel, 6 byte loads, 7 coalesced stores.
__global__ void __launch_bounds__(256) render_cond_kernel(float *__restrict__ out, const uint8_t *__restrict__ render,
                                                          const uint8_t *__restrict__ normal, int n, int px) {
  const long total = (long)n * px;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int img = (int)(i / px), p = (int)(i % px);
    const uint8_t *r = render + i * 3, *m = normal + i * 3;
    float *o = out + (size_t)img * 7 * px + p;
    const uint8_t n0 = m[0], n1 = m[1], n2 = m[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) o[(size_t)c * px] = __fdiv_rn((float)r[c], 255.0f);
    o[(size_t)3 * px] = __fdiv_rn((float)n0, 255.0f);
    o[(size_t)4 * px] = __fdiv_rn((float)n1, 255.0f);
    o[(size_t)5 * px] = __fdiv_rn((float)n2, 255.0f);
    o[(size_t)6 * px] = (n0 | n1 | n2) ? 1.0f : 0.0f;   // np.linalg.norm(normal, axis=-1) > 0
  }
}

}  // namespace hav

using namespace hav;

namespace hav {   // fir_cl.cu
cudaError_t launch_blur4_cl_tma(void *out, const void *x, const float *kernel, int batch, int in_h, int in_w, int channels, int out_h,
                                int out_w, int pad_x0, int pad_y0, const float *noise, float noise_weight, int noise_per_sample,
                                const float *bias, int act, cudaStream_t st);
}

extern "C" int hav_make_render_cond(float *out, const uint8_t *render, const uint8_t *normal, int n, int pixels, void *stream) {
  if (n < 0 || pixels < 1) return HAV_E_SHAPE;
  if (n == 0) return HAV_OK;
  if (out == nullptr || render == nullptr || normal == nullptr) return HAV_E_NULL;
  const long total = (long)n * pixels;
  long want = (total + 255) / 256;
  const int grid = (int)(want < (long)kSMs * 16 ? want : (long)kSMs * 16);
  render_cond_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, render, normal, n, pixels);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_fused_bias_act(float *out, const float *x, const float *bias, const float *ref, int64_t numel,
                                  int64_t step_b, int64_t size_b, int act, int grad, float alpha, float scale,
                                  void *stream) {
  if (numel < 0) return HAV_E_SHAPE;
  if (numel == 0) return HAV_OK;
  if (out == nullptr || x == nullptr) return HAV_E_NULL;
  if (bias != nullptr && (step_b < 1 || size_b < 1)) return HAV_E_SHAPE;
  const int mode = act * 10 + grad;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (numel % 4 == 0) && (bias == nullptr || step_b % 4 == 0) &&
                   (((uintptr_t)out | (uintptr_t)x | (uintptr_t)ref) & 15) == 0;
  if (vec) {
    int64_t n4 = numel / 4;
    int64_t want = (n4 + 255) / 256;
    int grid = (int)(want < (int64_t)kSMs * 16 ? want : (int64_t)kSMs * 16);
    bias_act_vec4_kernel<<<grid, 256, 0, st>>>((float4 *)out, (const float4 *)x, bias, (const float4 *)ref, n4,
                                               bias != nullptr ? step_b / 4 : 1, bias != nullptr ? size_b : 1, mode,
                                               alpha, scale);
  } else {
    int64_t want = (numel + 255) / 256;
    int grid = (int)(want < (int64_t)kSMs * 16 ? want : (int64_t)kSMs * 16);
    bias_act_scalar_kernel<<<grid, 256, 0, st>>>(out, x, bias, ref, numel, bias != nullptr ? step_b : 1,
                                                 bias != nullptr ? size_b : 1, mode, alpha, scale);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_bias_act_backward_splits(int batch, int channels, int64_t inner) {
  if (batch < 1 || channels < 1 || inner < 1) return 0;
  const int64_t total = (int64_t)batch * inner;
  int64_t splits = (2 * kSMs + channels - 1) / channels;        // about two CTAs per SM over all channels ...
  const int64_t cap = (total + 2047) / 2048;                    // ... of at least 2048 elements each
  if (splits > cap) splits = cap;
  if (splits > 64) splits = 64;
  return splits < 1 ? 1 : (int)splits;
}

extern "C" int hav_bias_act_backward(float *grad_input, float *partials, const float *grad_out, const float *ref, int batch, int channels,
                                     int64_t inner, int splits, float alpha, float scale, void *stream) {
  if (batch < 0 || channels < 1 || inner < 1 || splits < 1 || splits > 64 || channels > 65535) return HAV_E_SHAPE;
  if (batch == 0) return HAV_OK;
  if (grad_input == nullptr || partials == nullptr || grad_out == nullptr || ref == nullptr) return HAV_E_NULL;
  const int64_t total = (int64_t)batch * inner;
  int64_t per = (total + splits - 1) / splits;
  per = (per + 3) / 4 * 4;
  const bool vec = (inner % 4 == 0) && (((uintptr_t)grad_input | (uintptr_t)grad_out | (uintptr_t)ref) & 15) == 0;
  dim3 grid(splits, channels);
  if (vec) bias_act_backward_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(grad_input, partials, grad_out, ref, batch, channels, inner, per, alpha, scale);
  else bias_act_backward_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(grad_input, partials, grad_out, ref, batch, channels, inner, per, alpha, scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_noise_bias_act(float *out, const float *x, const float *bias, const float *noise, const float *noise_weight, int batch,
                                  int channels, int64_t inner, int noise_per_sample, float alpha, float scale, void *stream) {
  if (batch < 0 || channels < 1 || inner < 1) return HAV_E_SHAPE;
  if (batch == 0) return HAV_OK;
  if (out == nullptr || x == nullptr || noise == nullptr || noise_weight == nullptr) return HAV_E_NULL;
  const bool vec = (inner % 4 == 0) && (((uintptr_t)out | (uintptr_t)x | (uintptr_t)noise) & 15) == 0;
  const long n_units = (long)batch * channels * (inner / (vec ? 4 : 1));
  const long want = (n_units + 255) / 256;
  const int grid = (int)(want < (long)kSMs * 16 ? want : (long)kSMs * 16);
  const long nb = noise_per_sample ? inner : 0;
  if (vec) noise_bias_act_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(out, x, bias, noise, noise_weight, n_units, inner, channels, nb, alpha, scale);
  else noise_bias_act_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(out, x, bias, noise, noise_weight, n_units, inner, channels, nb, alpha, scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_noise_bias_act_backward(float *grad_input, float *partials, float *noise_partials, const float *grad_out, const float *ref,
                                           const float *noise, int batch, int channels, int64_t inner, int noise_per_sample, int splits,
                                           float alpha, float scale, void *stream) {
  if (batch < 0 || channels < 1 || inner < 1 || splits < 1 || splits > 64 || channels > 65535) return HAV_E_SHAPE;
  if (batch == 0) return HAV_OK;
  if (grad_input == nullptr || partials == nullptr || noise_partials == nullptr || grad_out == nullptr || ref == nullptr || noise == nullptr)
    return HAV_E_NULL;
  const int64_t total = (int64_t)batch * inner;
  int64_t per = (total + splits - 1) / splits;
  per = (per + 3) / 4 * 4;
  const bool vec = (inner % 4 == 0) && (((uintptr_t)grad_input | (uintptr_t)grad_out | (uintptr_t)ref | (uintptr_t)noise) & 15) == 0;
  dim3 grid(splits, channels);
  const long nb = noise_per_sample ? inner : 0;
  if (vec) bias_act_backward_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(grad_input, partials, grad_out, ref, batch, channels, inner, per, alpha, scale, noise, nb, noise_partials);
  else bias_act_backward_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(grad_input, partials, grad_out, ref, batch, channels, inner, per, alpha, scale, noise, nb, noise_partials);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_upfirdn2d(float *out, const float *x, const float *kernel, int major, int in_h, int in_w,
                             int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0,
                             int pad_x1, int pad_y0, int pad_y1, void *stream) {
  if (major < 0 || in_h < 1 || in_w < 1 || minor < 1 || kh < 1 || kw < 1) return HAV_E_SHAPE;
  if (up_x < 1 || up_y < 1 || down_x < 1 || down_y < 1) return HAV_E_SHAPE;
  if (kh > kMaxTaps || kw > kMaxTaps) return HAV_E_SHAPE;
  const int out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) / down_y;  // upfirdn2d_kernel.cu:236-241
  const int out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) / down_x;
  if (out_h < 1 || out_w < 1) return HAV_E_SHAPE;
  if (major == 0) return HAV_OK;
  if (out == nullptr || x == nullptr || kernel == nullptr) return HAV_E_NULL;
  if ((int64_t)major * minor > 65535 && minor != 1) return HAV_E_SHAPE;  // gridDim.y chunking needs minor == 1
  UfdParams p;
  p.in_h = in_h, p.in_w = in_w, p.minor = minor, p.out_h = out_h, p.out_w = out_w, p.kh = kh, p.kw = kw;
  p.up_x = up_x, p.up_y = up_y, p.down_x = down_x, p.down_y = down_y, p.pad_x0 = pad_x0, p.pad_y0 = pad_y0;
  p.win_h = ((kTileH - 1) * down_y + kh - 1) / up_y + 2;
  p.win_w = ((kTileW - 1) * down_x + kw - 1) / up_x + 2;
  if (minor == 1 && up_x == up_y && down_x == down_y && kh == kw && pad_x0 >= 0 && pad_y0 >= 0) {
    // the StyleUNet's own shapes (model/styleUnet.py:29-87, 371-422) take the specialised kernels
    cudaError_t fe = cudaErrorInvalidValue;
    const int64_t n_img = (int64_t)major;
    static const bool old_blur = getenv("HAV_BLUR_TILED") != nullptr;   // A/B aid: the shared-memory tile kernel of round 1
    if (up_x == 1 && down_x == 1 && kh == 4 && !old_blur && ((uintptr_t)x & 15) == 0)
      fe = launch_blur4x4_stream(out, x, kernel, p, n_img, (cudaStream_t)stream);
    else if (up_x == 1 && down_x == 1 && kh == 4) fe = launch_ufd_fast<1, 1, 4, 4>(out, x, kernel, p, n_img, (cudaStream_t)stream);
    else if (up_x == 1 && down_x == 2 && kh == 4) fe = launch_ufd_fast<1, 2, 4, 4>(out, x, kernel, p, n_img, (cudaStream_t)stream);
    else if (up_x == 2 && down_x == 1 && kh == 4) fe = launch_ufd_fast<2, 1, 4, 4>(out, x, kernel, p, n_img, (cudaStream_t)stream);
    else if (up_x == 1 && down_x == 2 && kh == 2) fe = launch_ufd_fast<1, 2, 2, 2>(out, x, kernel, p, n_img, (cudaStream_t)stream);
    else if (up_x == 2 && down_x == 1 && kh == 2) fe = launch_ufd_fast<2, 1, 2, 2>(out, x, kernel, p, n_img, (cudaStream_t)stream);
    if (fe == cudaSuccess) return HAV_OK;
    if (fe != cudaErrorInvalidValue) return (int)fe;
  }
  const size_t smem = (size_t)(kMaxTaps * kMaxTaps + p.win_h * p.win_w) * sizeof(float);
  if (smem > 200 * 1024) return HAV_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(upfirdn2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int tiles_x = (out_w + kTileW - 1) / kTileW, tiles_y = (out_h + kTileH - 1) / kTileH;
  const int64_t imgs = (int64_t)major * minor;
  // gridDim.y is limited to 65535: loop over image chunks
  for (int64_t i0 = 0; i0 < imgs; i0 += 65535) {
    int ny = (int)((imgs - i0) < 65535 ? (imgs - i0) : 65535);
    dim3 grid(tiles_x * tiles_y, ny);
    upfirdn2d_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(out + (size_t)i0 * out_h * out_w,
                                                                 x + (size_t)i0 * in_h * in_w, kernel, p, tiles_x);
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_upfirdn2d_cl(void *out, const void *x, const float *kernel, int batch, int in_h, int in_w, int channels,
                                int kh, int kw, int up, int down, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                                const float *noise, float noise_weight, int noise_per_sample, const float *bias, int act,
                                void *stream) {
  if (batch < 0 || in_h < 1 || in_w < 1 || channels < 8 || (channels & 7) || kh < 1 || kw < 1 || up < 1 || down < 1) return HAV_E_SHAPE;
  if (kh > kMaxTaps || kw > kMaxTaps) return HAV_E_SHAPE;
  const int out_h = (in_h * up + pad_y0 + pad_y1 - kh + down) / down, out_w = (in_w * up + pad_x0 + pad_x1 - kw + down) / down;
  if (out_h < 1 || out_w < 1) return HAV_E_SHAPE;
  if (batch == 0) return HAV_OK;
  if (out == nullptr || x == nullptr || kernel == nullptr) return HAV_E_NULL;
  if (up == 1 && down == 1 && kh == 4 && kw == 4 && (channels & 63) == 0) {
    // every Blur of the StyleUNet: TMA-tiled kernel (fir_cl.cu); HAV_FIR_GENERIC=1 keeps the generic kernel for A/B runs
    static const bool generic = getenv("HAV_FIR_GENERIC") != nullptr;
    if (!generic) {
      cudaError_t fe = launch_blur4_cl_tma(out, x, kernel, batch, in_h, in_w, channels, out_h, out_w, pad_x0, pad_y0, noise, noise_weight,
                                           noise_per_sample, bias, act, (cudaStream_t)stream);
      if (fe == cudaSuccess) return HAV_OK;
      if (fe != cudaErrorNotSupported) return (int)fe;
    }
  }
  UfdClParams p;
  p.B = batch, p.in_h = in_h, p.in_w = in_w, p.C = channels, p.out_h = out_h, p.out_w = out_w, p.kh = kh, p.kw = kw;
  p.up = up, p.down = down, p.pad_x0 = pad_x0, p.pad_y0 = pad_y0, p.act = act, p.noise_weight = noise_weight;
  p.noise_bstride = noise_per_sample ? out_h * out_w : 0;
  const long total = (long)batch * out_h * out_w * (channels / 8);
  long want = (total + 255) / 256;
  const int grid = (int)(want < (long)kSMs * 32 ? want : (long)kSMs * 32);
  upfirdn2d_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint16_t *)out, (const uint16_t *)x, kernel, noise, bias, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_get_rays(float *ray_batch, int height, int width, const float intr[4], const float c2w[12],
                            float near, float far, void *stream) {
  if (height < 1 || width < 1 || (int64_t)height * width > (int64_t)1 << 30) return HAV_E_SHAPE;
  if (ray_batch == nullptr || intr == nullptr || c2w == nullptr) return HAV_E_NULL;
  if (intr[0] == 0.0f || intr[1] == 0.0f) return HAV_E_VALUE;
  RayGen g;
  for (int i = 0; i < 4; ++i) g.cam[i] = intr[i];
  for (int i = 0; i < 12; ++i) g.cam[4 + i] = c2w[i];
  g.near = near, g.far = far, g.H = height, g.W = width;
  const int n = height * width;
  int grid = (n + 255) / 256;
  if (grid > kSMs * 8) grid = kSMs * 8;
  get_rays_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ray_batch, g);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}
