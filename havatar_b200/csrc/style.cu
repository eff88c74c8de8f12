// Modulation vectors and demodulation factors of EVERY modulated convolution of one network in two launches
// (model/styleUnet.py:237-258 per layer: s = EqualLinear(style), demod = rsqrt(sum((scale * W * s)^2) + eps)).
// The reference (and the per-layer path here) runs one small linear + one reduction per layer -- ~40 dependent 4-5 us
// launches on the critical path of an inference frame, each reading the full 3x3 weight tensor for the demodulation.  Here a
// device table describes the layers once; kernel 1 computes all s[b,ci] (one warp per modulation row), kernel 2 all
// demod[b,co] from the per-layer tap-summed squares wsq[co,ci] = sum_k W[co,ci,k]^2 (cached by the caller per weight version:
// demod = rsqrt(scale^2 * sum_ci s^2 * wsq + eps), a ninth of the bytes).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/havatar_b200.h"

namespace hav {
namespace style {

constexpr int kMaxBatch = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// layer that owns global row `row` of the concatenated row space (prefix[l] = first row of layer l, prefix[n] = total)
__device__ __forceinline__ int find_layer(const int *__restrict__ prefix, int n, int row) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(prefix + mid) <= row) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) modulate_kernel(float *__restrict__ s_all, const float *__restrict__ latent, int B, int n_latent, int D,
                                                       const hav_style_layer *__restrict__ L, const int *__restrict__ prefix, int n_layers,
                                                       int total_rows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= total_rows) return;
  const int l = find_layer(prefix, n_layers, warp);
  const hav_style_layer ly = L[l];
  const int ci = warp - __ldg(prefix + l);
  const float *w = ly.mod_w + (size_t)ci * D;
  float acc[kMaxBatch];
#pragma unroll
  for (int b = 0; b < kMaxBatch; ++b) acc[b] = 0.0f;
  for (int d = lane; d < D; d += 32) {
    const float wv = __ldg(w + d);
#pragma unroll
    for (int b = 0; b < kMaxBatch; ++b)
      if (b < B) acc[b] = fmaf(wv, __ldg(latent + ((size_t)b * n_latent + ly.latent_index) * D + d), acc[b]);
  }
  const float bias = ly.mod_b != nullptr ? __ldg(ly.mod_b + ci) * ly.mod_lr_mul : 0.0f;
#pragma unroll
  for (int b = 0; b < kMaxBatch; ++b) {
    if (b >= B) break;
    const float v = warp_sum(acc[b]);
    if (lane == 0) s_all[ly.s_off + (size_t)b * ly.cin + ci] = fmaf(v, ly.mod_scale, bias);
  }
}

__global__ void __launch_bounds__(256) demod_kernel(float *__restrict__ d_all, const float *__restrict__ s_all, int B,
                                                    const hav_style_layer *__restrict__ L, const int *__restrict__ prefix, int n_layers,
                                                    int total_rows, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= total_rows) return;
  const int l = find_layer(prefix, n_layers, warp);
  const hav_style_layer ly = L[l];
  if (ly.wsq == nullptr) return;
  const int co = warp - __ldg(prefix + l);
  const float *q = ly.wsq + (size_t)co * ly.cin;
  float acc[kMaxBatch];
#pragma unroll
  for (int b = 0; b < kMaxBatch; ++b) acc[b] = 0.0f;
  for (int ci = lane; ci < ly.cin; ci += 32) {
    const float qv = __ldg(q + ci);
#pragma unroll
    for (int b = 0; b < kMaxBatch; ++b)
      if (b < B) {
        const float sv = s_all[ly.s_off + (size_t)b * ly.cin + ci] * ly.conv_scale;
        acc[b] = fmaf(qv, sv * sv, acc[b]);
      }
  }
#pragma unroll
  for (int b = 0; b < kMaxBatch; ++b) {
    if (b >= B) break;
    const float v = warp_sum(acc[b]);
    if (lane == 0) d_all[ly.d_off + (size_t)b * ly.cout + co] = rsqrtf(v + eps);
  }
}

// wsq[co,ci] = sum_k w[co,ci,k]^2
__global__ void __launch_bounds__(256) tap_squares_kernel(float *__restrict__ wsq, const float *__restrict__ w, long n, int kk) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float q = 0.0f;
    for (int t = 0; t < kk; ++t) {
      const float v = __ldg(w + i * kk + t);
      q = fmaf(v, v, q);
    }
    wsq[i] = q;
  }
}

}  // namespace style
}  // namespace hav

using namespace hav;

extern "C" int hav_conv_tap_squares(float *wsq, const float *w, int cout, int cin, int ksize, void *stream) {
  if (wsq == nullptr || w == nullptr) return HAV_E_NULL;
  if (cout < 1 || cin < 1 || ksize < 1) return HAV_E_SHAPE;
  const long n = (long)cout * cin;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  style::tap_squares_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(wsq, w, n, ksize * ksize);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_style_plan_run(float *s_all, float *d_all, const float *latent, int batch, int n_latent, int style_dim,
                                  const hav_style_layer *layers, const int *s_prefix, const int *d_prefix, int n_layers, int s_rows,
                                  int d_rows, float eps, void *stream) {
  if (s_all == nullptr || latent == nullptr || layers == nullptr || s_prefix == nullptr) return HAV_E_NULL;
  if (batch < 1 || batch > style::kMaxBatch || n_latent < 1 || style_dim < 1 || n_layers < 1 || s_rows < 1 || d_rows < 0) return HAV_E_SHAPE;
  if (d_rows > 0 && (d_all == nullptr || d_prefix == nullptr)) return HAV_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  style::modulate_kernel<<<(s_rows + 7) / 8, 256, 0, st>>>(s_all, latent, batch, n_latent, style_dim, layers, s_prefix, n_layers, s_rows);
  if (d_rows > 0)
    style::demod_kernel<<<(d_rows + 7) / 8, 256, 0, st>>>(d_all, s_all, batch, layers, d_prefix, n_layers, d_rows, eps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}
