// fp32 CUDA-core fused render kernel: the reference-exact mode (HAV_PREC_FP32).
//
// One thread owns one ray and marches it front to back: depth sample -> skinning warp -> bi-plane
// fetch -> positional encoding -> MLP (fp32 FMA, weights broadcast from L1) -> composite.  No
// per-sample tensor ever reaches HBM (the reference materialises ~6-8 KB per sample,
// model/nerf_trainer.py:141-163).  Summation order inside each ray matches the reference
// (sequential transmittance product), so parity with the oracle is ~1e-6.
#include "render_common.cuh"
#include "render_internal.h"

namespace hav {

constexpr int kHeadN = 68;  // fp32 head matrix columns: 0..63 fc_rgbFeat, 64 fc_alpha, 65..67 zero pad

// ---- weight packing: nn.Linear [N,K] row-major -> [K][N] so that a warp-uniform float4 load yields
// four output neurons of one input (model/nerf_model.py:46-51) ----
__global__ void pack_mlp_fp32_kernel(const float *__restrict__ w0, const float *__restrict__ b0,
                                     const float *__restrict__ w1, const float *__restrict__ b1,
                                     const float *__restrict__ wa, const float *__restrict__ ba,
                                     const float *__restrict__ wf, const float *__restrict__ bf,
                                     const float *__restrict__ wr, const float *__restrict__ br, float *out) {
  float *W0t = out + kOffW0t, *W1t = out + kOffW1t, *Wht = out + kOffWht;
  float *B0 = out + kOffB0, *B1 = out + kOffB1, *Bh = out + kOffBh, *Wr = out + kOffWr, *Br = out + kOffBr;
  int i = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
  for (int t = i; t < kIn * kHid; t += n) W0t[t] = w0[(t % kHid) * kIn + t / kHid];
  for (int t = i; t < kHid * kHid; t += n) W1t[t] = w1[(t % kHid) * kHid + t / kHid];
  for (int t = i; t < kHid * kHeadN; t += n) {
    int k = t / kHeadN, c = t % kHeadN;
    Wht[t] = c < kRgbFeat ? wf[c * kHid + k] : (c == kRgbFeat ? wa[k] : 0.0f);
  }
  for (int t = i; t < kHid; t += n) B0[t] = b0[t], B1[t] = b1[t];
  for (int t = i; t < kHeadN; t += n) Bh[t] = t < kRgbFeat ? bf[t] : (t == kRgbFeat ? ba[0] : 0.0f);
  for (int t = i; t < 3 * kRgbFeat; t += n) Wr[t] = wr[t];
  for (int t = i; t < 3; t += n) Br[t] = br[t];
}

void launch_pack_mlp_fp32(const hav_render_args *a, float *out, cudaStream_t st) {
  pack_mlp_fp32_kernel<<<32, 256, 0, st>>>(a->w0, a->b0, a->w1, a->b1, a->w_alpha, a->b_alpha, a->w_feat, a->b_feat,
                                           a->w_rgb, a->b_rgb, out);
}

// acc[n] = b[n] + sum_k xs[k][tid] * Wt[k][n]
template <int N>
__device__ __forceinline__ void dense(const float *xs, int K, const float *__restrict__ Wt, const float *__restrict__ b,
                                      float (&acc)[N]) {
#pragma unroll
  for (int n = 0; n < N; ++n) acc[n] = __ldg(b + n);
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float xk = xs[k * kRaysPerBlock];
    const float4 *wr = reinterpret_cast<const float4 *>(Wt + (size_t)k * N);
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      float4 w = __ldg(wr + q);
      acc[4 * q + 0] = fmaf(xk, w.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(xk, w.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(xk, w.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(xk, w.w, acc[4 * q + 3]);
    }
  }
}

// smem: xs[kIn][128] activations (x, then h0, then h1 in rows 0..127) | sums[kOut][128] composite accumulators
constexpr int kSimtSmem = (kIn + kOut) * kRaysPerBlock * (int)sizeof(float);

__global__ void __launch_bounds__(kRaysPerBlock, 1) render_fp32_kernel(const RenderDev P) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  float *xs = smem + tid;                                  // xs[k*128]
  float *sums = smem + kIn * kRaysPerBlock + tid;          // sums[c*128]
  const int g = blockIdx.x * kRaysPerBlock + tid;
  const Ray ray = load_ray(P, g);
  const int gi = ray.valid ? g : 0;
  const float *T = P.invT + (size_t)ray.b * 12;
  float Tm[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) Tm[i] = __ldg(T + i);
  const size_t plane_sz = (size_t)P.PH * P.PW;
  const float *pl0 = P.planes + ((size_t)0 * P.B + ray.b) * kPlaneC * plane_sz;  // XY plane of this frame
  const float *pl1 = P.planes + ((size_t)1 * P.B + ray.b) * kPlaneC * plane_sz;  // ZY plane
  float *zcol = P.zbuf + (size_t)blockIdx.x * P.Sf * kRaysPerBlock + tid;
  float *wcol = P.wbuf + (size_t)blockIdx.x * P.Sc * kRaysPerBlock + tid;
  float bgc[3] = {0.f, 0.f, 0.f};
  if (P.bg != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) bgc[c] = __ldg(P.bg + (size_t)gi * 3 + c);
  }

  const int npass = P.nfine > 0 ? 2 : 1;
  for (int pass = 0; pass < npass; ++pass) {
    const int S = pass == 0 ? P.Sc : P.Sf;
    const float *noise = pass == 0 ? P.noise_c : P.noise_f;
    Composite cs;
    cs.reset();
#pragma unroll 1
    for (int c = 0; c < kOut; ++c) sums[c * kRaysPerBlock] = 0.0f;
    float z_cur = pass == 0 ? coarse_z(P, ray, gi, 0) : zcol[0];
    float dist_prev = 0.0f;
#pragma unroll 1
    for (int s = 0; s < S; ++s) {
      // ---- depth + sample distance (model/nerf_trainer.py:129-141, utils/nerf_util.py:36-38)
      float z_next = 0.0f, dist;
      if (s + 1 < S) {
        z_next = pass == 0 ? coarse_z(P, ray, gi, s + 1) : zcol[(s + 1) * kRaysPerBlock];
        dist = z_next - z_cur;
      } else {
        dist = dist_prev;
      }
      dist_prev = dist;
      const float z = z_cur;
      z_cur = z_next;
      float p[3], pc[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) p[j] = __fadd_rn(ray.o[j], __fmul_rn(ray.d[j], z));  // separate mul/add like ATen (:141)
      // ---- skinning warp (model/Skinning_Field.py:70-98)
      skin_warp(P, Tm, p, pc);
      // ---- bi-plane features, feature index = 2*c + plane (utils/util.py:359-392, model/nerf_model.py:88-99)
      {
        float qx = pc[0] * P.ps[0] + P.pt[0], qy = pc[1] * P.ps[1] + P.pt[1], qz = pc[2] * P.ps[2] + P.pt[2];
        Taps t0 = bilinear_taps(qx, qy, P.PH, P.PW);
        Taps t1 = bilinear_taps(qz, qy, P.PH, P.PW);
#pragma unroll 4
        for (int c = 0; c < kPlaneC; ++c) {
          const float *a = pl0 + c * plane_sz, *bq = pl1 + c * plane_sz;
          float f0 = 0.0f, f1 = 0.0f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (t0.w[k] != 0.0f) f0 += __ldg(a + t0.off[k]) * t0.w[k];
            if (t1.w[k] != 0.0f) f1 += __ldg(bq + t1.off[k]) * t1.w[k];
          }
          xs[(2 * c) * kRaysPerBlock] = f0;
          xs[(2 * c + 1) * kRaysPerBlock] = f1;
        }
      }
      // ---- positional encoding, order [f][sin|cos][xyz], cos = sin(a + pi/2) (model/network/embedder.py:32-61)
#pragma unroll
      for (int f = 0; f < kFreqs; ++f) {
        const float fr = (float)(1 << f);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float ang = pc[j] * fr;
          xs[(kFeat + f * 6 + j) * kRaysPerBlock] = sinf(ang);
          xs[(kFeat + f * 6 + 3 + j) * kRaysPerBlock] = sinf(ang + 1.57079632679489661923f);
        }
      }
      // ---- MLP (model/nerf_model.py:101-117)
      float alpha_raw, rgb_raw[3];
      {
        float acc[kHid];
        dense<kHid>(xs, kIn, P.W0t, P.b0, acc);
#pragma unroll
        for (int n = 0; n < kHid; ++n) xs[n * kRaysPerBlock] = fmaxf(acc[n], 0.0f);
        dense<kHid>(xs, kHid, P.W1t, P.b1, acc);
#pragma unroll
        for (int n = 0; n < kHid; ++n) xs[n * kRaysPerBlock] = fmaxf(acc[n], 0.0f);
      }
      float hd[kHeadN];
      dense<kHeadN>(xs, kHid, P.Wht, P.bh, hd);
      alpha_raw = hd[kRgbFeat];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float r = __ldg(P.br + j);
#pragma unroll
        for (int c = 0; c < kRgbFeat; ++c) r = fmaf(hd[c], __ldg(P.Wr + j * kRgbFeat + c), r);
        rgb_raw[j] = r;
      }
      // ---- composite (utils/nerf_util.py:28-73)
      float nz = noise != nullptr ? __ldg(noise + (size_t)gi * S + s) : 0.0f;
      float w = cs.step(alpha_raw, nz, dist * ray.dnorm, z);
      if (pass == 0 && npass == 2) wcol[s * kRaysPerBlock] = w;
#pragma unroll
      for (int j = 0; j < 3; ++j) sums[j * kRaysPerBlock] += w * sigmoidf_exact(rgb_raw[j]);
#pragma unroll
      for (int c = 0; c < kRgbFeat; ++c) sums[(3 + c) * kRaysPerBlock] += w * hd[c];
    }
    // ---- write the ray (utils/nerf_util.py:62-71)
    if (ray.valid) {
      float *rgb = (pass == 0 ? P.rgb_c : P.rgb_f) + (size_t)g * kOut;
#pragma unroll 1
      for (int c = 0; c < kOut; ++c) {
        float v = sums[c * kRaysPerBlock];
        if (c < 3 && P.bg != nullptr) v = v + (1.0f - cs.acc) * bgc[c];
        rgb[c] = v;
      }
      (pass == 0 ? P.depth_c : P.depth_f)[g] = cs.depth;
      (pass == 0 ? P.acc_c : P.acc_f)[g] = cs.acc;
      if (pass == npass - 1) P.wmax[g] = cs.wmax;
    }
    // ---- hierarchical resampling (utils/nerf_util.py:76-117, model/nerf_trainer.py:165-170)
    if (pass == 0 && npass == 2) {
      auto zc = [&](int s) { return coarse_z(P, ray, gi, s); };
      sample_pdf_merge(zc, P.Sc, P.nfine, wcol, kRaysPerBlock,
                       P.u_rand != nullptr ? P.u_rand + (size_t)gi * P.nfine : nullptr, zcol,
                         (ray.valid && P.pdf_inds != nullptr) ? P.pdf_inds + (size_t)g * P.nfine : nullptr);
      if (ray.valid && P.z_fine != nullptr)
        for (int j = 0; j < P.Sf; ++j) P.z_fine[(size_t)g * P.Sf + j] = zcol[j * kRaysPerBlock];
    }
  }
}

cudaError_t launch_render_fp32(const RenderDev &P, int num_blocks, cudaStream_t st) {
  // the attribute is per device / context: set it on every launch (a process may drive several GPUs)
  cudaError_t e = cudaFuncSetAttribute(render_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSimtSmem);
  if (e != cudaSuccess) return e;
  render_fp32_kernel<<<num_blocks, kRaysPerBlock, kSimtSmem, st>>>(P);
  return cudaGetLastError();
}

}  // namespace hav

// ------------------------------------------------------------------------------------------------
// hav_sample_pdf: utils/nerf_util.py:76-117 alone (the device function the render kernels run), one thread per row
// ------------------------------------------------------------------------------------------------
namespace hav {
__global__ void __launch_bounds__(128) sample_pdf_kernel(const float *__restrict__ bins, const float *__restrict__ weights,
                                                         const float *__restrict__ u, int n, int m, int nfine, float *__restrict__ samples,
                                                         int32_t *__restrict__ inds, float *__restrict__ scratch) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int Sc = m + 1;
  float *wcol = scratch + (size_t)r * Sc;
  for (int j = 1; j <= Sc - 2; ++j) wcol[j] = weights[(size_t)r * (m - 1) + j - 1];
  float zs[kMaxFine];
  const float *b = bins + (size_t)r * m;
  sample_pdf_core([&](int j) { return b[j]; }, Sc, nfine, wcol, 1, u != nullptr ? u + (size_t)r * nfine : nullptr, zs,
                  inds != nullptr ? inds + (size_t)r * nfine : nullptr);
  for (int k = 0; k < nfine; ++k) samples[(size_t)r * nfine + k] = zs[k];
}
}  // namespace hav

extern "C" int hav_sample_pdf(const float *bins, const float *weights, const float *u, int n, int m, int nfine, float *samples,
                              int32_t *inds, float *scratch, void *stream) {
  if (n < 0 || m < 2 || m + 1 > hav::kMaxSamples || nfine < 1 || nfine > hav::kMaxFine) return HAV_E_SHAPE;
  if (n == 0) return HAV_OK;
  if (bins == nullptr || weights == nullptr || samples == nullptr || scratch == nullptr) return HAV_E_NULL;
  hav::sample_pdf_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bins, weights, u, n, m, nfine, samples, inds, scratch);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

