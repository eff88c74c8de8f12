// 4x4 FIR (up == down == 1) on channels-last fp16 tensors [B,H,W,C], C % 64 == 0: the Blur after every transposed and before
// every stride-2 convolution of the StyleUNet (model/styleUnet.py:69-87, :264-287) in the inference hand-over layout, with the
// StyledConv tail (noise, bias, leaky-relu, :593-599) fused in.
//
// TMA-tiled: the input is described by a rank-4 tensor map (C, W, H, B); a tile = 32 x 8 output pixels x 64 channels, whose
// (32+3) x (8+3) x 64-channel input window (one pixel's 64 channels = one 128-byte line) arrives as ONE cp.async.bulk.tensor
// box load signalled on an mbarrier.  The zero padding of the filter is the tensor map's out-of-bounds fill (coordinates start
// at -pad), so the kernel has no boundary code on the load side.  Persistent CTAs walk the tile list with a two-stage ring: the
// box of tile i+1 is in flight while tile i is filtered.  A thread owns one output column x 8 channels and marches down the
// rows of the tile: four 16-byte shared-memory loads per input row (conflict-free: 8 lanes cover one pixel's line), horizontal
// sums kept for the last three rows in registers (rank-one taps: 4 + 4 FMAs per output and channel; anything else the general
// 16-tap form), fp32 accumulation, one 16-byte store per output pixel and channel octet.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/havatar_b200.h"
#include "tc_common.cuh"

namespace hav {
namespace fircl {

using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

constexpr int kTW = 32, kTH = 8, kCB = 64;              // output tile and channel block
constexpr int kIW = kTW + 3, kIH = kTH + 3;
constexpr int kStageBytes = kIW * kIH * kCB * 2;        // 49280
constexpr int kStages = 2;
constexpr int kSmemBytes = kStages * kStageBytes + 64;  // + mbarriers
constexpr int kThreads = 256, kCtasPerSm = 2;             // 2 x 98.6 KB of shared memory, 16 warps per SM

struct Params {
  int B, C, out_h, out_w, pad_x0, pad_y0, act, noise_bstride;
  int tiles_x, tiles_y, cblocks;
  long n_tiles;
  float noise_weight;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
  const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = __half22float2(h[e]);
    f[2 * e] = v.x, f[2 * e + 1] = v.y;
  }
}

template <bool kSep>
__device__ __forceinline__ void filter_tile(const uint8_t *__restrict__ st, uint16_t *__restrict__ out, const float *__restrict__ noise,
                                            const float (&bs)[8], const Params &p, int b, int oy0, int ox, int c0, const float (&kx)[4],
                                            const float (&ky)[4], const float (&k2)[16]) {
  // input row r of the window feeds output rows r-3 .. r (taps 3 .. 0 of the flipped kernel).  Sliding state: separable -> the
  // horizontal sums of the last three input rows (the window rotates through H[0..3] by renaming); general -> partial output rows
  float H[4][8], part[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) H[i][e] = 0.0f, part[i][e] = 0.0f;
  const int xl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  const uint4 *row = reinterpret_cast<const uint4 *>(st) + xl * 8 + cq;
  const bool col_ok = ox < p.out_w;
  const int total = min(kTH, p.out_h - oy0) + 3;
  uint16_t *op = out + (((size_t)b * p.out_h + oy0) * p.out_w + ox) * p.C + c0 + cq * 8 - (size_t)3 * p.out_w * p.C;   // output row r - 3
  const float *np = noise != nullptr ? noise + (size_t)b * p.noise_bstride + (size_t)oy0 * p.out_w + ox - (size_t)3 * p.out_w : nullptr;
  auto one_row = [&](int r, const float (&Hm3)[8], const float (&Hm2)[8], const float (&Hm1)[8], float (&Hnew)[8]) {
    float t[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack8(row[j * 8], t[j]);
    float o[8];
    if (kSep) {
#pragma unroll
      for (int e = 0; e < 8; ++e) Hnew[e] = fmaf(t[3][e], kx[3], fmaf(t[2][e], kx[2], fmaf(t[1][e], kx[1], t[0][e] * kx[0])));
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(Hnew[e], ky[3], fmaf(Hm1[e], ky[2], fmaf(Hm2[e], ky[1], Hm3[e] * ky[0])));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) part[i][e] = fmaf(t[j][e], k2[i * 4 + j], part[i][e]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o[e] = part[3][e];
        part[3][e] = part[2][e], part[2][e] = part[1][e], part[1][e] = part[0][e], part[0][e] = 0.0f;
      }
    }
    if (r >= 3 && col_ok) {
      const float nz = np != nullptr ? p.noise_weight * __ldg(np) : 0.0f;
      uint4 q;
      __half2 *qh = reinterpret_cast<__half2 *>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v0 = o[2 * e] + nz + bs[2 * e], v1 = o[2 * e + 1] + nz + bs[2 * e + 1];
        if (p.act) {
          v0 = (v0 > 0.0f ? v0 : 0.2f * v0) * 1.41421356237309515f;
          v1 = (v1 > 0.0f ? v1 : 0.2f * v1) * 1.41421356237309515f;
        }
        qh[e] = __floats2half2_rn(v0, v1);
      }
      *reinterpret_cast<uint4 *>(op) = q;
    }
    row += kIW * 8, op += (size_t)p.out_w * p.C;
    if (np != nullptr) np += p.out_w;
  };
#pragma unroll 1
  for (int r = 0; r < total; r += 4) {
    one_row(r, H[1], H[2], H[3], H[0]);
    if (r + 1 < total) one_row(r + 1, H[2], H[3], H[0], H[1]);
    if (r + 2 < total) one_row(r + 2, H[3], H[0], H[1], H[2]);
    if (r + 3 < total) one_row(r + 3, H[0], H[1], H[2], H[3]);
  }
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm) blur4_cl_tma_kernel(const __grid_constant__ CUtensorMap xmap, uint16_t *__restrict__ out,
                                                                   const float *__restrict__ kernel, const float *__restrict__ noise,
                                                                   const float *__restrict__ bias, const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem), bar0 = sbase + kStages * kStageBytes;
  // flipped taps (upfirdn2d correlates with the flipped kernel) and their rank-one factorisation k2[i][j] = ky[i] * kx[j]
  float k2[16], kx[4], ky[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) k2[i] = __ldg(kernel + 15 - i);
  int best = 0;
#pragma unroll
  for (int i = 1; i < 16; ++i)
    if (fabsf(k2[i]) > fabsf(k2[best])) best = i;
  const float piv = k2[best];
  bool sep = piv != 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) kx[j] = k2[(best >> 2) * 4 + j];
#pragma unroll
  for (int i = 0; i < 4; ++i) ky[i] = sep ? k2[i * 4 + (best & 3)] / piv : 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) sep = sep && fabsf(k2[i] - ky[i >> 2] * kx[i & 3]) <= 1e-6f * fabsf(piv);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long t, int s) {   // thread 0 only: the box load of tile t into stage s
    long r = t;
    const int tx = (int)(r % p.tiles_x); r /= p.tiles_x;
    const int ty = (int)(r % p.tiles_y); r /= p.tiles_y;
    const int cb = (int)(r % p.cblocks);
    const int b = (int)(r / p.cblocks);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + s * 8), "r"((uint32_t)kStageBytes) : "memory");
    tma_load_4d(sbase + s * kStageBytes, &xmap, cb * kCB, tx * kTW - p.pad_x0, ty * kTH - p.pad_y0, b, bar0 + s * 8);
  };
  if (threadIdx.x == 0 && (long)blockIdx.x < p.n_tiles) issue(blockIdx.x, 0);
  int it = 0;
  for (long t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
    const int s = it & 1;
    // stage s^1 was read by iteration it-1, which every thread left through the __syncthreads below
    if (threadIdx.x == 0 && t + gridDim.x < p.n_tiles) issue(t + gridDim.x, s ^ 1);
    long r = t;
    const int tx = (int)(r % p.tiles_x); r /= p.tiles_x;
    const int ty = (int)(r % p.tiles_y); r /= p.tiles_y;
    const int cb = (int)(r % p.cblocks);
    const int b = (int)(r / p.cblocks);
    const int ox = tx * kTW + (threadIdx.x >> 3), c0 = cb * kCB;
    float bs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bs[e] = bias != nullptr ? __ldg(bias + c0 + (threadIdx.x & 7) * 8 + e) : 0.0f;
    mbar_wait(bar0 + s * 8, (it >> 1) & 1);
    if (sep) filter_tile<true>(smem + s * kStageBytes, out, noise, bs, p, b, ty * kTH, ox, c0, kx, ky, k2);
    else filter_tile<false>(smem + s * kStageBytes, out, noise, bs, p, b, ty * kTH, ox, c0, kx, ky, k2);
    __syncthreads();
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return (EncodeTiledFn)sym;
  }();
  return fn;
}

}  // namespace fircl

// the 4x4, up = down = 1 case of hav_upfirdn2d_cl with C % 64 == 0.  Returns cudaErrorNotSupported when the tensor cannot be
// described by a tensor map (the caller then runs the generic kernel).
cudaError_t launch_blur4_cl_tma(void *out, const void *x, const float *kernel, int batch, int in_h, int in_w, int channels, int out_h,
                                int out_w, int pad_x0, int pad_y0, const float *noise, float noise_weight, int noise_per_sample,
                                const float *bias, int act, cudaStream_t st) {
  using namespace fircl;
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr || (channels % kCB) != 0 || ((uintptr_t)x & 15) != 0) return cudaErrorNotSupported;
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)in_w, (cuuint64_t)in_h, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)channels * 2, (cuuint64_t)in_w * channels * 2, (cuuint64_t)in_h * in_w * channels * 2};
  const cuuint32_t box[4] = {kCB, kIW, kIH, 1}, estr[4] = {1, 1, 1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorNotSupported;
  Params p;
  p.B = batch, p.C = channels, p.out_h = out_h, p.out_w = out_w, p.pad_x0 = pad_x0, p.pad_y0 = pad_y0, p.act = act;
  p.noise_bstride = noise_per_sample ? out_h * out_w : 0, p.noise_weight = noise_weight;
  p.tiles_x = (out_w + kTW - 1) / kTW, p.tiles_y = (out_h + kTH - 1) / kTH, p.cblocks = channels / kCB;
  p.n_tiles = (long)batch * p.cblocks * p.tiles_y * p.tiles_x;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(p.n_tiles < (long)sms * kCtasPerSm ? p.n_tiles : (long)sms * kCtasPerSm);
  cudaError_t e = cudaFuncSetAttribute(blur4_cl_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return e;
  blur4_cl_tma_kernel<<<grid, kThreads, kSmemBytes, st>>>(map, (uint16_t *)out, kernel, noise, bias, p);
  return cudaGetLastError();
}

}  // namespace hav
