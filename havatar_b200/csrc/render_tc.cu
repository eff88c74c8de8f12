// Tensor-core fused render kernel (HAV_PREC_FP16 / HAV_PREC_BF16): tcgen05.mma with fp32 accumulators in
// TMEM for the three dense layers, everything else on the CUDA cores of the same persistent CTA.
//
// One CTA per SM, two self-contained warpgroups (WG) of 128 threads.  A WG owns a block of 128 rays and
// marches it front to back; each march step is one M=128 tile (row m = ray m at sample s):
//
//   thread-per-row : depth -> o+d*z -> 2-bone skinning warp -> bilinear tap set of both planes -> PE (fp16, smem)
//   warp-cooperative: bi-plane gather, 8 lanes per 128-byte channels-last texel, HFMA2 blend -> A operand (smem)
//   tcgen05.mma     : L0 [128 x 192] x [192 x 128]      (K = 128 features | 48 PE | bias column | pad)
//   thread-per-row : TMEM -> relu -> fp16 -> smem (A operand of the next layer), same for L1 (K = 128 + bias)
//   tcgen05.mma     : head [128 x 144] x [144 x 80]     (64 rgb-features | sigma | 3 rgb pre-composed | pad)
//   thread-per-row : alpha composite straight out of TMEM (utils/nerf_util.py:28-73), registers only
//
// While one WG waits on its MMAs the other one runs its CUDA-core phases.  The weights live in shared memory
// for the lifetime of the CTA in the canonical no-swizzle K-major UMMA layout ([K/8][rows][8]); no per-sample
// tensor ever touches HBM.  fc_rgb (64 -> 3) has no activation in front of it (model/nerf_model.py:110-115),
// so it is folded into the head GEMM as three pre-multiplied columns (fc_rgb.weight @ fc_rgbFeat.weight).
#include "tc_common.cuh"

namespace hav {
namespace tc {

// ------------------------------------------------------------------------------------------------
// packing kernels (run once per call, microseconds)
// ------------------------------------------------------------------------------------------------
template <bool kBF16>
__device__ __forceinline__ uint16_t to16(float v) {
  if (kBF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v));
}
// HAV_RENDER_CHECK_RANGE: an operand that does not fit fp16 (it would become inf) is reported in bit 0 of *status
template <bool kBF16>
__device__ __forceinline__ void range_note(float v, int32_t *status) {
  if (!kBF16 && status != nullptr && fabsf(v) > 65504.0f) atomicOr(status, 1);
}

// Weight image = the exact bytes of the shared-memory weight region: three K-major matrices in [K/8][rows][8]
// order.  Internal K order of L0: 0..63 plane-0 channels, 64..127 plane-1 channels (the reference interleaves
// them as 2c+plane, model/nerf_model.py:99), 128..175 PE, 176 bias.
template <bool kBF16, bool kLo = false>
__global__ void pack_mlp_16_kernel(const float *__restrict__ w0, const float *__restrict__ b0,
                                   const float *__restrict__ w1, const float *__restrict__ b1,
                                   const float *__restrict__ wa, const float *__restrict__ ba,
                                   const float *__restrict__ wf, const float *__restrict__ bf,
                                   const float *__restrict__ wr, const float *__restrict__ br, uint16_t *img, int32_t *status) {
  const int total = kWImgBytes / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int byte = i * 2;
    float v = 0.0f;
    if (byte < kW1Off) {
      int e = i, chunk = e / (128 * 8), n = (e / 8) % 128, k = chunk * 8 + e % 8;
      if (k < 64) v = w0[n * kIn + 2 * k];
      else if (k < 128) v = w0[n * kIn + 2 * (k - 64) + 1];
      else if (k < kIn) v = w0[n * kIn + k];
      else if (k == kIn) v = b0[n];
    } else if (byte < kWHOff) {
      int e = i - kW1Off / 2, chunk = e / (128 * 8), n = (e / 8) % 128, k = chunk * 8 + e % 8;
      if (k < kHid) v = w1[n * kHid + k];
      else if (k == kHid) v = b1[n];
    } else {
      int e = i - kWHOff / 2, chunk = e / (kNH * 8), n = (e / 8) % kNH, k = chunk * 8 + e % 8;
      if (n < kRgbFeat) {
        if (k < kHid) v = wf[n * kHid + k];
        else if (k == kHid) v = bf[n];
      } else if (n == kRgbFeat) {
        if (k < kHid) v = wa[k];
        else if (k == kHid) v = ba[0];
      } else if (n < kRgbFeat + 4) {   // fc_rgb o fc_rgbFeat
        const int j = n - kRgbFeat - 1;
        if (k < kHid) {
          for (int c = 0; c < kRgbFeat; ++c) v = fmaf(wr[j * kRgbFeat + c], wf[c * kHid + k], v);
        } else if (k == kHid) {
          v = br[j];
          for (int c = 0; c < kRgbFeat; ++c) v = fmaf(wr[j * kRgbFeat + c], bf[c], v);
        }
      }
    }
    range_note<kBF16>(v, status);
    if (kLo) v = v - __half2float(__float2half_rn(v));   // split precision: the residual of the fp16 rounding (fp16 mode only)
    img[i] = to16<kBF16>(v);
  }
}

// planes [2,B,64,H,W] fp32 (model/nerf_model.py:85) -> [2B][H+3][W+3][64] 16-bit channels-last with a zero
// border (1 texel before, 2 after): one texel = one 128-byte line, and F.grid_sample's padding_mode='zeros'
// (utils/util.py:404) becomes a plain in-bounds read.  One block per (plane*B+b, y) row.
template <bool kBF16>
__global__ void __launch_bounds__(256) pack_planes_kernel(const float *__restrict__ planes, uint16_t *__restrict__ out,
                                                          int H, int W, int32_t *status) {
  extern __shared__ float tile[];   // [64][W+1]
  const int img = blockIdx.y, y = blockIdx.x;
  const int Hp = H + kPadLo + kPadHi, Wp = W + kPadLo + kPadHi;
  const float *src = planes + (size_t)img * kPlaneC * H * W + (size_t)y * W;
  for (int i = threadIdx.x; i < kPlaneC * W; i += blockDim.x) {
    int c = i / W, x = i % W;
    tile[c * (W + 1) + x] = __ldg(src + (size_t)c * H * W + x);
  }
  __syncthreads();
  uint16_t *dst = out + (((size_t)img * Hp + (y + kPadLo)) * Wp + kPadLo) * kPlaneC;
  for (int i = threadIdx.x; i < kPlaneC * W; i += blockDim.x) {
    int x = i / kPlaneC, c = i % kPlaneC;
    range_note<kBF16>(tile[c * (W + 1) + x], status);
    dst[(size_t)x * kPlaneC + c] = to16<kBF16>(tile[c * (W + 1) + x]);
  }
}

// ------------------------------------------------------------------------------------------------
// the render kernel
// ------------------------------------------------------------------------------------------------
// bilinear tap base + fractional weights on the zero-bordered plane (see pack_planes_kernel)
__device__ __forceinline__ void plane_taps(float gx, float gy, int H, int W, int img, int &off, float &wx, float &wy) {
  float ix = fminf(fmaxf(unnorm(gx, W), -1.0f), (float)W);
  float iy = fminf(fmaxf(unnorm(gy, H), -1.0f), (float)H);
  float x0f = floorf(ix), y0f = floorf(iy);
  wx = ix - x0f, wy = iy - y0f;
  const int Hp = H + kPadLo + kPadHi, Wp = W + kPadLo + kPadHi;
  off = (img * Hp + ((int)y0f + kPadLo)) * Wp + ((int)x0f + kPadLo);
}

template <bool kBF16, bool kTS>
__global__ void __launch_bounds__(kThreads, 1) render_tc_kernel(const RenderDev P, int num_ray_blocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp = t >> 5, lane = tid & 31;
  const uint32_t smem_base = smem_u32(smem);
  uint8_t *Abuf = smem + kSmemA + wg * kABytes;
  const uint32_t A_addr = smem_base + kSmemA + wg * kABytes;
  Stage *stage = reinterpret_cast<Stage *>(smem + kSmemStage + wg * kStageBytes);
  const uint32_t bar = smem_base + kSmemBar + wg * 8;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSmemBar + 32);

  // ---- one-time setup: TMEM, barriers, weights, the constant bias column of both A buffers ----
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kSmemBar + 32),
                 "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    mbar_init(smem_base + kSmemBar, 1);
    mbar_init(smem_base + kSmemBar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.wimg);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = tid; i < kWImgBytes / 16; i += kThreads) dst[i] = __ldg(src + i);
    const uint16_t one = kBF16 ? 0x3F80 : 0x3C00;
    *reinterpret_cast<uint4 *>(Abuf + kOnesChunk * kChunkA + t * 16) = make_uint4(one, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(Abuf + (kOnesChunk + 1) * kChunkA + t * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc0 = tmem_base + wg * 256, tm_acc1 = tm_acc0 + 128;       // column offsets
  const uint32_t tm_lane = (uint32_t)(warp * 32) << 16;                        // this warp's 32 TMEM lanes
  uint32_t phase = 0;

  constexpr uint32_t kIdesc128 = instr_desc(128, kBF16), kIdescH = instr_desc(kNH, kBF16);
  const uint32_t W0_addr = smem_base + kW0Off, W1_addr = smem_base + kW1Off, WH_addr = smem_base + kWHOff;
  const int img_stride_b = P.B;   // plane p of frame b is image p*B + b
  const int Wp = P.PW + kPadLo + kPadHi;
  const uint4 *planes = reinterpret_cast<const uint4 *>(P.planes_cl);

  for (int rb = blockIdx.x * kWGs + wg; rb < num_ray_blocks; rb += gridDim.x * kWGs) {
    const int g = rb * kRaysPerBlock + t;
    const Ray ray = load_ray(P, g);
    const int gi = ray.valid ? g : 0;
    float Tm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Tm[i] = __ldg(P.invT + (size_t)ray.b * 12 + i);
    const int slot = blockIdx.x * kWGs + wg;
    float *zcol = P.zbuf + (size_t)slot * P.Sf * kRaysPerBlock + t;
    float *wcol = P.wbuf + (size_t)slot * P.Sc * kRaysPerBlock + t;
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
      float sums[kOut];
#pragma unroll
      for (int c = 0; c < kOut; ++c) sums[c] = 0.0f;
      float z_cur = pass == 0 ? coarse_z(P, ray, gi, 0) : zcol[0];
      float dist_prev = 0.0f;
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        // ---- row thread: depth, point, skinning warp (nerf_trainer.py:129-146, Skinning_Field.py:70-98)
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
        for (int j = 0; j < 3; ++j) p[j] = fmaf(ray.d[j], z, ray.o[j]);
        skin_warp(P, Tm, p, pc);
        // ---- tap descriptors for the cooperative gather (util.py:359-406)
        {
          float qx = pc[0] * P.ps[0] + P.pt[0], qy = pc[1] * P.ps[1] + P.pt[1], qz = pc[2] * P.ps[2] + P.pt[2];
          Stage st;
          plane_taps(qx, qy, P.PH, P.PW, ray.b, st.off0, st.wx0, st.wy0);
          plane_taps(qz, qy, P.PH, P.PW, img_stride_b + ray.b, st.off1, st.wx1, st.wy1);
          st.pad0 = st.pad1 = 0;
          reinterpret_cast<uint4 *>(stage + t)[0] = make_uint4(st.off0, st.off1, __float_as_uint(st.wx0), __float_as_uint(st.wy0));
          reinterpret_cast<uint2 *>(stage + t)[2] = make_uint2(__float_as_uint(st.wx1), __float_as_uint(st.wy1));
        }
        // ---- positional encoding -> A chunks 16..21 (embedder.py:32-61; order [f][sin|cos][xyz])
        {
          float sn[3], cn[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) sincosf(pc[j], &sn[j], &cn[j]);
          uint32_t pk[24];
#pragma unroll
          for (int f = 0; f < kFreqs; ++f) {
            // 6 values of this octave: sin x,y,z, cos x,y,z -> packed pairs (sx,sy) (sz,cx) (cy,cz)
            pk[f * 3 + 0] = pack2<kBF16>(sn[0], sn[1]);
            pk[f * 3 + 1] = pack2<kBF16>(sn[2], cn[0]);
            pk[f * 3 + 2] = pack2<kBF16>(cn[1], cn[2]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {   // double the angle
              float s2 = 2.0f * sn[j] * cn[j], c2 = fmaf(-2.0f * sn[j], sn[j], 1.0f);
              sn[j] = s2, cn[j] = c2;
            }
          }
#pragma unroll
          for (int c = 0; c < 6; ++c)
            *reinterpret_cast<uint4 *>(Abuf + (16 + c) * kChunkA + t * 16) =
                make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        }
        bar_wg(wg);
        // ---- cooperative gather: 16 lanes per row (2 planes x 8 channel octets), 2 rows per step
        {
          const int sub = lane >> 4, plane = (lane >> 3) & 1, oct = lane & 7;
#pragma unroll 2
          for (int it = 0; it < 16; ++it) {
            const int row = warp * 32 + it * 2 + sub;
            const Stage *sp = stage + row;
            const int off = plane ? sp->off1 : sp->off0;
            const float wx = plane ? sp->wx1 : sp->wx0, wy = plane ? sp->wy1 : sp->wy0;
            const uint4 *tp = planes + (size_t)off * 8 + oct;
            const uint4 t00 = __ldg(tp), t01 = __ldg(tp + 8), t10 = __ldg(tp + (size_t)Wp * 8), t11 = __ldg(tp + (size_t)Wp * 8 + 8);
            const float ux = 1.0f - wx, uy = 1.0f - wy;
            const uint32_t w00 = pack2<kBF16>(ux * uy, ux * uy), w01 = pack2<kBF16>(wx * uy, wx * uy);
            const uint32_t w10 = pack2<kBF16>(ux * wy, ux * wy), w11 = pack2<kBF16>(wx * wy, wx * wy);
            uint4 r;
            r.x = fma2<kBF16>(t11.x, w11, fma2<kBF16>(t10.x, w10, fma2<kBF16>(t01.x, w01, mul2<kBF16>(t00.x, w00))));
            r.y = fma2<kBF16>(t11.y, w11, fma2<kBF16>(t10.y, w10, fma2<kBF16>(t01.y, w01, mul2<kBF16>(t00.y, w00))));
            r.z = fma2<kBF16>(t11.z, w11, fma2<kBF16>(t10.z, w10, fma2<kBF16>(t01.z, w01, mul2<kBF16>(t00.z, w00))));
            r.w = fma2<kBF16>(t11.w, w11, fma2<kBF16>(t10.w, w10, fma2<kBF16>(t01.w, w01, mul2<kBF16>(t00.w, w00))));
            *reinterpret_cast<uint4 *>(Abuf + (plane * 8 + oct) * kChunkA + row * 16) = r;
          }
        }
        fence_async_smem();
        bar_wg(wg);
        // ---- L0: [128 x 192] x [192 x 128] -> acc0
        if (t == 0) {
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kK0 / 16; ++k)
            umma_ss(tm_acc0, smem_desc(A_addr + 2 * k * kChunkA, kChunkA, 128), smem_desc(W0_addr + 2 * k * kChunkB, kChunkB, 128),
                    kIdesc128, k > 0);
          umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue 0: relu -> 16 bit -> A operand of L1 (model/nerf_model.py:105-106).  TS: written back over
        //      the accumulator's own columns (two values per column) + the bias column; SS: smem chunks 0..15
        hidden_epilogue<kBF16, kTS>(tm_acc0 + tm_lane, Abuf, t);
        bar_wg(wg);
        // ---- L1: [128 x 128 (+bias)] x [.. x 128] -> acc1
        if (t == 0) {
          tc_fence_after();
          issue_hidden<kTS>(tm_acc1, tm_acc0, A_addr, W1_addr, kChunkB, kIdesc128);
          umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        hidden_epilogue<kBF16, kTS>(tm_acc1 + tm_lane, Abuf, t);
        bar_wg(wg);
        // ---- head: [128 x 128 (+bias)] x [.. x 80] -> acc0 cols 0..79
        if (t == 0) {
          tc_fence_after();
          issue_hidden<kTS>(tm_acc0, tm_acc1, A_addr, WH_addr, kChunkBH, kIdescH);
          umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- composite (utils/nerf_util.py:28-73): cols 64 = sigma, 65..67 = rgb logits, 0..63 = features
        {
          uint32_t h[4];
          HAV_TMEM_LD4(h, tm_acc0 + tm_lane + kRgbFeat);
          tmem_wait_ld();
          const float nz = noise != nullptr ? __ldg(noise + (size_t)gi * S + s) : 0.0f;
          const float w = cs.step(__uint_as_float(h[0]), nz, dist * ray.dnorm, z);
          if (pass == 0 && npass == 2) wcol[s * kRaysPerBlock] = w;
#pragma unroll
          for (int j = 0; j < 3; ++j) sums[j] = fmaf(w, sigmoidf_exact(__uint_as_float(h[1 + j])), sums[j]);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t r[32];
            HAV_TMEM_LD32(r, tm_acc0 + tm_lane + q * 32);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) sums[3 + q * 32 + c] = fmaf(w, __uint_as_float(r[c]), sums[3 + q * 32 + c]);
          }
        }
        tc_fence_before();   // the next tile's L0 overwrites acc0 after two more barriers
      }
      // ---- write the ray (utils/nerf_util.py:62-71)
      if (ray.valid) {
        float *rgb = (pass == 0 ? P.rgb_c : P.rgb_f) + (size_t)g * kOut;
#pragma unroll
        for (int c = 0; c < kOut; ++c) {
          float v = sums[c];
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
        sample_pdf_merge(zc, P.Sc, P.nfine, wcol, kRaysPerBlock, P.u_rand != nullptr ? P.u_rand + (size_t)gi * P.nfine : nullptr,
                         zcol,
                         (ray.valid && P.pdf_inds != nullptr) ? P.pdf_inds + (size_t)g * P.nfine : nullptr);
        if (ray.valid && P.z_fine != nullptr)
          for (int j = 0; j < P.Sf; ++j) P.z_fine[(size_t)g * P.Sf + j] = zcol[j * kRaysPerBlock];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
uint64_t tc_weight_image_bytes() { return tc::kWImgBytes; }

uint64_t tc_planes_bytes(int nimg, int H, int W) {
  return (uint64_t)nimg * (H + tc::kPadLo + tc::kPadHi) * (W + tc::kPadLo + tc::kPadHi) * kPlaneC * 2;
}

static int sm_count() {   // of the CURRENT device (not cached: a process may drive several GPUs)
  int n = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  return n;
}

int tc_num_ctas(int num_ray_blocks) {
  int want = (num_ray_blocks + tc::kWGs - 1) / tc::kWGs;
  int sms = sm_count();
  return want < sms ? (want > 0 ? want : 1) : sms;
}
int tc_scratch_slots(int num_ray_blocks) { return tc_num_ctas(num_ray_blocks) * tc::kWGs; }

void launch_pack_mlp_16(const hav_render_args *a, uint8_t *wimg, cudaStream_t st, int32_t *status) {
  if (a->precision == HAV_PREC_BF16)
    tc::pack_mlp_16_kernel<true><<<54, 256, 0, st>>>(a->w0, a->b0, a->w1, a->b1, a->w_alpha, a->b_alpha, a->w_feat, a->b_feat,
                                                      a->w_rgb, a->b_rgb, (uint16_t *)wimg, status);
  else
    tc::pack_mlp_16_kernel<false><<<54, 256, 0, st>>>(a->w0, a->b0, a->w1, a->b1, a->w_alpha, a->b_alpha, a->w_feat, a->b_feat,
                                                       a->w_rgb, a->b_rgb, (uint16_t *)wimg, status);
}

void launch_pack_mlp_16_lo(const hav_render_args *a, uint8_t *wimg_lo, cudaStream_t st) {
  tc::pack_mlp_16_kernel<false, true><<<54, 256, 0, st>>>(a->w0, a->b0, a->w1, a->b1, a->w_alpha, a->b_alpha, a->w_feat, a->b_feat,
                                                           a->w_rgb, a->b_rgb, (uint16_t *)wimg_lo, nullptr);
}

namespace tc {
// planes [2,B,64,H,W] fp32 -> [2B][H+3][W+3][64] fp32 channels-last with the same zero border as the 16-bit layout (split mode)
__global__ void __launch_bounds__(256) pack_planes_f32_kernel(const float *__restrict__ planes, float *__restrict__ out, int H, int W) {
  extern __shared__ float tile[];   // [64][W+1]
  const int img = blockIdx.y, y = blockIdx.x;
  const int Hp = H + kPadLo + kPadHi, Wp = W + kPadLo + kPadHi;
  const float *src = planes + (size_t)img * kPlaneC * H * W + (size_t)y * W;
  for (int i = threadIdx.x; i < kPlaneC * W; i += blockDim.x) {
    int c = i / W, x = i % W;
    tile[c * (W + 1) + x] = __ldg(src + (size_t)c * H * W + x);
  }
  __syncthreads();
  float *dst = out + (((size_t)img * Hp + (y + kPadLo)) * Wp + kPadLo) * kPlaneC;
  for (int i = threadIdx.x; i < kPlaneC * W; i += blockDim.x) {
    int x = i / kPlaneC, c = i % kPlaneC;
    dst[(size_t)x * kPlaneC + c] = tile[c * (W + 1) + x];
  }
}
}  // namespace tc

cudaError_t launch_pack_planes_f32(const float *planes, float *out, int nimg, int H, int W, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, 2 * tc_planes_bytes(nimg, H, W), st);
  if (e != cudaSuccess) return e;
  const size_t smem = (size_t)kPlaneC * (W + 1) * sizeof(float);
  e = cudaFuncSetAttribute(tc::pack_planes_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc::pack_planes_f32_kernel<<<dim3(H, nimg), 256, smem, st>>>(planes, out, H, W);
  return cudaGetLastError();
}

cudaError_t launch_pack_planes_16(const float *planes, uint16_t *out, int nimg, int H, int W, bool bf16, cudaStream_t st,
                                  int32_t *status) {
  cudaError_t e = cudaMemsetAsync(out, 0, tc_planes_bytes(nimg, H, W), st);
  if (e != cudaSuccess) return e;
  const size_t smem = (size_t)kPlaneC * (W + 1) * sizeof(float);
  dim3 grid(H, nimg);
  if (bf16) {
    e = cudaFuncSetAttribute(tc::pack_planes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc::pack_planes_kernel<true><<<grid, 256, smem, st>>>(planes, out, H, W, status);
  } else {
    e = cudaFuncSetAttribute(tc::pack_planes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc::pack_planes_kernel<false><<<grid, 256, smem, st>>>(planes, out, H, W, status);
  }
  return cudaGetLastError();
}

template <bool kBF16, bool kTS>
static cudaError_t launch_tc(const RenderDev &P, int num_ray_blocks, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tc::render_tc_kernel<kBF16, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  if (e != cudaSuccess) return e;
  tc::render_tc_kernel<kBF16, kTS><<<tc_num_ctas(num_ray_blocks), tc::kThreads, tc::kSmemBytes, st>>>(P, num_ray_blocks);
  return cudaGetLastError();
}

cudaError_t launch_render_16(const RenderDev &P, int num_ray_blocks, bool bf16, cudaStream_t st) {
  static const bool v1 = getenv("HAV_TC_V1") != nullptr;   // debugging aids: the v1 kernel (no warp specialisation) ...
  static const bool ss = getenv("HAV_TC_SS") != nullptr;   // ... and v1 with hidden activations through smem instead of TMEM
  if (!v1 && !ss) return launch_render_16_v2(P, num_ray_blocks, bf16, st);
  if (ss) return bf16 ? launch_tc<true, false>(P, num_ray_blocks, st) : launch_tc<false, false>(P, num_ray_blocks, st);
  return bf16 ? launch_tc<true, true>(P, num_ray_blocks, st) : launch_tc<false, true>(P, num_ray_blocks, st);
}

}  // namespace hav
