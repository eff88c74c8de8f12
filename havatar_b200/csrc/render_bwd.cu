// Backward of the fused render (gradients of Trainer.predict_and_render_radiance, model/nerf_trainer.py:120-201, which the
// reference obtains from autograd over ~140 ATen ops per chunk and pass).  Three kernels:
//
//   render_bwd_kernel : one warpgroup marches a block of 128 rays front to back exactly like the forward kernel
//       (render_tc.cu), one M = 128 tile per sample index.  Per tile it RECOMPUTES the forward (gather -> L0 -> L1 -> head on
//       tcgen05, hidden activations in TMEM), runs the composite backward per ray in registers (suffix sums come from the
//       forward's own outputs minus a running prefix, so there is no back-to-front pass and no per-sample state in HBM),
//       then the three data-gradient GEMMs  d_h2 = d_out . W_head,  d_h1 = (d_h2 * relu') . W1,  d_x = (d_h1 * relu') . W0
//       on tcgen05 with the gradient rows as the A operand in TMEM and the SAME shared-memory weight image as the B operand,
//       read through an MN-major descriptor (B'[in][out] = B[out][in]: no transposed copy of the weights).  The row thread
//       then scatters d_x: bi-plane texel gradients (vector red.add into a channels-last fp32 image), the plane-coordinate
//       and positional-encoding gradients -> d(canonical point) -> skinning-weight gradients (trilinear scatter).
//       Every tile also leaves its six 16-bit operand images (x, h1, h2, d_out, d_pre1, d_pre0) in HBM, chunk-major
//       [F/8][128 rows][8] -- the byte image of a K-major A operand, which read with the SAMPLE index as K is the canonical
//       MN-major operand of the weight-gradient GEMM.
//   wgrad_kernel      : dW = sum over tiles of d_pre^T . act, a split-K tcgen05 GEMM over those images (both operands
//       MN-major from shared memory, bulk-copied as they lie); the bias gradients fall out of the constant-one column.
//   finalize kernels  : back to the reference's parameter layouts (un-permute L0's plane interleave, un-compose
//       fc_rgb o fc_rgbFeat), channels-last -> NCHW for the plane gradients.
//
// 16-bit gradient operands carry a power-of-two loss scale chosen on the device from max|upstream gradient|.
#include "tc_common.cuh"

namespace hav {
namespace bwd {

using namespace tc;

constexpr int kThr = 128;
constexpr int kSmA = kWImgBytes;
constexpr int kSmStage = kSmA + kABytes;
constexpr int kSmD = kSmStage + kStageBytes;      // d(feature)/d(ix), d(feature)/d(iy) of every row: 32 chunks like the A operand
constexpr int kDChunks = 32;
constexpr int kSmBarB = kSmD + kDChunks * kChunkA;
constexpr int kSmWin = kSmBarB + 64;              // per-warp min / max of the tap coordinates: [4 warps][8] ints
constexpr int kSmBytesB = kSmWin + 4 * 8 * 4;
static_assert(kSmBytesB <= 232448, "shared memory budget");
constexpr int kWinTexels = 128;                   // texel window of the tensor-core scatter (one TMEM lane per texel)
constexpr int kC0 = 0, kC1 = 128, kC2 = 256, kC3 = 384;   // TMEM column blocks

// per-tile operand images in HBM
constexpr int kChunk = 2048;
constexpr int kXCh = 22, kHCh = 16, kOCh = 10;
constexpr int kOffX = 0;
constexpr int kOffH1 = kOffX + kXCh * kChunk;
constexpr int kOffH2 = kOffH1 + kHCh * kChunk;
constexpr int kOffDO = kOffH2 + kHCh * kChunk;
constexpr int kOffD1 = kOffDO + kOCh * kChunk;
constexpr int kOffD0 = kOffD1 + kHCh * kChunk;
constexpr int kTileBytes = kOffD0 + kHCh * kChunk;   // 196608

// weight-gradient accumulator image [128 rows][kDWCols] fp32: L0 | L1 | head
constexpr int kDW0 = 0, kDW1 = 192, kDWH = 336, kDWCols = 480;

struct BwdDev {
  const float *g_rgb[2], *g_depth[2], *g_acc[2];   // [pass]: coarse, fine (NULL = zero)
  const float *scale;                               // device: {scale, 1/scale}
  float *gplanes_cl;                                // [2B][H+3][W+3][64] fp32, same texel indexing as the packed planes
  float *gwvol;                                     // [2][D][H][W]
  uint8_t *dump;                                    // [tiles][kTileBytes]
  int swap_mn;                                      // debugging aid: swap LBO / SBO of the MN-major descriptors
};

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <bool kBF16>
__device__ __forceinline__ void unpack2(uint32_t v, float &lo, float &hi) {
  if (kBF16) {
    lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xFFFF0000u);
  } else {
    float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&v));
    lo = f.x, hi = f.y;
  }
}
template <bool kBF16>
__device__ __forceinline__ uint32_t sub2(uint32_t a, uint32_t b) {
  uint32_t d;
  if (kBF16) asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
#define HAV_TMEM_LD16(r, taddr)                                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),        \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])   \
               : "r"(taddr))

// MN-major no-swizzle descriptor: core matrix = 8 K-rows of 16 bytes (8 contiguous MN elements); SBO = distance between
// 8-element MN groups, LBO = distance between 8-row K groups (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr, uint32_t k_group_bytes, uint32_t mn_group_bytes, int swap) {
  return swap ? smem_desc(addr, mn_group_bytes, k_group_bytes) : smem_desc(addr, k_group_bytes, mn_group_bytes);
}
constexpr uint32_t kBMajorMN = 1u << 16, kAMajorMN = 1u << 15;

// F.grid_sample(4-D, zeros, align_corners) tap base on the zero-bordered packed plane + whether the coordinate is inside the
// range where the interpolant depends on it (outside every corner is padding: zero gradient, as in ATen's backward)
__device__ __forceinline__ void plane_taps_b(float gx, float gy, int H, int W, int img, int &off, float &wx, float &wy, bool &in,
                                             int &px, int &py) {
  const float ux = unnorm(gx, W), uy = unnorm(gy, H);
  in = ux > -1.0f && ux < (float)W && uy > -1.0f && uy < (float)H;
  float ix = fminf(fmaxf(ux, -1.0f), (float)W), iy = fminf(fmaxf(uy, -1.0f), (float)H);
  float x0f = floorf(ix), y0f = floorf(iy);
  wx = ix - x0f, wy = iy - y0f;
  const int Hp = H + kPadLo + kPadHi, Wp = W + kPadLo + kPadHi;
  px = (int)x0f + kPadLo, py = img * Hp + ((int)y0f + kPadLo);   // column / row of tap (y0,x0) in the stacked padded images
  off = py * Wp + px;
}

// backward of trilinear_border (render_common.cuh) with respect to the volume: the same corners and weights, scattered
__device__ __forceinline__ void trilinear_border_scatter(float *__restrict__ gvol, int D, int H, int W, float x, float y, float z,
                                                         float g) {
  float ix = fminf(fmaxf(unnorm(x, W), 0.0f), (float)(W - 1));
  float iy = fminf(fmaxf(unnorm(y, H), 0.0f), (float)(H - 1));
  float iz = fminf(fmaxf(unnorm(z, D), 0.0f), (float)(D - 1));
  float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
  int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
  float wx1 = ix - x0f, wy1 = iy - y0f, wz1 = iz - z0f;
  float wx0 = (x0f + 1.0f) - ix, wy0 = (y0f + 1.0f) - iy, wz0 = (z0f + 1.0f) - iz;
  bool xe = x0 + 1 <= W - 1, ye = y0 + 1 <= H - 1, ze = z0 + 1 <= D - 1;
  float *p0 = gvol + ((size_t)z0 * H) * W, *p1 = p0 + (size_t)H * W;
  const int r0 = y0 * W + x0, r1 = r0 + W;
  atomicAdd(p0 + r0, g * ((wx0 * wy0) * wz0));
  if (xe) atomicAdd(p0 + r0 + 1, g * ((wx1 * wy0) * wz0));
  if (ye) atomicAdd(p0 + r1, g * ((wx0 * wy1) * wz0));
  if (xe && ye) atomicAdd(p0 + r1 + 1, g * ((wx1 * wy1) * wz0));
  if (ze) {
    atomicAdd(p1 + r0, g * ((wx0 * wy0) * wz1));
    if (xe) atomicAdd(p1 + r0 + 1, g * ((wx1 * wy0) * wz1));
    if (ye) atomicAdd(p1 + r1, g * ((wx0 * wy1) * wz1));
    if (xe && ye) atomicAdd(p1 + r1 + 1, g * ((wx1 * wy1) * wz1));
  }
}

// relu (kRelu) or relu-mask by the packed activations at tm_mask (!kRelu), 16-bit pack of one 128-column fp32 accumulator
// row back over its own first 64 columns (the A operand of the next GEMM) + a copy into the tile's operand image in HBM
template <bool kBF16, bool kRelu>
__device__ __forceinline__ void epilogue_pack_dump(uint32_t tm_row, uint32_t tm_mask, uint8_t *dst, int t) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[32];
    HAV_TMEM_LD32(r, tm_row + q * 32);
    uint32_t m[16];
    if (!kRelu) HAV_TMEM_LD16(m, tm_mask + q * 16);
    tmem_wait_ld();
    uint32_t v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (kRelu) {
        v[c] = pack_relu<kBF16>(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1]));
      } else {
        const float lo = (m[c] & 0x0000FFFFu) ? __uint_as_float(r[2 * c]) : 0.0f;
        const float hi = (m[c] & 0xFFFF0000u) ? __uint_as_float(r[2 * c + 1]) : 0.0f;
        v[c] = pack2<kBF16>(lo, hi);
      }
    }
    HAV_TMEM_ST16(tm_row + q * 16, v);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4 *>(dst + (q * 4 + c) * kChunk + t * 16) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  }
  if (kRelu) {   // constant-one bias column of the forward GEMMs
    uint32_t one[8] = {kBF16 ? 0x3F80u : 0x3C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    HAV_TMEM_ST8(tm_row + 64, one);
  }
  tmem_wait_st();
  tc_fence_before();
}

template <bool kBF16>
__global__ void __launch_bounds__(kThr, 1) render_bwd_kernel(const RenderDev P, const BwdDev Q, int num_ray_blocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t smem_base = smem_u32(smem);
  uint8_t *Abuf = smem + kSmA;
  const uint32_t A_addr = smem_base + kSmA;
  Stage *stage = reinterpret_cast<Stage *>(smem + kSmStage);
  uint8_t *Dbuf = smem + kSmD;
  int *win_s = reinterpret_cast<int *>(smem + kSmWin);
  const uint32_t bar = smem_base + kSmBarB;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSmBarB + 32);

  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kSmBarB + 32), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 32) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.wimg);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = t; i < kWImgBytes / 16; i += kThr) dst[i] = __ldg(src + i);
    const uint16_t one = kBF16 ? 0x3F80 : 0x3C00;
    *reinterpret_cast<uint4 *>(Abuf + kOnesChunk * kChunkA + t * 16) = make_uint4(one, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(Abuf + (kOnesChunk + 1) * kChunkA + t * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_lane = (uint32_t)(warp * 32) << 16;
  const uint32_t tmC0 = tmem_base + kC0, tmC1 = tmem_base + kC1, tmC2 = tmem_base + kC2, tmC3 = tmem_base + kC3;
  uint32_t phase = 0;

  constexpr uint32_t kIdesc128 = instr_desc(128, kBF16), kIdescH = instr_desc(kNH, kBF16);
  constexpr uint32_t kIdescD128 = instr_desc(128, kBF16) | kBMajorMN, kIdescDX = instr_desc(kIn, kBF16) | kBMajorMN;
  constexpr uint32_t kIdescSc = instr_desc(kPlaneC, kBF16) | kAMajorMN | kBMajorMN;
  const uint32_t W0_addr = smem_base + kW0Off, W1_addr = smem_base + kW1Off, WH_addr = smem_base + kWHOff;
  const int Wp = P.PW + kPadLo + kPadHi;
  const uint4 *planes = reinterpret_cast<const uint4 *>(P.planes_cl);
  const float gscale = __ldg(Q.scale), ginv = __ldg(Q.scale + 1);
  const size_t vs = (size_t)P.VD * P.VH * P.VW;
  const int Stot = P.Sc + P.Sf;

#define HAV_MMA_ROUND(body)    \
  if (t == 0) {                \
    tc_fence_after();          \
    body;                      \
    umma_commit(bar);          \
  }                            \
  mbar_wait(bar, phase);       \
  phase ^= 1;                  \
  tc_fence_after();

  for (int rb = blockIdx.x; rb < num_ray_blocks; rb += gridDim.x) {
    const int g = rb * kRaysPerBlock + t;
    const Ray ray = load_ray(P, g);
    const int gi = ray.valid ? g : 0;
    float Tm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Tm[i] = __ldg(P.invT + (size_t)ray.b * 12 + i);
    float bgc[3] = {0.f, 0.f, 0.f};
    if (P.bg != nullptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) bgc[c] = __ldg(P.bg + (size_t)gi * 3 + c);
    }
    const int npass = P.nfine > 0 ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
      const int S = pass == 0 ? P.Sc : P.Sf;
      const float *noise = pass == 0 ? P.noise_c : P.noise_f;
      const float *zf = P.z_fine + (size_t)gi * P.Sf;
      // ---- upstream gradients of this ray and the total  sum_j (dL/dw_j) w_j  from the forward's outputs
      float G[kOut], gd = 0.0f, ga = 0.0f, total = 0.0f;
      {
        const float *rgb_o = (pass == 0 ? P.rgb_c : P.rgb_f) + (size_t)gi * kOut;
        const float acc_o = (pass == 0 ? P.acc_c : P.acc_f)[gi], depth_o = (pass == 0 ? P.depth_c : P.depth_f)[gi];
        const bool on = ray.valid;
        if (on && Q.g_depth[pass] != nullptr) gd = __ldg(Q.g_depth[pass] + gi);
        if (on && Q.g_acc[pass] != nullptr) ga = __ldg(Q.g_acc[pass] + gi);
#pragma unroll
        for (int c = 0; c < kOut; ++c) {
          G[c] = (on && Q.g_rgb[pass] != nullptr) ? __ldg(Q.g_rgb[pass] + (size_t)gi * kOut + c) : 0.0f;
          float v = rgb_o[c];
          if (c < 3 && P.bg != nullptr) {   // rgb += (1 - acc) * bg  (utils/nerf_util.py:70-71)
            v -= (1.0f - acc_o) * bgc[c];
            ga -= G[c] * bgc[c];
          }
          total = fmaf(G[c], v, total);
        }
        total = fmaf(gd, depth_o, total);
        total = fmaf(ga, acc_o, total);
      }
      float Tr = 1.0f, prefix = 0.0f;
      float z_cur = pass == 0 ? coarse_z(P, ray, gi, 0) : zf[0];
      float dist_prev = 0.0f;
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        uint8_t *tile = Q.dump + (size_t)((size_t)rb * Stot + (pass == 0 ? 0 : P.Sc) + s) * kTileBytes;
        // ================= forward recompute (as render_tc.cu) =================
        float z_next = 0.0f, dist;
        if (s + 1 < S) {
          z_next = pass == 0 ? coarse_z(P, ray, gi, s + 1) : zf[s + 1];
          dist = z_next - z_cur;
        } else {
          dist = dist_prev;
        }
        dist_prev = dist;
        const float z = z_cur;
        z_cur = z_next;
        float p[3], p1[3], pc[3], a0, a1, den;
#pragma unroll
        for (int j = 0; j < 3; ++j) p[j] = fmaf(ray.d[j], z, ray.o[j]);
        {   // Skinning_Field.py:70-98 (skin_warp of render_common.cuh with its intermediates kept)
          float q0 = p[0] + Tm[9], q1 = p[1] + Tm[10], q2 = p[2] + Tm[11];
#pragma unroll
          for (int j = 0; j < 3; ++j) p1[j] = q0 * Tm[j] + q1 * Tm[3 + j] + q2 * Tm[6 + j];
          float w0 = trilinear_border(P.wvol, P.VD, P.VH, P.VW, p[0] * P.ss[0] + P.st[0], p[1] * P.ss[1] + P.st[1], p[2] * P.ss[2] + P.st[2]);
          float w1 = trilinear_border(P.wvol + vs, P.VD, P.VH, P.VW, p1[0] * P.ss[0] + P.st[0], p1[1] * P.ss[1] + P.st[1],
                                      p1[2] * P.ss[2] + P.st[2]);
          den = (w0 + w1) + 1e-8f;
          a0 = w0 / den, a1 = w1 / den;
#pragma unroll
          for (int j = 0; j < 3; ++j) pc[j] = a0 * p[j] + a1 * p1[j];
        }
        Stage st;
        bool in0, in1;
        int tpx[2], tpy[2];
        {
          float qx = pc[0] * P.ps[0] + P.pt[0], qy = pc[1] * P.ps[1] + P.pt[1], qz = pc[2] * P.ps[2] + P.pt[2];
          plane_taps_b(qx, qy, P.PH, P.PW, ray.b, st.off0, st.wx0, st.wy0, in0, tpx[0], tpy[0]);
          plane_taps_b(qz, qy, P.PH, P.PW, P.B + ray.b, st.off1, st.wx1, st.wy1, in1, tpx[1], tpy[1]);
          // bounding box of this tile's taps per plane (warp reduction now, combined after the next barrier)
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            const int big = 0x3fffffff;
            const int xmn = __reduce_min_sync(0xffffffffu, ray.valid ? tpx[pl] : big), xmx = __reduce_max_sync(0xffffffffu, ray.valid ? tpx[pl] : -big);
            const int ymn = __reduce_min_sync(0xffffffffu, ray.valid ? tpy[pl] : big), ymx = __reduce_max_sync(0xffffffffu, ray.valid ? tpy[pl] : -big);
            if (lane == 0) *reinterpret_cast<int4 *>(win_s + warp * 8 + pl * 4) = make_int4(xmn, xmx, ymn, ymx);
          }
          reinterpret_cast<uint4 *>(stage + t)[0] = make_uint4(st.off0, st.off1, __float_as_uint(st.wx0), __float_as_uint(st.wy0));
          reinterpret_cast<uint2 *>(stage + t)[2] = make_uint2(__float_as_uint(st.wx1), __float_as_uint(st.wy1));
        }
        float sn0[3], cn0[3];
        {
#pragma unroll
          for (int j = 0; j < 3; ++j) sincosf(pc[j], &sn0[j], &cn0[j]);
          float sn[3] = {sn0[0], sn0[1], sn0[2]}, cn[3] = {cn0[0], cn0[1], cn0[2]};
          uint32_t pk[24];
#pragma unroll
          for (int f = 0; f < kFreqs; ++f) {
            pk[f * 3 + 0] = pack2<kBF16>(sn[0], sn[1]);
            pk[f * 3 + 1] = pack2<kBF16>(sn[2], cn[0]);
            pk[f * 3 + 2] = pack2<kBF16>(cn[1], cn[2]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              float s2 = 2.0f * sn[j] * cn[j], c2 = fmaf(-2.0f * sn[j], sn[j], 1.0f);
              sn[j] = s2, cn[j] = c2;
            }
          }
#pragma unroll
          for (int c = 0; c < 6; ++c)
            *reinterpret_cast<uint4 *>(Abuf + (16 + c) * kChunkA + t * 16) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        }
        bar_wg(0);
        int wxmin[2], wymin[2], wW[2];   // texel window of this tile per plane: origin and row pitch (see the scatter below)
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          int4 b = *reinterpret_cast<const int4 *>(win_s + pl * 4);
#pragma unroll
          for (int wq = 1; wq < 4; ++wq) {
            const int4 o = *reinterpret_cast<const int4 *>(win_s + wq * 8 + pl * 4);
            b.x = min(b.x, o.x), b.y = max(b.y, o.y), b.z = min(b.z, o.z), b.w = max(b.w, o.w);
          }
          wxmin[pl] = b.x, wymin[pl] = b.z;
          wW[pl] = b.y >= b.x ? min(b.y - b.x + 2, kWinTexels) : 1;
        }
        {   // cooperative bi-plane gather: 16 lanes per row (2 planes x 8 channel octets), 2 rows per step
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
            // coordinate derivatives of the interpolant, kept for the backward of this tile (util.py:395-406):
            //   d/d(ix) = (t01 - t00) uy + (t11 - t10) wy ,  d/d(iy) = (t10 - t00) ux + (t11 - t01) wx
            const uint32_t a4[4] = {t00.x, t00.y, t00.z, t00.w}, b4[4] = {t01.x, t01.y, t01.z, t01.w};
            const uint32_t c4[4] = {t10.x, t10.y, t10.z, t10.w}, d4[4] = {t11.x, t11.y, t11.z, t11.w};
            const uint32_t ux2 = pack2<kBF16>(ux, ux), uy2 = pack2<kBF16>(uy, uy), wx2 = pack2<kBF16>(wx, wx), wy2 = pack2<kBF16>(wy, wy);
            uint32_t dx4[4], dy4[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              dx4[h] = fma2<kBF16>(sub2<kBF16>(b4[h], a4[h]), uy2, mul2<kBF16>(sub2<kBF16>(d4[h], c4[h]), wy2));
              dy4[h] = fma2<kBF16>(sub2<kBF16>(c4[h], a4[h]), ux2, mul2<kBF16>(sub2<kBF16>(d4[h], b4[h]), wx2));
            }
            *reinterpret_cast<uint4 *>(Dbuf + ((plane * 8 + oct) * 2) * kChunkA + row * 16) = make_uint4(dx4[0], dx4[1], dx4[2], dx4[3]);
            *reinterpret_cast<uint4 *>(Dbuf + ((plane * 8 + oct) * 2 + 1) * kChunkA + row * 16) = make_uint4(dy4[0], dy4[1], dy4[2], dy4[3]);
          }
        }
        fence_async_smem();
        bar_wg(0);
        // ---- L0 -> C0 (+ the tile's x image to HBM while the MMAs run)
        if (t == 0) {
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kK0 / 16; ++k)
            umma_ss(tmC0, smem_desc(A_addr + 2 * k * kChunkA, kChunkA, 128), smem_desc(W0_addr + 2 * k * kChunkB, kChunkB, 128),
                    kIdesc128, k > 0);
          umma_commit(bar);
        }
#pragma unroll
        for (int c = 0; c < kXCh; ++c)
          *reinterpret_cast<uint4 *>(tile + kOffX + c * kChunk + t * 16) = *reinterpret_cast<const uint4 *>(Abuf + c * kChunkA + t * 16);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        epilogue_pack_dump<kBF16, true>(tmC0 + tm_lane, 0, tile + kOffH1, t);     // h1 = relu(.)  (nerf_model.py:105-106)
        bar_wg(0);
        HAV_MMA_ROUND(issue_hidden<true>(tmC1, tmC0, 0, W1_addr, kChunkB, kIdesc128));
        epilogue_pack_dump<kBF16, true>(tmC1 + tm_lane, 0, tile + kOffH2, t);     // h2
        bar_wg(0);
        HAV_MMA_ROUND(issue_hidden<true>(tmC2, tmC1, 0, WH_addr, kChunkBH, kIdescH));
        // ================= composite backward (utils/nerf_util.py:28-73) =================
        float wgt;
        {
          uint32_t h[4];
          HAV_TMEM_LD4(h, tmC2 + tm_lane + kRgbFeat);
          tmem_wait_ld();
          const float nz = noise != nullptr ? __ldg(noise + (size_t)gi * S + s) : 0.0f;
          const float dn = dist * ray.dnorm;
          const float sin_ = __uint_as_float(h[0]) + nz;
          const float sigma = fmaxf(sin_, 0.0f);
          const float e = expf(-sigma * dn);
          const float alpha = 1.0f - e;
          wgt = alpha * Tr;
          const float one_m = (1.0f - alpha) + 1e-10f;
          float sg[3], dot = 0.0f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            sg[j] = sigmoidf_exact(__uint_as_float(h[1 + j]));
            dot = fmaf(G[j], sg[j], dot);
          }
          const float ws = wgt * gscale;
          uint32_t vlast[8];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t r[32];
            HAV_TMEM_LD32(r, tmC2 + tm_lane + q * 32);
            tmem_wait_ld();
            uint32_t v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              dot = fmaf(G[3 + q * 32 + 2 * c], __uint_as_float(r[2 * c]), dot);
              dot = fmaf(G[3 + q * 32 + 2 * c + 1], __uint_as_float(r[2 * c + 1]), dot);
              v[c] = pack2<kBF16>(G[3 + q * 32 + 2 * c] * ws, G[3 + q * 32 + 2 * c + 1] * ws);   // d f = G * w
            }
            HAV_TMEM_ST16(tmC2 + tm_lane + q * 16, v);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4 *>(tile + kOffDO + (q * 4 + c) * kChunk + t * 16) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
          const float gw = dot + gd * z + ga;              // dL/dw_s
          prefix = fmaf(gw, wgt, prefix);
          const float dalpha = gw * Tr - (total - prefix) / one_m;
          const float dsig = sin_ > 0.0f ? dalpha * dn * e : 0.0f;
          Tr *= one_m;
          float dl[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) dl[j] = G[j] * wgt * sg[j] * (1.0f - sg[j]);   // sigmoid on rgb only (:45-46)
          vlast[0] = pack2<kBF16>(dsig * gscale, dl[0] * gscale);
          vlast[1] = pack2<kBF16>(dl[1] * gscale, dl[2] * gscale);
#pragma unroll
          for (int c = 2; c < 8; ++c) vlast[c] = 0u;
          HAV_TMEM_ST8(tmC2 + tm_lane + 32, vlast);
          *reinterpret_cast<uint4 *>(tile + kOffDO + 8 * kChunk + t * 16) = make_uint4(vlast[0], vlast[1], 0u, 0u);
          *reinterpret_cast<uint4 *>(tile + kOffDO + 9 * kChunk + t * 16) = make_uint4(0u, 0u, 0u, 0u);
          tmem_wait_st();
          tc_fence_before();
        }
        bar_wg(0);
        // ================= data gradients: the weight image read MN-major is W^T =================
        HAV_MMA_ROUND({
          for (int k = 0; k < kNH / 16; ++k)
            umma_ts(tmC3, tmC2 + k * 8, smem_desc_mn(WH_addr + k * 256, 128, kChunkBH, Q.swap_mn), kIdescD128, k > 0);
        });
        epilogue_pack_dump<kBF16, false>(tmC3 + tm_lane, tmC1 + tm_lane, tile + kOffD1, t);   // d_pre1 = d_h2 * [h2 > 0]
        bar_wg(0);
        HAV_MMA_ROUND({
          for (int k = 0; k < kHid / 16; ++k)
            umma_ts(tmC2, tmC3 + k * 8, smem_desc_mn(W1_addr + k * 256, 128, kChunkB, Q.swap_mn), kIdescD128, k > 0);
        });
        epilogue_pack_dump<kBF16, false>(tmC2 + tm_lane, tmC0 + tm_lane, tile + kOffD0, t);   // d_pre0 = d_h1 * [h1 > 0]
        bar_wg(0);
        HAV_MMA_ROUND({
          for (int k = 0; k < kHid / 16; ++k)
            umma_ts(tmC0, tmC2 + k * 8, smem_desc_mn(W0_addr + k * 256, 128, kChunkB, Q.swap_mn), kIdescDX, k > 0);
        });
        // ================= d_x (C0..C1, 176 columns) -> planes, canonical point, skinning weights =================
        {
          float dpc[3] = {0.f, 0.f, 0.f};
          {   // positional encoding, columns 128..175 ordered [f][sin|cos][xyz] (embedder.py:32-61)
            uint32_t r[48];
            HAV_TMEM_LD32(r, tmC0 + tm_lane + 128);
            HAV_TMEM_LD16((r + 32), tmC0 + tm_lane + 160);
            tmem_wait_ld();
            float sn[3] = {sn0[0], sn0[1], sn0[2]}, cn[3] = {cn0[0], cn0[1], cn0[2]};
            float fr = 1.0f;
#pragma unroll
            for (int f = 0; f < kFreqs; ++f) {
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                dpc[j] += fr * (__uint_as_float(r[f * 6 + j]) * cn[j] - __uint_as_float(r[f * 6 + 3 + j]) * sn[j]);   // loss-scaled
                float s2 = 2.0f * sn[j] * cn[j], c2 = fmaf(-2.0f * sn[j], sn[j], 1.0f);
                sn[j] = s2, cn[j] = c2;
              }
              fr *= 2.0f;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) dpc[j] *= ginv;
          }
          // bi-plane features (util.py:359-406).  Coordinate gradients: <d_feat, d(feature)/d(ix, iy)> with the derivatives
          // saved by the gather.  Texel gradients: rows of a tile share most of their texels (neighbouring pixels are a
          // fraction of a texel apart), so instead of 4 taps x 64 channels of global reductions per row and plane the tile is
          // reduced on the tensor core first:  G[texel][ch] = S[texel][row] . d_feat[row][ch], S = bilinear weights of the rows
          // over a window of <= 128 texels around the tile's taps (one TMEM lane per window texel); only window texels are
          // then added to the global image.  Rows whose taps fall outside the window take the direct path.
          int m00[2];
          bool inwin[2];
#pragma unroll
          for (int plane = 0; plane < 2; ++plane) {
            m00[plane] = (tpy[plane] - wymin[plane]) * wW[plane] + (tpx[plane] - wxmin[plane]);
            inwin[plane] = ray.valid && tpx[plane] - wxmin[plane] + 1 < wW[plane] && m00[plane] + wW[plane] + 1 < kWinTexels;
          }
#pragma unroll 1
          for (int plane = 0; plane < 2; ++plane) {
            const int off = plane ? st.off1 : st.off0;
            const float wx = plane ? st.wx1 : st.wx0, wy = plane ? st.wy1 : st.wy0;
            const float ux = 1.0f - wx, uy = 1.0f - wy;
            const float tw[4] = {ux * uy, wx * uy, ux * wy, wx * wy};
            const int toff[4] = {0, 1, Wp, Wp + 1};
            float dfx = 0.0f, dfy = 0.0f;
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {
              uint32_t r[32];
              HAV_TMEM_LD32(r, tmC0 + tm_lane + plane * 64 + q * 32);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint8_t *dp = Dbuf + ((plane * 8 + q * 4 + i) * 2) * kChunkA + t * 16;
                const uint4 vx = *reinterpret_cast<const uint4 *>(dp), vy = *reinterpret_cast<const uint4 *>(dp + kChunkA);
                const uint32_t x4[4] = {vx.x, vx.y, vx.z, vx.w}, y4[4] = {vy.x, vy.y, vy.z, vy.w};
#pragma unroll
                for (int h2 = 0; h2 < 4; ++h2) {
                  float lo, hi;
                  unpack2<kBF16>(x4[h2], lo, hi);
                  dfx = fmaf(__uint_as_float(r[i * 8 + h2 * 2]), lo, dfx);
                  dfx = fmaf(__uint_as_float(r[i * 8 + h2 * 2 + 1]), hi, dfx);
                  unpack2<kBF16>(y4[h2], lo, hi);
                  dfy = fmaf(__uint_as_float(r[i * 8 + h2 * 2]), lo, dfy);
                  dfy = fmaf(__uint_as_float(r[i * 8 + h2 * 2 + 1]), hi, dfy);
                }
                // d_feat (still loss-scaled) as the 16-bit B operand of the scatter GEMM, in the dead X buffer
                *reinterpret_cast<uint4 *>(Abuf + (plane * 8 + q * 4 + i) * kChunkA + t * 16) =
                    make_uint4(pack2<kBF16>(__uint_as_float(r[i * 8]), __uint_as_float(r[i * 8 + 1])),
                               pack2<kBF16>(__uint_as_float(r[i * 8 + 2]), __uint_as_float(r[i * 8 + 3])),
                               pack2<kBF16>(__uint_as_float(r[i * 8 + 4]), __uint_as_float(r[i * 8 + 5])),
                               pack2<kBF16>(__uint_as_float(r[i * 8 + 6]), __uint_as_float(r[i * 8 + 7])));
              }
              if (ray.valid && !inwin[plane]) {   // direct path: w_tap * d_feat straight into the channels-last image
#pragma unroll 1
                for (int tap = 0; tap < 4; ++tap) {
                  const float wv = tw[tap] * ginv;
                  if (wv != 0.0f) {
                    float *gp = Q.gplanes_cl + (size_t)(off + toff[tap]) * kPlaneC + q * 32;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                      red_add_v4(gp + i * 4, __uint_as_float(r[i * 4]) * wv, __uint_as_float(r[i * 4 + 1]) * wv,
                                 __uint_as_float(r[i * 4 + 2]) * wv, __uint_as_float(r[i * 4 + 3]) * wv);
                  }
                }
              }
            }
            if (plane ? in1 : in0) {
              const float gx = dfx * ginv * 0.5f * (float)(P.PW - 1), gy = dfy * ginv * 0.5f * (float)(P.PH - 1);
              // plane 0 is sampled at (qx, qy), plane 1 at (qz, qy)  (util.py:378-381)
              if (plane == 0) dpc[0] = fmaf(gx, P.ps[0], dpc[0]);
              else dpc[2] = fmaf(gx, P.ps[2], dpc[2]);
              dpc[1] = fmaf(gy, P.ps[1], dpc[1]);
            }
          }
          // skinning weights (Skinning_Field.py:85-95): pc = a0 p + a1 p1, a_i = w_i / (w0 + w1 + 1e-8)
          if (ray.valid) {
            float da0 = 0.0f, da1 = 0.0f;
#pragma unroll
            for (int j = 0; j < 3; ++j) da0 = fmaf(dpc[j], p[j], da0), da1 = fmaf(dpc[j], p1[j], da1);
            const float common = a0 * da0 + a1 * da1;
            const float dw0 = (da0 - common) / den, dw1 = (da1 - common) / den;
            trilinear_border_scatter(Q.gwvol, P.VD, P.VH, P.VW, p[0] * P.ss[0] + P.st[0], p[1] * P.ss[1] + P.st[1],
                                     p[2] * P.ss[2] + P.st[2], dw0);
            trilinear_border_scatter(Q.gwvol + vs, P.VD, P.VH, P.VW, p1[0] * P.ss[0] + P.st[0], p1[1] * P.ss[1] + P.st[1],
                                     p1[2] * P.ss[2] + P.st[2], dw1);
          }
          // ---- scatter GEMM.  S lives where the coordinate derivatives were (every thread is done reading them after the
          //      barrier), MN-major: element (texel m, row r) at [m / 8][r][m % 8], so row thread r owns and rewrites its whole
          //      column (16 x 16 bytes per plane) every tile -- no clearing pass.
          bar_wg(0);
#pragma unroll 1
          for (int plane = 0; plane < 2; ++plane) {
            const float wx = plane ? st.wx1 : st.wx0, wy = plane ? st.wy1 : st.wy0;
            const float ux = 1.0f - wx, uy = 1.0f - wy;
            const int mt[4] = {m00[plane], m00[plane] + 1, m00[plane] + wW[plane], m00[plane] + wW[plane] + 1};
            const uint32_t p01 = pack2<kBF16>(ux * uy, wx * uy), p23 = pack2<kBF16>(ux * wy, wx * wy);
            const uint16_t hw[4] = {(uint16_t)(p01 & 0xFFFFu), (uint16_t)(p01 >> 16), (uint16_t)(p23 & 0xFFFFu), (uint16_t)(p23 >> 16)};
            uint8_t *Sp = Dbuf + plane * (16 * kChunk) + t * 16;
#pragma unroll
            for (int gq = 0; gq < 16; ++gq) *reinterpret_cast<uint4 *>(Sp + gq * kChunk) = make_uint4(0u, 0u, 0u, 0u);
            if (inwin[plane]) {   // same-thread stores: the four weights land after the zero fill of this column
#pragma unroll
              for (int tap = 0; tap < 4; ++tap)
                *reinterpret_cast<volatile uint16_t *>(Sp + (mt[tap] >> 3) * kChunk + (mt[tap] & 7) * 2) = hw[tap];
            }
          }
          fence_async_smem();
          tc_fence_before();
          bar_wg(0);
          HAV_MMA_ROUND({
            for (int plane = 0; plane < 2; ++plane)
              for (int k = 0; k < 8; ++k)
                umma_ss(tmC2 + plane * 64, smem_desc_mn(smem_base + kSmD + plane * (16 * kChunk) + k * 256, 128, kChunk, Q.swap_mn),
                        smem_desc_mn(A_addr + plane * 8 * kChunkA + k * 256, 128, kChunkA, Q.swap_mn), kIdescSc, k > 0);
          });
          // window texel m = this thread's TMEM lane: add its 64 channels to the global image unless the row is empty
#pragma unroll 1
          for (int plane = 0; plane < 2; ++plane) {
            const size_t texel = (size_t)(wymin[plane] + t / wW[plane]) * Wp + (wxmin[plane] + t % wW[plane]);
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {
              uint32_t r[32];
              HAV_TMEM_LD32(r, tmC2 + tm_lane + plane * 64 + q * 32);
              tmem_wait_ld();
              uint32_t any = 0u;
#pragma unroll
              for (int i = 0; i < 32; ++i) any |= r[i];
              if ((any & 0x7FFFFFFFu) != 0u) {
                float *gp = Q.gplanes_cl + texel * kPlaneC + q * 32;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  red_add_v4(gp + i * 4, __uint_as_float(r[i * 4]) * ginv, __uint_as_float(r[i * 4 + 1]) * ginv,
                             __uint_as_float(r[i * 4 + 2]) * ginv, __uint_as_float(r[i * 4 + 3]) * ginv);
              }
            }
          }
        }
        tc_fence_before();   // the next tile's L0 overwrites C0
      }
    }
  }
#undef HAV_MMA_ROUND
  tc_fence_before();
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
}


// ------------------------------------------------------------------------------------------------
// weight gradients: split-K tcgen05 GEMM over the per-tile operand images
// ------------------------------------------------------------------------------------------------
constexpr int kWgX = 0;                         // x      : 24 chunks (22 from HBM + constant one / zero chunk)
constexpr int kWgH1 = kWgX + 24 * kChunk;       // h1     : 18 chunks (16 + one / zero)
constexpr int kWgH2 = kWgH1 + 18 * kChunk;      // h2     : 18 chunks
constexpr int kWgD0 = kWgH2 + 18 * kChunk;      // d_pre0 : 16 chunks
constexpr int kWgD1 = kWgD0 + 16 * kChunk;      // d_pre1 : 16 chunks
constexpr int kWgDO = kWgD1 + 16 * kChunk;      // d_out  : 16 chunks (10 from HBM + 6 zero)
constexpr int kWgBar = kWgDO + 16 * kChunk;     // 221184
constexpr int kWgSmem = kWgBar + 64;
static_assert(kWgSmem <= 232448, "shared memory budget");

__device__ __forceinline__ void mbar_expect_tx_b(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_b(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// dW[n][k] += sum_samples d_pre[sample][n] * act[sample][k]: A = d_pre (M = 128 output features), B = act (N input features),
// K = the 128 samples of a tile; both operands are the tile images as they lie (MN-major, K group = 8 samples = 128 bytes,
// MN group = one 2 KB chunk).  One thread drives copies and MMAs; the accumulators stay in TMEM across all tiles of the CTA.
template <bool kBF16>
__global__ void __launch_bounds__(kThr, 1) wgrad_kernel(const uint8_t *__restrict__ dump, int ntiles, float *__restrict__ dW, int swap_mn) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int t = threadIdx.x, warp = t >> 5;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_ld = smem_base + kWgBar, bar_mma = bar_ld + 8;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kWgBar + 32);
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kWgBar + 32), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 32) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {   // constant chunks: the one column (bias gradient) and the zero padding
    const uint16_t one = kBF16 ? 0x3F80 : 0x3C00;
    const uint4 o4 = make_uint4(one, 0u, 0u, 0u), z4 = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(smem + kWgX + 22 * kChunk + t * 16) = o4;
    *reinterpret_cast<uint4 *>(smem + kWgX + 23 * kChunk + t * 16) = z4;
    *reinterpret_cast<uint4 *>(smem + kWgH1 + 16 * kChunk + t * 16) = o4;
    *reinterpret_cast<uint4 *>(smem + kWgH1 + 17 * kChunk + t * 16) = z4;
    *reinterpret_cast<uint4 *>(smem + kWgH2 + 16 * kChunk + t * 16) = o4;
    *reinterpret_cast<uint4 *>(smem + kWgH2 + 17 * kChunk + t * 16) = z4;
#pragma unroll
    for (int c = kOCh; c < 16; ++c) *reinterpret_cast<uint4 *>(smem + kWgDO + c * kChunk + t * 16) = z4;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (t == 0) {
    constexpr uint32_t kMN = kAMajorMN | kBMajorMN;
    constexpr uint32_t id0 = instr_desc(192, kBF16) | kMN, id1 = instr_desc(144, kBF16) | kMN;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint8_t *src = dump + (size_t)tile * kTileBytes;
      mbar_expect_tx_b(bar_ld, kTileBytes);
      bulk_g2s_b(smem_base + kWgX, src + kOffX, kXCh * kChunk, bar_ld);
      bulk_g2s_b(smem_base + kWgH1, src + kOffH1, kHCh * kChunk, bar_ld);
      bulk_g2s_b(smem_base + kWgH2, src + kOffH2, kHCh * kChunk, bar_ld);
      bulk_g2s_b(smem_base + kWgDO, src + kOffDO, kOCh * kChunk, bar_ld);
      bulk_g2s_b(smem_base + kWgD1, src + kOffD1, kHCh * kChunk, bar_ld);
      bulk_g2s_b(smem_base + kWgD0, src + kOffD0, kHCh * kChunk, bar_ld);
      mbar_wait(bar_ld, ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u, ko = k * 256;
        umma_ss(tmem_base + kDW0, smem_desc_mn(smem_base + kWgD0 + ko, 128, kChunk, swap_mn),
                smem_desc_mn(smem_base + kWgX + ko, 128, kChunk, swap_mn), id0, acc);
        umma_ss(tmem_base + kDW1, smem_desc_mn(smem_base + kWgD1 + ko, 128, kChunk, swap_mn),
                smem_desc_mn(smem_base + kWgH1 + ko, 128, kChunk, swap_mn), id1, acc);
        umma_ss(tmem_base + kDWH, smem_desc_mn(smem_base + kWgDO + ko, 128, kChunk, swap_mn),
                smem_desc_mn(smem_base + kWgH2 + ko, 128, kChunk, swap_mn), id1, acc);
      }
      umma_commit(bar_mma);
      mbar_wait(bar_mma, ph);   // the operands may be overwritten by the next tile's copies
      ph ^= 1;
    }
  }
  __syncthreads();
  tc_fence_after();
  // accumulators -> global (row n = TMEM lane)
  const uint32_t tm_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  float *dst = dW + (size_t)t * kDWCols;
#pragma unroll 1
  for (int q = 0; q < kDWCols / 32; ++q) {
    uint32_t r[32];
    HAV_TMEM_LD32(r, tm_row + q * 32);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      red_add_v4(dst + q * 32 + i * 4, __uint_as_float(r[i * 4]), __uint_as_float(r[i * 4 + 1]), __uint_as_float(r[i * 4 + 2]),
                 __uint_as_float(r[i * 4 + 3]));
  }
  tc_fence_before();
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
}

// ------------------------------------------------------------------------------------------------
// small kernels: loss scale, layout finalisation
// ------------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const float *__restrict__ g, size_t n, unsigned int *__restrict__ out) {
  float m = 0.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = fabsf(g[i]);
    if (v < 3.0e38f) m = fmaxf(m, v);   // ignore inf / nan when choosing the scale
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}
// scale = the power of two that brings max|g| to [512, 1024): 16-bit gradient operands then sit well inside the fp16 range
__global__ void scale_kernel(const unsigned int *__restrict__ amax, float *__restrict__ scale, float fixed) {
  float s = fixed;
  if (!(fixed > 0.0f)) {
    const float m = __uint_as_float(*amax);
    s = 1.0f;
    if (m > 0.0f) {
      int e;
      frexpf(m, &e);          // m = f * 2^e, f in [0.5, 1)
      s = ldexpf(1.0f, min(max(10 - e, -60), 60));
    }
  }
  scale[0] = s, scale[1] = 1.0f / s;
}

// channels-last fp32 gradient image [2B][H+3][W+3][64] -> planes gradient [2B][64][H][W] (NCHW); one block per (image, row)
__global__ void __launch_bounds__(256) planes_grad_finalize_kernel(const float *__restrict__ gcl, float *__restrict__ out, int H, int W) {
  extern __shared__ float tile[];   // [W][65]
  const int img = blockIdx.y, y = blockIdx.x;
  const int Hp = H + kPadLo + kPadHi, Wp = W + kPadLo + kPadHi;
  const float *src = gcl + (((size_t)img * Hp + (y + kPadLo)) * Wp + kPadLo) * kPlaneC;
  for (int i = threadIdx.x; i < kPlaneC * W; i += blockDim.x) tile[(i / kPlaneC) * (kPlaneC + 1) + (i % kPlaneC)] = src[i];
  __syncthreads();
  float *dst = out + (size_t)img * kPlaneC * H * W + (size_t)y * W;
  for (int i = threadIdx.x; i < kPlaneC * W; i += blockDim.x) {
    const int c = i / W, x = i % W;
    dst[(size_t)c * H * W + x] = tile[x * (kPlaneC + 1) + c];
  }
}

// accumulator image [128][480] -> the reference's parameter gradients.  Internal K order of L0 is plane-0 channels | plane-1
// channels | PE | bias (render_tc.cu pack_mlp_16_kernel); head rows are fc_rgbFeat (0..63), fc_alpha (64) and the composed
// fc_rgb o fc_rgbFeat rows (65..67):  logit = Wr (Wf h + bf) + br, so with D = d(loss)/d(composed rows):
//   dWf += Wr^T D[:, :128],  dbf += Wr^T D[:, 128],  dWr = D[:, :128] Wf^T + D[:, 128] bf^T,  dbr = D[:, 128].
struct MlpGrads {
  float *w0, *b0, *w1, *b1, *wa, *ba, *wf, *bf, *wr, *br;
};
__global__ void mlp_grad_finalize_kernel(const float *__restrict__ dW, const float *__restrict__ scale, const float *__restrict__ wf,
                                         const float *__restrict__ bf, const float *__restrict__ wr, MlpGrads G) {
  const float inv = scale[1];
  const int nth = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = tid; i < kHid * kIn; i += nth) {
    const int n = i / kIn, j = i % kIn;
    const int k = j < kFeat ? (j & 1) * kPlaneC + (j >> 1) : j;
    G.w0[i] = dW[n * kDWCols + kDW0 + k] * inv;
  }
  for (int i = tid; i < kHid * kHid; i += nth) G.w1[i] = dW[(i / kHid) * kDWCols + kDW1 + (i % kHid)] * inv;
  for (int n = tid; n < kHid; n += nth) {
    G.b0[n] = dW[n * kDWCols + kDW0 + kIn] * inv;
    G.b1[n] = dW[n * kDWCols + kDW1 + kHid] * inv;
    G.wa[n] = dW[kRgbFeat * kDWCols + kDWH + n] * inv;
  }
  if (tid == 0) G.ba[0] = dW[kRgbFeat * kDWCols + kDWH + kHid] * inv;
  for (int i = tid; i < kRgbFeat * (kHid + 1); i += nth) {   // fc_rgbFeat weight (k < 128) and bias (k == 128)
    const int c = i / (kHid + 1), k = i % (kHid + 1);
    float v = dW[c * kDWCols + kDWH + k];
#pragma unroll
    for (int j = 0; j < 3; ++j) v = fmaf(wr[j * kRgbFeat + c], dW[(kRgbFeat + 1 + j) * kDWCols + kDWH + k], v);
    if (k < kHid) G.wf[c * kHid + k] = v * inv;
    else G.bf[c] = v * inv;
  }
  for (int i = tid; i < 3 * kRgbFeat; i += nth) {
    const int j = i / kRgbFeat, c = i % kRgbFeat;
    const float *D = dW + (kRgbFeat + 1 + j) * kDWCols + kDWH;
    float v = D[kHid] * bf[c];
    for (int k = 0; k < kHid; ++k) v = fmaf(D[k], wf[c * kHid + k], v);
    G.wr[i] = v * inv;
  }
  if (tid < 3) G.br[tid] = dW[(kRgbFeat + 1 + tid) * kDWCols + kDWH + kHid] * inv;
}

}  // namespace bwd

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline uint64_t align_up_b(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct BwdLayout {
  uint64_t scalars, dW, wimg, planes_cl, gplanes_cl, dump, total, gplanes_bytes;
  int num_blocks, tiles;
};

static BwdLayout bwd_layout(const hav_render_args *a) {
  BwdLayout L;
  memset(&L, 0, sizeof(L));
  const int64_t total = (int64_t)a->batch * a->rays;
  L.num_blocks = (int)((total + kRaysPerBlock - 1) / kRaysPerBlock);
  const int Sf = a->num_fine > 0 ? (a->num_coarse + 1) / 2 + a->num_fine : 0;
  L.tiles = L.num_blocks * (a->num_coarse + Sf);
  uint64_t off = 0;
  L.scalars = off, off = align_up_b(off + 256, 256);
  L.dW = off, off = align_up_b(off + (uint64_t)kHid * bwd::kDWCols * 4, 256);
  L.wimg = off, off = align_up_b(off + tc_weight_image_bytes(), 256);
  L.planes_cl = off, off = align_up_b(off + tc_planes_bytes(2 * a->batch, a->plane_h, a->plane_w), 256);
  L.gplanes_bytes = tc_planes_bytes(2 * a->batch, a->plane_h, a->plane_w) * 2;   // fp32 instead of 16 bit
  L.gplanes_cl = off, off = align_up_b(off + L.gplanes_bytes, 256);
  L.dump = off, off = align_up_b(off + (uint64_t)L.tiles * bwd::kTileBytes, 256);
  L.total = off;
  return L;
}

static int bwd_check(const hav_render_bwd_args *b) {
  if (b == nullptr || b->fwd == nullptr) return HAV_E_NULL;
  if (b->struct_bytes != sizeof(hav_render_bwd_args)) return HAV_E_VALUE;
  const hav_render_args *a = b->fwd;
  int rc = render_check_args(a);
  if (rc != HAV_OK) return rc;
  if (a->precision != HAV_PREC_BF16 && a->precision != HAV_PREC_FP16) return HAV_E_VALUE;
  if ((int64_t)a->batch * a->rays == 0) return HAV_OK;
  if (a->num_fine > 0 && a->z_fine == nullptr) return HAV_E_NULL;
  const void *req[] = {b->g_planes, b->g_wvol, b->g_w0, b->g_b0, b->g_w1, b->g_b1, b->g_w_alpha, b->g_b_alpha,
                       b->g_w_feat, b->g_b_feat, b->g_w_rgb, b->g_b_rgb};
  for (const void *p : req)
    if (p == nullptr) return HAV_E_NULL;
  return HAV_OK;
}

}  // namespace hav

using namespace hav;

extern "C" uint64_t hav_render_backward_workspace_bytes(const hav_render_bwd_args *b) {
  if (bwd_check(b) != HAV_OK) return 0;
  return bwd_layout(b->fwd).total;
}

extern "C" int hav_render_backward(const hav_render_bwd_args *b, void *stream) {
  int rc = bwd_check(b);
  if (rc != HAV_OK) return rc;
  const hav_render_args *a = b->fwd;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)a->batch * a->rays;
  const size_t plane_elems = (size_t)2 * a->batch * kPlaneC * a->plane_h * a->plane_w;
  const size_t vol_elems = (size_t)2 * a->vol_d * a->vol_h * a->vol_w;
  cudaError_t e;
#define HAV_TRY(x) if ((e = (x)) != cudaSuccess) return (int)e
  if (total == 0) {   // empty ray batch: all gradients are zero
    HAV_TRY(cudaMemsetAsync(b->g_planes, 0, plane_elems * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_wvol, 0, vol_elems * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_w0, 0, kHid * kIn * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_b0, 0, kHid * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_w1, 0, kHid * kHid * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_b1, 0, kHid * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_w_alpha, 0, kHid * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_b_alpha, 0, 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_w_feat, 0, kRgbFeat * kHid * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_b_feat, 0, kRgbFeat * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_w_rgb, 0, 3 * kRgbFeat * 4, st));
    HAV_TRY(cudaMemsetAsync(b->g_b_rgb, 0, 3 * 4, st));
    return HAV_OK;
  }
  int dev = 0, major = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (major != 10) return HAV_E_ARCH;
  BwdLayout L = bwd_layout(a);
  if (b->workspace == nullptr) return HAV_E_NULL;
  if (b->workspace_bytes < L.total || ((uintptr_t)b->workspace & 255) != 0) return HAV_E_WORKSPACE;
  uint8_t *ws = (uint8_t *)b->workspace;
  const bool bf16 = a->precision == HAV_PREC_BF16;
  static const int swap_mn = getenv("HAV_BWD_SWAP_MN") != nullptr ? 1 : 0;

  RenderDev P;
  render_fill_dev(a, P);
  P.wimg = ws + L.wimg;
  P.planes_cl = (const uint16_t *)(ws + L.planes_cl);
  launch_pack_mlp_16(a, ws + L.wimg, st);
  HAV_TRY(launch_pack_planes_16(a->planes, (uint16_t *)(ws + L.planes_cl), 2 * a->batch, a->plane_h, a->plane_w, bf16, st));

  // loss scale from max |upstream gradient|
  unsigned int *amax = (unsigned int *)(ws + L.scalars);
  float *scale = (float *)(ws + L.scalars) + 4;
  HAV_TRY(cudaMemsetAsync(ws + L.scalars, 0, 256, st));
  HAV_TRY(cudaMemsetAsync(ws + L.dW, 0, (size_t)kHid * bwd::kDWCols * 4, st));
  HAV_TRY(cudaMemsetAsync(ws + L.gplanes_cl, 0, L.gplanes_bytes, st));
  HAV_TRY(cudaMemsetAsync(b->g_wvol, 0, vol_elems * 4, st));
  if (!(b->grad_scale > 0.0f)) {
    const struct { const float *p; size_t n; } gs[] = {
        {b->g_rgb_coarse, (size_t)total * kOut}, {b->g_depth_coarse, (size_t)total}, {b->g_acc_coarse, (size_t)total},
        {b->g_rgb_fine, (size_t)total * kOut},   {b->g_depth_fine, (size_t)total},   {b->g_acc_fine, (size_t)total}};
    for (const auto &g : gs)
      if (g.p != nullptr) {
        const int blocks = (int)((g.n + 1023) / 1024 < 592 ? (g.n + 1023) / 1024 : 592);
        bwd::absmax_kernel<<<blocks, 256, 0, st>>>(g.p, g.n, amax);
      }
  }
  bwd::scale_kernel<<<1, 1, 0, st>>>(amax, scale, b->grad_scale);

  bwd::BwdDev Q;
  memset(&Q, 0, sizeof(Q));
  Q.g_rgb[0] = b->g_rgb_coarse, Q.g_depth[0] = b->g_depth_coarse, Q.g_acc[0] = b->g_acc_coarse;
  Q.g_rgb[1] = b->g_rgb_fine, Q.g_depth[1] = b->g_depth_fine, Q.g_acc[1] = b->g_acc_fine;
  Q.scale = scale;
  Q.gplanes_cl = (float *)(ws + L.gplanes_cl);
  Q.gwvol = b->g_wvol;
  Q.dump = ws + L.dump;
  Q.swap_mn = swap_mn;
  const int grid = L.num_blocks < sms ? L.num_blocks : sms;
  if (bf16) {
    HAV_TRY(cudaFuncSetAttribute(bwd::render_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmBytesB));
    bwd::render_bwd_kernel<true><<<grid, bwd::kThr, bwd::kSmBytesB, st>>>(P, Q, L.num_blocks);
  } else {
    HAV_TRY(cudaFuncSetAttribute(bwd::render_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmBytesB));
    bwd::render_bwd_kernel<false><<<grid, bwd::kThr, bwd::kSmBytesB, st>>>(P, Q, L.num_blocks);
  }
  HAV_TRY(cudaGetLastError());
  const int wgrid = L.tiles < sms ? L.tiles : sms;
  if (bf16) {
    HAV_TRY(cudaFuncSetAttribute(bwd::wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kWgSmem));
    bwd::wgrad_kernel<true><<<wgrid, bwd::kThr, bwd::kWgSmem, st>>>(ws + L.dump, L.tiles, (float *)(ws + L.dW), swap_mn);
  } else {
    HAV_TRY(cudaFuncSetAttribute(bwd::wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kWgSmem));
    bwd::wgrad_kernel<false><<<wgrid, bwd::kThr, bwd::kWgSmem, st>>>(ws + L.dump, L.tiles, (float *)(ws + L.dW), swap_mn);
  }
  HAV_TRY(cudaGetLastError());
  bwd::MlpGrads G{b->g_w0, b->g_b0, b->g_w1, b->g_b1, b->g_w_alpha, b->g_b_alpha, b->g_w_feat, b->g_b_feat, b->g_w_rgb, b->g_b_rgb};
  bwd::mlp_grad_finalize_kernel<<<64, 256, 0, st>>>((const float *)(ws + L.dW), scale, a->w_feat, a->b_feat, a->w_rgb, G);
  const size_t fsm = (size_t)a->plane_w * (kPlaneC + 1) * sizeof(float);
  HAV_TRY(cudaFuncSetAttribute(bwd::planes_grad_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
  bwd::planes_grad_finalize_kernel<<<dim3(a->plane_h, 2 * a->batch), 256, fsm, st>>>((const float *)(ws + L.gplanes_cl), b->g_planes,
                                                                                      a->plane_h, a->plane_w);
  HAV_TRY(cudaGetLastError());
#undef HAV_TRY
  return HAV_OK;
}
