// Shared pieces of the tensor-core render kernels: operand layouts, tcgen05 / TMEM / mbarrier PTX wrappers,
// the hidden-layer epilogue.  See render_tc.cu (v1: two self-contained warpgroups) and render_tc2.cu
// (v2: producer / consumer warp specialisation) for the kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "render_common.cuh"
#include "render_internal.h"

namespace hav {
namespace tc {

constexpr int kWGs = 2;
constexpr int kThreads = kWGs * 128;
constexpr int kK0 = 192;                      // L0: 176 inputs + bias column (k = 176) + 15 zero columns
constexpr int kK1 = 144;                      // L1 / head: 128 inputs + bias column (k = 128) + 15 zero columns
constexpr int kNH = 80;                       // head rows: 0..63 fc_rgbFeat, 64 fc_alpha, 65..67 fc_rgb o fc_rgbFeat, pad
constexpr int kChunkB = 128 * 16;             // bytes of one 8-wide K chunk of a 128-row weight matrix
constexpr int kChunkBH = kNH * 16;
constexpr int kW0Off = 0;
constexpr int kW1Off = kW0Off + (kK0 / 8) * kChunkB;
constexpr int kWHOff = kW1Off + (kK1 / 8) * kChunkB;
constexpr int kWImgBytes = kWHOff + (kK1 / 8) * kChunkBH;   // 109056
constexpr int kChunkA = 128 * 16 + 16;        // A-operand chunk stride, +16 B so the gather's stores spread over banks
constexpr int kAChunks = 24;
constexpr int kABytes = kAChunks * kChunkA;   // 49536
constexpr int kOnesChunk = 22;                // chunk 22 = [1,0,..,0] per row (bias column), chunk 23 = zeros
constexpr int kStageBytes = 128 * 32;         // per-row tap descriptors handed from the row threads to the gather
constexpr int kSmemA = kWImgBytes;
constexpr int kSmemStage = kSmemA + kWGs * kABytes;
constexpr int kSmemBar = kSmemStage + kWGs * kStageBytes;
constexpr int kSmemBytes = kSmemBar + 64;
constexpr int kTmemCols = 512;
constexpr int kPadLo = 1, kPadHi = 2;         // zero border of the packed planes: "zeros" padding for free

struct Stage {   // 32 bytes
  int off0, off1;             // texel index of tap (y0,x0) in the packed plane array, plane 0 / plane 1
  float wx0, wy0, wx1, wy1;   // fractional weights of the +1 taps: plane 0 (x,y), plane 1 (x,y)
  int pad0, pad1;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_wg(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one lane of a converged warp (call from warp-uniform code: keeps the MMA operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep in hardware, do not spin
}
// plain polling variant (no suspend-time hint)
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, one K = 16 step
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: A rows = TMEM lanes, 16-bit elements packed two per 32-bit column
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// no-swizzle K-major shared-memory matrix descriptor: core matrix = 8 rows x 16 bytes, contiguous 128 B;
// SBO = distance between 8-row groups, LBO = distance between the two 8-wide K chunks of one K = 16 step.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(int n, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

#define HAV_TMEM_LD32(r, taddr)                                                                                       \
  asm volatile(                                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                      \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),        \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                     \
      : "r"(taddr))
#define HAV_TMEM_ST16(taddr, r)                                                                                   \
  asm volatile(                                                                                                   \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"    \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), \
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")
#define HAV_TMEM_ST8(taddr, r)                                                                    \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"           \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory")
#define HAV_TMEM_LD4(r, taddr)                                                 \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])                \
               : "r"(taddr))

template <bool kBF16>
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
  uint32_t d;
  if (kBF16) asm("cvt.rn.relu.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (kBF16) asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// d = a * w + d on packed 16-bit pairs (w = the same weight in both halves)
template <bool kBF16>
__device__ __forceinline__ uint32_t fma2(uint32_t a, uint32_t w, uint32_t c) {
  uint32_t d;
  if (kBF16) asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(w), "r"(c));
  else asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(w), "r"(c));
  return d;
}
template <bool kBF16>
__device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t w) {
  uint32_t d;
  if (kBF16) asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(w));
  else asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(w));
  return d;
}


// relu + 16-bit pack of one 128-column fp32 accumulator row (this thread's TMEM lane) into the A operand of the
// next layer.  kTS: in place over the accumulator's columns [0,64) as packed pairs, plus the constant bias
// column pair (1,0) at column 64 (columns 65..71 zero) -- reads of columns [32q, 32q+32) always precede the
// write of [16q, 16q+16).  !kTS: shared-memory chunks 0..15 (the constant chunk 22/23 carries the bias column).
// kCheck (fp16 only): *sat collects whether any packed value sits at the saturation value 65504 (0x7BFF).
template <bool kBF16, bool kTS, bool kCheck = false>
__device__ __forceinline__ void hidden_epilogue(uint32_t tm_row, uint8_t *Abuf, int t, uint32_t *sat = nullptr) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[32];
    HAV_TMEM_LD32(r, tm_row + q * 32);
    tmem_wait_ld();
    uint32_t v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = pack_relu<kBF16>(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1]));
    if (kCheck) {
#pragma unroll
      for (int c = 0; c < 16; ++c) *sat |= __vcmpeq2(v[c], 0x7BFF7BFFu);
    }
    if (kTS) {
      HAV_TMEM_ST16(tm_row + q * 16, v);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4 *>(Abuf + (q * 4 + c) * kChunkA + t * 16) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
  }
  if (kTS) {
    uint32_t one[8] = {kBF16 ? 0x3F80u : 0x3C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    HAV_TMEM_ST8(tm_row + 64, one);
    tmem_wait_st();
    tc_fence_before();
  } else {
    tc_fence_before();
    fence_async_smem();
  }
}

// the 8 + 1 K-steps of a hidden-layer GEMM: A = relu(previous accumulator) (TMEM or smem), B = weight matrix at w_addr
template <bool kTS>
__device__ __forceinline__ void issue_hidden(uint32_t tm_d, uint32_t tm_a, uint32_t A_addr, uint32_t w_addr, int chunk_b,
                                             uint32_t idesc) {
#pragma unroll
  for (int k = 0; k <= kHid / 16; ++k) {
    const uint64_t bdesc = smem_desc(w_addr + 2 * k * chunk_b, chunk_b, 128);
    if (kTS) umma_ts(tm_d, tm_a + k * 8, bdesc, idesc, k > 0);
    else umma_ss(tm_d, smem_desc(A_addr + (k < kHid / 16 ? 2 * k : kOnesChunk) * kChunkA, kChunkA, 128), bdesc, idesc, k > 0);
  }
}

}  // namespace tc
}  // namespace hav
