// Tensor-core fused render kernel, v3: CTA pairs (tcgen05 cta_group::2) + a split-precision mode.
//
// Launch: clusters of two CTAs (one per SM of a TPC), persistent, 512 threads each.  The warp-specialised structure of v2
// (render_tc2.cu: producer groups fill the K-major A operand X of one M = 128 tile per sample step, consumer warpgroups run
// L0 / L1 / head GEMMs, epilogues and the composite) is kept; what changes is the MMA:
//
//   * every GEMM is ONE M = 256 instruction stream issued by the leader CTA (rank 0) of the cluster for both CTAs' row tiles:
//     each CTA contributes its own 128-row A tile (X in its shared memory, or the hidden activations in its TMEM) and HALF of
//     the weight rows (B operand: N/2 rows per CTA), so the B-operand shared-memory reads per row tile and the per-CTA weight
//     footprint are halved (109 KB -> 55.5 KB);
//   * cross-CTA hand-shakes: each CTA synchronises its own 128 threads locally (named barrier / local mbarrier), then ONE
//     thread of the non-leader CTA relays the event to an mbarrier in the leader's shared memory (mapa / shared::cluster remote
//     arrive: one message per event instead of 128); completions come back through tcgen05.commit ... multicast::cluster.
//
// The halved weight footprint is what makes the split-precision mode fit (HAV_PREC_FP16X3): A = A_hi + A_lo and
// W = W_hi + W_lo as fp16 pairs, D = A_hi W_hi + A_lo W_hi + A_hi W_lo accumulated in fp32 TMEM (the dropped A_lo W_lo term
// is 2^-22 relative), fp32 planes / blends / positional encoding / composite: fp32-class results on the tensor cores.
// One pair per CTA in that mode (weights hi + lo 111 KB, X hi + lo 89 KB).
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace hav {
namespace tc3 {

using namespace tc;

constexpr int kThreads3 = 512;
constexpr int kConsumerRegs = 168, kProducerRegs = 88;
constexpr int kNH3 = 96;                                  // head rows padded to a multiple of 32 (2-SM A-from-TMEM MMA)
constexpr int kHalfChunkB = 64 * 16, kHalfChunkBH = (kNH3 / 2) * 16;
constexpr int kW0h = 0;
constexpr int kW1h = kW0h + (kK0 / 8) * kHalfChunkB;
constexpr int kWHh = kW1h + (kK1 / 8) * kHalfChunkB;
constexpr int kWHalfBytes = kWHh + (kK1 / 8) * kHalfChunkBH;   // 56832
constexpr int kXChunks = 22;
constexpr int kXBytes = kXChunks * kChunkA;               // 45408
constexpr int kConstBytes = 2 * kChunkA;
constexpr int kStageRow = 48;
constexpr int kStageBytes3 = 128 * kStageRow;
// barrier slots per pair (8 bytes each)
constexpr int kBarXFull = 0, kBarXFree = 1, kBarMma = 2, kBarZFine = 3, kBarReady = 4, kBarPeerFull = 5, kBarsPerPair = 6;
constexpr int kHCols = 136;                               // split mode: hidden activations hi [0,72) | lo [72,136) (TMEM columns)

enum { kModeF16 = 0, kModeBF16 = 1, kModeSplit = 2 };

template <int kMode>
struct Cfg {
  static constexpr bool bf16 = kMode == kModeBF16;
  static constexpr bool split = kMode == kModeSplit;
  static constexpr int pairs = split ? 1 : 2;
  static constexpr int sets = split ? 2 : 1;              // hi (+ lo) copies of the weights and of X
  static constexpr int smW = 0;
  static constexpr int smX = smW + sets * kWHalfBytes;
  static constexpr int smConst = smX + pairs * sets * kXBytes;
  static constexpr int smStage = smConst + kConstBytes;
  static constexpr int smBar = smStage + pairs * 2 * kStageBytes3;
  static constexpr int smBytes = smBar + 128;
};
static_assert(Cfg<kModeF16>::smBytes <= 232448 && Cfg<kModeSplit>::smBytes <= 232448, "shared memory budget");

// ---- cluster / 2-CTA PTX ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait on a barrier of THIS CTA whose arrivals may come from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bar_named(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// completion of every MMA issued so far by this thread -> the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_ts2(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// M = 256 (both CTAs' 128-row tiles), N = n, fp16 / bf16 operands, fp32 accumulate
__host__ __device__ constexpr uint32_t instr_desc2(int n, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}

// v = hi + lo with hi = fp16(v), lo = fp16(v - hi): two packed pairs
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// tap base + bilinear weights; 16-bit modes: weights pre-packed as duplicated 16-bit pairs (HFMA2 blend), split mode: fp32
template <bool kBF16>
__device__ __forceinline__ void plane_taps_packed(float gx, float gy, int H, int W, int img, int &off, uint32_t (&w)[4]) {
  float ix = fminf(fmaxf(unnorm(gx, W), -1.0f), (float)W);
  float iy = fminf(fmaxf(unnorm(gy, H), -1.0f), (float)H);
  float x0f = floorf(ix), y0f = floorf(iy);
  float wx = ix - x0f, wy = iy - y0f, ux = 1.0f - wx, uy = 1.0f - wy;
  const int Hp = H + kPadLo + kPadHi, Wp = W + kPadLo + kPadHi;
  off = (img * Hp + ((int)y0f + kPadLo)) * Wp + ((int)x0f + kPadLo);
  float a = ux * uy, b = wx * uy, c = ux * wy, d = wx * wy;
  w[0] = pack2<kBF16>(a, a), w[1] = pack2<kBF16>(b, b), w[2] = pack2<kBF16>(c, c), w[3] = pack2<kBF16>(d, d);
}
// fp32 taps in ATen's own arithmetic (grid_sampler_2d: nw = (ix_se - ix)(iy_se - iy), ...), zero border handles 'zeros' padding
__device__ __forceinline__ void plane_taps_f32(float gx, float gy, int H, int W, int img, int &off, float (&w)[4]) {
  float ix = fminf(fmaxf(unnorm(gx, W), -1.0f), (float)W);
  float iy = fminf(fmaxf(unnorm(gy, H), -1.0f), (float)H);
  float x0f = floorf(ix), y0f = floorf(iy);
  float wx1 = ix - x0f, wy1 = iy - y0f;
  float wx0 = (x0f + 1.0f) - ix, wy0 = (y0f + 1.0f) - iy;
  const int Hp = H + kPadLo + kPadHi, Wp = W + kPadLo + kPadHi;
  off = (img * Hp + ((int)y0f + kPadLo)) * Wp + ((int)x0f + kPadLo);
  w[0] = wx0 * wy0, w[1] = wx1 * wy0, w[2] = wx0 * wy1, w[3] = wx1 * wy1;
}

// ray block of (iteration, cluster, CTA rank, pair); >= num_ray_blocks = a dummy block (rows masked, barriers still served)
template <int kPairs>
__device__ __forceinline__ int block_of(int it, int pair) {
  return ((it * (int)num_clusters_x() + (int)cluster_id_x()) * 2 + (int)cluster_ctarank()) * kPairs + pair;
}

// ------------------------------------------------------------------------------------------------
// producer
// ------------------------------------------------------------------------------------------------
template <int kMode>
__device__ __forceinline__ void producer_loop(const RenderDev &P, int num_ray_blocks, int iters, uint8_t *smem, uint32_t smem_base,
                                              int pair, int t) {
  using C = Cfg<kMode>;
  constexpr bool kBF16 = C::bf16, kSplit = C::split;
  // split mode: EIGHT producer warps serve the single pair -- two threads per row (tp = 0..255: row t = tp & 127, half = tp >> 7;
  // half 0 builds the tap descriptors and the positional encoding of frequencies 0..3, half 1 frequencies 4..7) and each warp
  // gathers 16 rows instead of 32: the fp32 gather + 48 exact sines per row made the 4-warp producer the bottleneck
  const int tp = t, half = kSplit ? (tp >> 7) : 0;
  t = tp & 127;
  const int warp = tp >> 5, lane = tp & 31;
  constexpr int kRowsPerWarp = kSplit ? 16 : 32;
  uint8_t *X = smem + C::smX + pair * C::sets * kXBytes;          // split: X_hi, then X_lo at + kXBytes
  uint8_t *stage_base = smem + C::smStage + pair * 2 * kStageBytes3;
  const uint32_t bars = smem_base + C::smBar + pair * kBarsPerPair * 8;
  const uint32_t bar_full = bars + kBarXFull * 8;
  const uint32_t bar_free = bars + kBarXFree * 8, bar_zfine = bars + kBarZFine * 8;
  const int Wp = P.PW + kPadLo + kPadHi;
  uint32_t n = 0, zfine_phase = 0;

  for (int it = 0; it < iters; ++it) {
    const int rb = block_of<C::pairs>(it, pair);
    const int g = rb < num_ray_blocks ? rb * kRaysPerBlock + t : P.total_rays;   // dummy block: every row invalid
    const Ray ray = load_ray(P, g);
    const int gi = ray.valid ? g : 0;
    float Tm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Tm[i] = __ldg(P.invT + (size_t)ray.b * 12 + i);
    const int slot = blockIdx.x * C::pairs + pair;
    const float *zcol = P.zbuf + (size_t)slot * P.Sf * kRaysPerBlock + t;
    const int npass = P.nfine > 0 ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
      const int S = pass == 0 ? P.Sc : P.Sf;
      if (pass == 1) {
        mbar_wait(bar_zfine, zfine_phase);
        zfine_phase ^= 1;
      }
#pragma unroll 1
      for (int s = 0; s < S; ++s, ++n) {
        uint8_t *stage = stage_base + (n & 1) * kStageBytes3;
        const float z = pass == 0 ? coarse_z(P, ray, gi, s) : __ldcg(zcol + s * kRaysPerBlock);
        float p[3], pc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) p[j] = kSplit ? __fadd_rn(ray.o[j], __fmul_rn(ray.d[j], z)) : fmaf(ray.d[j], z, ray.o[j]);
        skin_warp(P, Tm, p, pc);
        {
          float qx = pc[0] * P.ps[0] + P.pt[0], qy = pc[1] * P.ps[1] + P.pt[1], qz = pc[2] * P.ps[2] + P.pt[2];
          int off0, off1;
          uint4 *sr = reinterpret_cast<uint4 *>(stage + t * kStageRow);
          if constexpr (kSplit) {
            if (half == 0) {
              float w0[4], w1[4];
              plane_taps_f32(qx, qy, P.PH, P.PW, ray.b, off0, w0);
              plane_taps_f32(qz, qy, P.PH, P.PW, P.B + ray.b, off1, w1);
              sr[0] = make_uint4(off0, off1, 0u, 0u);
              sr[1] = make_uint4(__float_as_uint(w0[0]), __float_as_uint(w0[1]), __float_as_uint(w0[2]), __float_as_uint(w0[3]));
              sr[2] = make_uint4(__float_as_uint(w1[0]), __float_as_uint(w1[1]), __float_as_uint(w1[2]), __float_as_uint(w1[3]));
            }
          } else {
            uint32_t w0[4], w1[4];
            plane_taps_packed<kBF16>(qx, qy, P.PH, P.PW, ray.b, off0, w0);
            plane_taps_packed<kBF16>(qz, qy, P.PH, P.PW, P.B + ray.b, off1, w1);
            sr[0] = make_uint4(off0, off1, 0u, 0u);
            sr[1] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
            sr[2] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
          }
        }
        // positional encoding, order [f][sin|cos][xyz] (model/network/embedder.py:32-61)
        uint32_t pk[kSplit ? 12 : 24], pl[kSplit ? 12 : 1];
        if constexpr (kSplit) {   // exact: sin(a), sin(a + pi/2) per frequency, like the reference; this thread's 4 frequencies
          const float f0 = half == 0 ? 1.0f : 16.0f;
#pragma unroll
          for (int f = 0; f < kFreqs / 2; ++f) {
            const float fr = f0 * (float)(1 << f);
            float sn[3], cn[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float ang = pc[j] * fr;
              sn[j] = sinf(ang), cn[j] = sinf(ang + 1.57079632679489661923f);
            }
            split2(sn[0], sn[1], pk[f * 3 + 0], pl[f * 3 + 0]);
            split2(sn[2], cn[0], pk[f * 3 + 1], pl[f * 3 + 1]);
            split2(cn[1], cn[2], pk[f * 3 + 2], pl[f * 3 + 2]);
          }
        } else {
          float sn[3], cn[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) __sincosf(pc[j], &sn[j], &cn[j]);
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
        }
        if constexpr (kSplit) asm volatile("bar.sync 3, 256;" ::: "memory");   // tap descriptors of all 128 rows are visible
        else bar_named(3 + pair);
        if (n > 0) mbar_wait(bar_free, (n - 1) & 1);          // L0 of the previous tile has finished reading X (both CTAs)
        if constexpr (kSplit) {      // 24 of the 48 encoding values = 3 chunks: half 0 -> chunks 16..18, half 1 -> 19..21
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            uint8_t *dst = X + (16 + 3 * half + c) * kChunkA + t * 16;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
            *reinterpret_cast<uint4 *>(dst + kXBytes) = make_uint4(pl[c * 4], pl[c * 4 + 1], pl[c * 4 + 2], pl[c * 4 + 3]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 6; ++c)
            *reinterpret_cast<uint4 *>(X + (16 + c) * kChunkA + t * 16) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        }
        if constexpr (!kSplit) {
          // cooperative gather, 16-bit planes: one step = 4 consecutive rows x 1 plane x 8 channel octets (see render_tc2.cu)
          const uint4 *planes = reinterpret_cast<const uint4 *>(P.planes_cl);
          const int oct = lane & 7, rsub = lane >> 3;
          const uint8_t *sp0 = stage + (warp * 32 + rsub) * kStageRow;
          uint8_t *xrow = X + oct * kChunkA + (warp * 32 + rsub) * 16;
          const size_t row_pitch = (size_t)Wp * 8;
          uint4 ta[4], tb[4], wa, wb;
          int offn;
          auto load_desc = [&](int i, uint4 &w) {
            const uint8_t *sp = sp0 + (i >> 1) * 4 * kStageRow;
            w = *reinterpret_cast<const uint4 *>(sp + 16 + (i & 1) * 16);
            return *reinterpret_cast<const int *>(sp + (i & 1) * 4);
          };
          auto issue = [&](int off, uint4 (&tt)[4]) {
            const uint4 *tp = planes + (size_t)off * 8 + oct;
            tt[0] = __ldg(tp), tt[1] = __ldg(tp + 8), tt[2] = __ldg(tp + row_pitch), tt[3] = __ldg(tp + row_pitch + 8);
          };
          auto blend = [&](const uint4 (&tt)[4], const uint4 &w, int i) {
            uint4 r;
            r.x = fma2<kBF16>(tt[3].x, w.w, fma2<kBF16>(tt[2].x, w.z, fma2<kBF16>(tt[1].x, w.y, mul2<kBF16>(tt[0].x, w.x))));
            r.y = fma2<kBF16>(tt[3].y, w.w, fma2<kBF16>(tt[2].y, w.z, fma2<kBF16>(tt[1].y, w.y, mul2<kBF16>(tt[0].y, w.x))));
            r.z = fma2<kBF16>(tt[3].z, w.w, fma2<kBF16>(tt[2].z, w.z, fma2<kBF16>(tt[1].z, w.y, mul2<kBF16>(tt[0].z, w.x))));
            r.w = fma2<kBF16>(tt[3].w, w.w, fma2<kBF16>(tt[2].w, w.z, fma2<kBF16>(tt[1].w, w.y, mul2<kBF16>(tt[0].w, w.x))));
            *reinterpret_cast<uint4 *>(xrow + (i & 1) * 8 * kChunkA + (i >> 1) * 64) = r;
          };
          issue(load_desc(0, wa), ta);
          offn = load_desc(1, wb);
#pragma unroll 1
          for (int i = 0; i < 16; i += 2) {
            issue(offn, tb);
            uint4 wn;
            if (i + 2 < 16) offn = load_desc(i + 2, wn);
            blend(ta, wa, i);
            if (i + 2 < 16) {
              issue(offn, ta);
              wa = wn;
              offn = load_desc(i + 3, wn);
            }
            blend(tb, wb, i + 1);
            wb = wn;
          }
        } else {
          // cooperative gather, fp32 planes: one step = 2 consecutive rows x 1 plane x 16 channel quads (a texel = 256 bytes =
          // 16 lanes x 16 B); fp32 blend in ATen's order (nw, ne, sw, se), then hi / lo split into X_hi / X_lo
          const float4 *planes = reinterpret_cast<const float4 *>(P.planes_cl);
          const int quad = lane & 15, rsub = lane >> 4;
          const uint8_t *sp0 = stage + (warp * kRowsPerWarp + rsub) * kStageRow;
          uint8_t *xrow = X + (quad >> 1) * kChunkA + (warp * kRowsPerWarp + rsub) * 16 + (quad & 1) * 8;
          const size_t row_pitch = (size_t)Wp * 16;
          float4 ta[4], tb[4], wa, wb;
          int offn;
          auto load_desc = [&](int i, float4 &w) {   // step i: rows 2*(i>>1), +1; plane i&1
            const uint8_t *sp = sp0 + (i >> 1) * 2 * kStageRow;
            w = *reinterpret_cast<const float4 *>(sp + 16 + (i & 1) * 16);
            return *reinterpret_cast<const int *>(sp + (i & 1) * 4);
          };
          auto issue = [&](int off, float4 (&tt)[4]) {
            const float4 *tp = planes + (size_t)off * 16 + quad;
            tt[0] = __ldg(tp), tt[1] = __ldg(tp + 16), tt[2] = __ldg(tp + row_pitch), tt[3] = __ldg(tp + row_pitch + 16);
          };
          auto blend = [&](const float4 (&tt)[4], const float4 &w, int i) {
            float r[4];
            const float *a0 = &tt[0].x, *a1 = &tt[1].x, *a2 = &tt[2].x, *a3 = &tt[3].x;
#pragma unroll
            for (int c = 0; c < 4; ++c) r[c] = fmaf(a3[c], w.w, fmaf(a2[c], w.z, fmaf(a1[c], w.y, a0[c] * w.x)));
            uint2 hi, lo;
            split2(r[0], r[1], hi.x, lo.x);
            split2(r[2], r[3], hi.y, lo.y);
            uint8_t *dst = xrow + (i & 1) * 8 * kChunkA + (i >> 1) * 32;
            *reinterpret_cast<uint2 *>(dst) = hi;
            *reinterpret_cast<uint2 *>(dst + kXBytes) = lo;
          };
          issue(load_desc(0, wa), ta);
          offn = load_desc(1, wb);
          constexpr int kSteps = kRowsPerWarp;        // 2 rows x 1 plane per step
#pragma unroll 1
          for (int i = 0; i < kSteps; i += 2) {
            issue(offn, tb);
            float4 wn;
            if (i + 2 < kSteps) offn = load_desc(i + 2, wn);
            blend(ta, wa, i);
            if (i + 2 < kSteps) {
              issue(offn, ta);
              wa = wn;
              offn = load_desc(i + 3, wn);
            }
            blend(tb, wb, i + 1);
            wb = wn;
          }
        }
        fence_async_all();                  // X (this CTA's shared memory) is read by the async proxy of an MMA the LEADER issues
        mbar_arrive_local(bar_full);        // the non-leader's consumer warp 0 relays the completed barrier to the leader
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// consumer
// ------------------------------------------------------------------------------------------------
// split mode: relu + hi / lo split of one 128-column fp32 accumulator row into the A operand of the next layer at tm_h:
// hi pairs in columns [0,64), the constant bias pair (1, 0) at 64 (65..71 zero), lo pairs in [72,136)
__device__ __forceinline__ void hidden_epilogue_split(uint32_t tm_acc_row, uint32_t tm_h_row) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t r[32];
    HAV_TMEM_LD32(r, tm_acc_row + q * 32);
    tmem_wait_ld();
    uint32_t vh[16], vl[16];
#pragma unroll
    for (int c = 0; c < 16; ++c)
      split2(fmaxf(__uint_as_float(r[2 * c]), 0.0f), fmaxf(__uint_as_float(r[2 * c + 1]), 0.0f), vh[c], vl[c]);
    HAV_TMEM_ST16(tm_h_row + q * 16, vh);
    HAV_TMEM_ST16(tm_h_row + 72 + q * 16, vl);
  }
  uint32_t one[8] = {0x3C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  HAV_TMEM_ST8(tm_h_row + 64, one);
  tmem_wait_st();
  tc_fence_before();
}

template <int kMode, bool kCheck>
__device__ __forceinline__ void consumer_loop(const RenderDev &P, int num_ray_blocks, int iters, uint32_t smem_base, uint32_t tmem_base,
                                              int pair, int warp, int t, bool leader) {
  using C = Cfg<kMode>;
  constexpr bool kBF16 = C::bf16, kSplit = C::split;
  const bool issuer = warp == 0 && leader;
  const uint32_t bars = smem_base + C::smBar + pair * kBarsPerPair * 8;
  const uint32_t bar_full = bars + kBarXFull * 8, bar_free = bars + kBarXFree * 8, bar_mma = bars + kBarMma * 8,
                 bar_zfine = bars + kBarZFine * 8, bar_ready = bars + kBarReady * 8;
  const uint32_t bar_peer_full = bars + kBarPeerFull * 8;
  const uint32_t bar_ready_leader = map_to_cta(bar_ready, 0), bar_peer_full_leader = map_to_cta(bar_peer_full, 0);
  const bool relay = warp == 0 && !leader;
  const uint32_t X_addr = smem_base + C::smX + pair * C::sets * kXBytes, C_addr = smem_base + C::smConst;
  const uint32_t W_addr = smem_base + C::smW;
  const uint32_t tm_acc0 = tmem_base + pair * 256, tm_acc1 = tm_acc0 + 128;
  const uint32_t tm_h = tmem_base + 256;                       // split mode only (one pair): hidden activations hi | lo
  const uint32_t tm_lane = (uint32_t)(warp * 32) << 16;
  constexpr uint32_t kIdesc128 = instr_desc2(128, kBF16), kIdescH = instr_desc2(kNH3, kBF16);
  uint32_t n = 0, mma_phase = 0, ready_phase = 0, sat = 0;

  // one layer's MMAs.  A: X (shared memory, kSS) or hidden activations (TMEM); B: this layer's half weight image(s)
  auto issue_l0 = [&]() {
    uint32_t acc = 0;
#pragma unroll
    for (int term = 0; term < (kSplit ? 3 : 1); ++term) {
      const uint32_t xa = X_addr + (term == 1 ? kXBytes : 0);                    // term 1: A_lo
      const uint32_t wb = W_addr + kW0h + (term == 2 ? kWHalfBytes : 0);         // term 2: W_lo
#pragma unroll
      for (int k = 0; k < kK0 / 16; ++k) {
        if (term == 1 && k >= kXChunks / 2) continue;                            // lo of the constant bias column is zero
        const uint32_t a_addr = k < kXChunks / 2 ? xa + 2 * k * kChunkA : C_addr;
        umma_ss2(tm_acc0, smem_desc(a_addr, kChunkA, 128), smem_desc(wb + 2 * k * kHalfChunkB, kHalfChunkB, 128), kIdesc128, acc);
        acc = 1;
      }
    }
  };
  auto issue_hidden2 = [&](uint32_t tm_d, uint32_t tm_a, int w_off, int chunk_b, uint32_t idesc) {
    uint32_t acc = 0;
#pragma unroll
    for (int term = 0; term < (kSplit ? 3 : 1); ++term) {
      const uint32_t ta = tm_a + (term == 1 ? 72 : 0);
      const uint32_t wb = W_addr + w_off + (term == 2 ? kWHalfBytes : 0);
#pragma unroll
      for (int k = 0; k <= kHid / 16; ++k) {
        if (term == 1 && k == kHid / 16) continue;
        umma_ts2(tm_d, ta + k * 8, smem_desc(wb + 2 * k * chunk_b, chunk_b, 128), idesc, acc);
        acc = 1;
      }
    }
  };
  // every consumer thread of both CTAs has finished its TMEM accesses of this phase -> the leader may issue.  Local: the 128
  // threads of this CTA's consumer group meet at a named barrier; the non-leader's warp 0 then sends ONE remote arrive.
  auto arrive_ready = [&]() {
    bar_named(1 + pair);
    if (relay) {
      if (elect_one()) mbar_arrive_cluster(bar_ready_leader);
      __syncwarp();
    }
  };
  auto wait_ready = [&]() {
    mbar_wait_cluster(bar_ready, ready_phase & 1);
    ++ready_phase;
    tc_fence_after();
  };

  for (int it = 0; it < iters; ++it) {
    const int rb = block_of<C::pairs>(it, pair);
    const int g = rb < num_ray_blocks ? rb * kRaysPerBlock + t : P.total_rays;
    const Ray ray = load_ray(P, g);
    const int gi = ray.valid ? g : 0;
    const int slot = blockIdx.x * C::pairs + pair;
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
      for (int s = 0; s < S; ++s, ++n) {
        if (issuer) {
          if (n > 0) wait_ready();                       // both CTAs have read the previous head accumulator (acc0 is free)
          mbar_wait(bar_full, n & 1);                    // this CTA's X tile is in place ...
          mbar_wait_cluster(bar_peer_full, n & 1);       // ... and the peer's (relayed)
          tc_fence_after();
          if (elect_one()) {
            issue_l0();
            umma_commit2(bar_free);
            umma_commit2(bar_mma);
          }
          __syncwarp();
        }
        if (relay) {                                      // non-leader: forward "my X tile is full" to the leader
          mbar_wait(bar_full, n & 1);
          if (elect_one()) mbar_arrive_cluster(bar_peer_full_leader);
          __syncwarp();
        }
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
        const float nz = noise != nullptr ? __ldg(noise + (size_t)gi * S + s) : 0.0f;

        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        if constexpr (kSplit) hidden_epilogue_split(tm_acc0 + tm_lane, tm_h + tm_lane);
        else hidden_epilogue<kBF16, true, kCheck>(tm_acc0 + tm_lane, nullptr, t, &sat);
        arrive_ready();
        if (issuer) {
          wait_ready();
          if (elect_one()) {
            issue_hidden2(tm_acc1, kSplit ? tm_h : tm_acc0, kW1h, kHalfChunkB, kIdesc128);
            umma_commit2(bar_mma);
          }
          __syncwarp();
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        if constexpr (kSplit) hidden_epilogue_split(tm_acc1 + tm_lane, tm_h + tm_lane);
        else hidden_epilogue<kBF16, true, kCheck>(tm_acc1 + tm_lane, nullptr, t, &sat);
        arrive_ready();
        if (issuer) {
          wait_ready();
          if (elect_one()) {
            issue_hidden2(tm_acc0, kSplit ? tm_h : tm_acc1, kWHh, kHalfChunkBH, kIdescH);
            umma_commit2(bar_mma);
          }
          __syncwarp();
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        // composite (utils/nerf_util.py:28-73): cols 64 = sigma, 65..67 = rgb logits, 0..63 = features
        {
          uint32_t h[4];
          HAV_TMEM_LD4(h, tm_acc0 + tm_lane + kRgbFeat);
          tmem_wait_ld();
          const float w = kSplit ? cs.step<false>(__uint_as_float(h[0]), nz, dist * ray.dnorm, z)
                                 : cs.step<true>(__uint_as_float(h[0]), nz, dist * ray.dnorm, z);
          if (pass == 0 && npass == 2) wcol[s * kRaysPerBlock] = w;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float sg = kSplit ? sigmoidf_exact(__uint_as_float(h[1 + j])) : sigmoidf_fast(__uint_as_float(h[1 + j]));
            sums[j] = kSplit ? __fadd_rn(sums[j], __fmul_rn(w, sg)) : fmaf(w, sg, sums[j]);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t r[32];
            HAV_TMEM_LD32(r, tm_acc0 + tm_lane + q * 32);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) sums[3 + q * 32 + c] = fmaf(w, __uint_as_float(r[c]), sums[3 + q * 32 + c]);
          }
        }
        tc_fence_before();
        arrive_ready();     // this row has read the head accumulator; the next L0 (or the kernel's end) may proceed
      }
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
      if (pass == 0 && npass == 2) {
        auto zc = [&](int s) { return coarse_z(P, ray, gi, s); };
        sample_pdf_merge(zc, P.Sc, P.nfine, wcol, kRaysPerBlock, P.u_rand != nullptr ? P.u_rand + (size_t)gi * P.nfine : nullptr, zcol,
                         (ray.valid && P.pdf_inds != nullptr) ? P.pdf_inds + (size_t)g * P.nfine : nullptr);
        if (ray.valid && P.z_fine != nullptr)
          for (int j = 0; j < P.Sf; ++j) P.z_fine[(size_t)g * P.Sf + j] = zcol[j * kRaysPerBlock];
        __threadfence_block();
        mbar_arrive_local(bar_zfine);
      }
    }
  }
  if (issuer && n > 0) wait_ready();     // drain the last phase so that no arrival is in flight towards a CTA that has exited
  if (kCheck && sat != 0 && P.status != nullptr) atomicOr(P.status, 2);
}

template <int kMode, bool kCheck>
__global__ void __launch_bounds__(kThreads3, 1) render_tc3_kernel(const RenderDev P, int num_ray_blocks, int iters) {
  using C = Cfg<kMode>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t rank = cluster_ctarank();
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + C::smBar + 112);

  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + C::smBar + 112), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (tid == 32) {
    for (int p = 0; p < C::pairs; ++p) {
      const uint32_t b = smem_base + C::smBar + p * kBarsPerPair * 8;
      mbar_init(b + kBarXFull * 8, C::split ? 256 : 128);   // local: the producer threads of this CTA
      mbar_init(b + kBarPeerFull * 8, 1);  // leader's copy: the peer's relay thread ("the peer's X tile is full")
      mbar_init(b + kBarXFree * 8, 1);     // tcgen05.commit, multicast
      mbar_init(b + kBarMma * 8, 1);       // tcgen05.commit, multicast
      mbar_init(b + kBarZFine * 8, 128);   // local: consumer -> producer, fine depths written
      mbar_init(b + kBarReady * 8, 1);     // leader's copy: the peer's relay thread ("the peer's consumers passed this phase")
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    // this CTA's half of every weight matrix: rows [rank * N/2, (rank + 1) * N/2) of the [K/8][N][8] images (head: 96 rows,
    // the image holds 80 -> zero fill).  Split mode: hi image, then lo image.
    for (int set = 0; set < C::sets; ++set) {
      const uint4 *src = reinterpret_cast<const uint4 *>(P.wimg + (size_t)set * kWImgBytes);
      uint4 *dst = reinterpret_cast<uint4 *>(smem + C::smW + set * kWHalfBytes);
      for (int i = tid; i < kWHalfBytes / 16; i += kThreads3) {
        const int byte = i * 16;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (byte < kWHh) {
          const int m = byte < kW1h ? 0 : 1;
          const int rel = byte - (m ? kW1h : kW0h), chunk = rel / kHalfChunkB, row = (rel % kHalfChunkB) / 16;
          v = __ldg(src + ((m ? kW1Off : kW0Off) + chunk * kChunkB + ((int)rank * 64 + row) * 16) / 16);
        } else {
          const int rel = byte - kWHh, chunk = rel / kHalfChunkBH, row = (int)rank * (kNH3 / 2) + (rel % kHalfChunkBH) / 16;
          if (row < kNH) v = __ldg(src + (kWHOff + chunk * kChunkBH + row * 16) / 16);
        }
        dst[i] = v;
      }
    }
    if (tid < 128) {
      const uint32_t one = C::bf16 ? 0x3F80u : 0x3C00u;
      *reinterpret_cast<uint4 *>(smem + C::smConst + tid * 16) = make_uint4(one, 0u, 0u, 0u);
      *reinterpret_cast<uint4 *>(smem + C::smConst + kChunkA + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_async_all();
  tc_fence_before();
  __syncthreads();
  cluster_sync();          // both CTAs: barriers initialised, weights in place, TMEM allocated -- before any remote arrive / MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int t = tid & 127;
  // 16-bit modes: warps 0-3 / 4-7 = consumers of pair 0 / 1, warps 8-11 / 12-15 = producers of pair 0 / 1.
  // split mode (one pair): warps 0-3 consumers (4-7 idle), warps 8-15 producers (two threads per row).
  const int pair = (warp_u >> 2) & 1;
  if (warp_u < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
    if (pair < C::pairs) consumer_loop<kMode, kCheck>(P, num_ray_blocks, iters, smem_base, tmem_base, pair, warp_u & 3, t, rank == 0);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
    if (C::split) producer_loop<kMode>(P, num_ray_blocks, iters, smem, smem_base, 0, tid - 256);      // all eight warps, one pair
    else producer_loop<kMode>(P, num_ray_blocks, iters, smem, smem_base, pair, t);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();          // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
}

}  // namespace tc3

static int tc3_clusters(int num_ray_blocks, int pairs) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int per_cluster = 2 * pairs;
  int want = (num_ray_blocks + per_cluster - 1) / per_cluster;
  int maxc = sms / 2;
  return want < maxc ? (want > 0 ? want : 1) : maxc;
}
int tc3_scratch_slots(int num_ray_blocks) {
  // slots are indexed blockIdx.x * pairs + pair: bound by the larger of the two configurations
  int a = tc3_clusters(num_ray_blocks, 2) * 2 * 2, b = tc3_clusters(num_ray_blocks, 1) * 2;
  return a > b ? a : b;
}

template <int kMode, bool kCheck>
static cudaError_t launch_tc3(const RenderDev &P, int num_ray_blocks, cudaStream_t st) {
  using C = tc3::Cfg<kMode>;
  auto kern = tc3::render_tc3_kernel<kMode, kCheck>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smBytes);
  if (e != cudaSuccess) return e;
  const int clusters = tc3_clusters(num_ray_blocks, C::pairs);
  const int per_iter = clusters * 2 * C::pairs;
  const int iters = (num_ray_blocks + per_iter - 1) / per_iter;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * 2), cfg.blockDim = dim3(tc3::kThreads3), cfg.dynamicSmemBytes = C::smBytes, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, P, num_ray_blocks, iters);
}

// mode: 0 fp16, 1 bf16, 2 fp16 hi + lo split (fp32-class)
cudaError_t launch_render_16_v3(const RenderDev &P, int num_ray_blocks, int mode, cudaStream_t st) {
  if (mode == tc3::kModeSplit) return launch_tc3<tc3::kModeSplit, false>(P, num_ray_blocks, st);
  if (mode == tc3::kModeBF16) return launch_tc3<tc3::kModeBF16, false>(P, num_ray_blocks, st);
  if (P.status != nullptr) return launch_tc3<tc3::kModeF16, true>(P, num_ray_blocks, st);
  return launch_tc3<tc3::kModeF16, false>(P, num_ray_blocks, st);
}

}  // namespace hav
