// Tensor-core fused render kernel, v2: producer / consumer warp specialisation.
//
// One persistent CTA per SM, 512 threads, two independent PAIRS.  A pair = one consumer warpgroup (4 warps, owns
// 256 TMEM columns and the per-ray composite state) + one producer group (4 warps) marching the same block of 128
// rays; each march step is one M = 128 tile (row m = ray m at sample s).
//
//   producer, thread-per-row : depth -> o+d*z -> 2-bone skinning warp (Skinning_Field.py:70-98) -> bilinear tap
//                              descriptors of both planes (offset + 4 packed 16-bit weights) -> PE (registers)
//   producer, warp-cooperative: bi-plane gather, 8 lanes per 128-byte channels-last texel, HFMA2 blend, one
//                              STS.128 per lane into the K-major A operand X (shared memory); PE rows; arrive x_full
//   consumer, one thread      : wait x_full -> tcgen05.mma L0 (A = X in smem) -> commit x_free (producer may refill X)
//   consumer, thread-per-row : TMEM -> relu -> 16-bit back into the accumulator's own columns -> tcgen05.mma L1
//                              (A in TMEM) -> same -> head GEMM (A in TMEM) -> alpha composite out of TMEM
//
// The producer of a pair works on tile s+1 (its loads are latency bound) while the consumer runs the three GEMMs,
// two epilogues and the composite of tile s; the two pairs interleave on the tensor pipe.  Register budget is
// re-balanced with setmaxnreg (consumers keep 67 composite accumulators + 48 staging registers per thread).
#include "tc_common.cuh"

namespace hav {
namespace tc2 {

using namespace tc;

// Optional phase timing (build with HAV_NVCC_DEFS=-DHAV_TC_TIMING): thread 0 of pair 0 of CTA 0 accumulates the
// cycles it spends in each phase into P.zbuf-independent global counters (see scripts/time_phases.py).
#ifdef HAV_TC_TIMING
__device__ unsigned long long g_phase_cycles[32];
#define TICK(var) long long var = clock64()
#define TOCK(idx, var)                                                        \
  do {                                                                        \
    long long _n = clock64();                                                 \
    if (t == 0 && pair == 0 && blockIdx.x == 0) g_phase_cycles[idx] += (unsigned long long)(_n - var); \
    var = _n;                                                                 \
  } while (0)
#else
#define TICK(var)
#define TOCK(idx, var)
#endif

constexpr int kPairs = 2;
constexpr int kThreads2 = 512;
constexpr int kConsumerRegs = 168, kProducerRegs = 88;
constexpr int kXChunks = 22;                               // 16 feature chunks + 6 PE chunks
constexpr int kXBytes = kXChunks * kChunkA;                // 45408
constexpr int kConstBytes = 2 * kChunkA;                   // ones chunk [1,0..0] + zero chunk, shared by both pairs
constexpr int kStageRow = 48;                              // off0, off1, pad, pad | 4 weights plane 0 | 4 weights plane 1
constexpr int kStageBytes2 = 128 * kStageRow;              // one buffer; two buffers per pair
constexpr int kSmX = kWImgBytes;
constexpr int kSmConst = kSmX + kPairs * kXBytes;
constexpr int kSmStage = kSmConst + kConstBytes;
constexpr int kSmBar = kSmStage + kPairs * 2 * kStageBytes2;
constexpr int kSmemBytes2 = kSmBar + 128;
static_assert(kSmemBytes2 <= 232448, "shared memory budget");
// barrier slots per pair (8 bytes each)
constexpr int kBarXFull = 0, kBarXFree = 1, kBarMma = 2, kBarZFine = 3, kBarsPerPair = 4;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bar_named(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// tap base + the four bilinear weights, pre-packed as duplicated 16-bit pairs for the gather's HFMA2s
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

// ------------------------------------------------------------------------------------------------
// producer: fills X(s) for every tile of this pair's ray blocks
// ------------------------------------------------------------------------------------------------
template <bool kBF16>
__device__ __forceinline__ void producer_loop(const RenderDev &P, int num_ray_blocks, uint8_t *smem, uint32_t smem_base,
                                              int pair, int t) {
  const int warp = t >> 5, lane = t & 31;
  uint8_t *X = smem + kSmX + pair * kXBytes;
  uint8_t *stage_base = smem + kSmStage + pair * 2 * kStageBytes2;
  const uint32_t bars = smem_base + kSmBar + pair * kBarsPerPair * 8;
  const uint32_t bar_full = bars + kBarXFull * 8, bar_free = bars + kBarXFree * 8, bar_zfine = bars + kBarZFine * 8;
  const int Wp = P.PW + kPadLo + kPadHi;
  const uint4 *planes = reinterpret_cast<const uint4 *>(P.planes_cl);
  const int oct = lane & 7;
  uint32_t n = 0;          // tiles produced so far by this pair
  uint32_t zfine_phase = 0;

  for (int rb = blockIdx.x * kPairs + pair; rb < num_ray_blocks; rb += gridDim.x * kPairs) {
    const int g = rb * kRaysPerBlock + t;
    const Ray ray = load_ray(P, g);
    const int gi = ray.valid ? g : 0;
    float Tm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Tm[i] = __ldg(P.invT + (size_t)ray.b * 12 + i);
    const int slot = blockIdx.x * kPairs + pair;
    const float *zcol = P.zbuf + (size_t)slot * P.Sf * kRaysPerBlock + t;
    const int npass = P.nfine > 0 ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
      const int S = pass == 0 ? P.Sc : P.Sf;
      if (pass == 1) {   // fine depths come from the consumer's sample_pdf (global scratch)
        mbar_wait(bar_zfine, zfine_phase);
        zfine_phase ^= 1;
      }
#pragma unroll 1
      for (int s = 0; s < S; ++s, ++n) {
        uint8_t *stage = stage_base + (n & 1) * kStageBytes2;
        TICK(tk);
        // ---- row thread: depth, point, skinning warp, tap descriptors, PE
        const float z = pass == 0 ? coarse_z(P, ray, gi, s) : __ldcg(zcol + s * kRaysPerBlock);
        float p[3], pc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) p[j] = fmaf(ray.d[j], z, ray.o[j]);
        skin_warp(P, Tm, p, pc);
        {
          float qx = pc[0] * P.ps[0] + P.pt[0], qy = pc[1] * P.ps[1] + P.pt[1], qz = pc[2] * P.ps[2] + P.pt[2];
          int off0, off1;
          uint32_t w0[4], w1[4];
          plane_taps_packed<kBF16>(qx, qy, P.PH, P.PW, ray.b, off0, w0);
          plane_taps_packed<kBF16>(qz, qy, P.PH, P.PW, P.B + ray.b, off1, w1);
          uint4 *sr = reinterpret_cast<uint4 *>(stage + t * kStageRow);
          sr[0] = make_uint4(off0, off1, 0u, 0u);
          sr[1] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
          sr[2] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        }
        uint32_t pk[24];
        {
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
        TOCK(0, tk);
        bar_named(3 + pair);                                  // tap descriptors of all 128 rows are visible
        TOCK(1, tk);
        if (n > 0) mbar_wait(bar_free, (n - 1) & 1);          // L0 of the previous tile has finished reading X
        TOCK(2, tk);
#pragma unroll
        for (int c = 0; c < 6; ++c)
          *reinterpret_cast<uint4 *>(X + (16 + c) * kChunkA + t * 16) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        // ---- cooperative gather: one step = 4 consecutive rows x 1 plane x 8 channel octets (lane = row-in-group*8 +
        //      octet).  Neighbouring rays mostly share their texels (a texel spans ~4.5 pixels, and all rays of a tile
        //      sit at the same depth), so a step's 32 tap loads usually fall into 1-2 128-byte lines instead of 4.
        //      Software pipelined by hand: the tap loads of step i+1 are in flight while step i is blended.
        {
          const int rsub = lane >> 3;
          const uint8_t *sp0 = stage + (warp * 32 + rsub) * kStageRow;
          uint8_t *xrow = X + oct * kChunkA + (warp * 32 + rsub) * 16;
          const size_t row_pitch = (size_t)Wp * 8;
          uint4 ta[4], tb[4], wa, wb;
          int offn;
          // step i: rows 4*(i>>1) .. +3, plane i&1
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
        }
        TOCK(3, tk);
        fence_async_smem();
        mbar_arrive(bar_full);
        TOCK(4, tk);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// consumer: GEMMs, epilogues, composite, hierarchical resampling, output
// ------------------------------------------------------------------------------------------------
template <bool kBF16, bool kCheck>
__device__ __forceinline__ void consumer_loop(const RenderDev &P, int num_ray_blocks, uint32_t smem_base, uint32_t tmem_base,
                                              int pair, int warp, int t) {
  const bool issuer = warp == 0;   // `pair` and `warp` are warp-uniform (shuffled from lane 0 by the caller)
  const uint32_t bars = smem_base + kSmBar + pair * kBarsPerPair * 8;
  const uint32_t bar_full = bars + kBarXFull * 8, bar_free = bars + kBarXFree * 8, bar_mma = bars + kBarMma * 8,
                 bar_zfine = bars + kBarZFine * 8;
  const uint32_t X_addr = smem_base + kSmX + pair * kXBytes, C_addr = smem_base + kSmConst;
  const uint32_t W0_addr = smem_base + kW0Off, W1_addr = smem_base + kW1Off, WH_addr = smem_base + kWHOff;
  const uint32_t tm_acc0 = tmem_base + pair * 256, tm_acc1 = tm_acc0 + 128;
  const uint32_t tm_lane = (uint32_t)(warp * 32) << 16;
  constexpr uint32_t kIdesc128 = instr_desc(128, kBF16), kIdescH = instr_desc(kNH, kBF16);
  uint32_t n = 0, mma_phase = 0, sat = 0;

  for (int rb = blockIdx.x * kPairs + pair; rb < num_ray_blocks; rb += gridDim.x * kPairs) {
    const int g = rb * kRaysPerBlock + t;
    const Ray ray = load_ray(P, g);
    const int gi = ray.valid ? g : 0;
    const int slot = blockIdx.x * kPairs + pair;
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
        // ---- L0: [128 x 192] x [192 x 128] -> acc0, A = X (smem) + the constant bias chunk pair
        TICK(tk);
        if (issuer) {   // warp-uniform: descriptors stay in uniform registers, one elected lane issues
          mbar_wait(bar_full, n & 1);
          TOCK(8, tk);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kK0 / 16; ++k) {
              const uint32_t a_addr = k < kXChunks / 2 ? X_addr + 2 * k * kChunkA : C_addr;
              umma_ss(tm_acc0, smem_desc(a_addr, kChunkA, 128), smem_desc(W0_addr + 2 * k * kChunkB, kChunkB, 128), kIdesc128, k > 0);
            }
            umma_commit(bar_free);
            umma_commit(bar_mma);
          }
          __syncwarp();
        }
        // depth bookkeeping of this sample while the GEMM runs (nerf_trainer.py:129-141, nerf_util.py:36-38)
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

        TOCK(9, tk);
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        TOCK(10, tk);
        hidden_epilogue<kBF16, true, kCheck>(tm_acc0 + tm_lane, nullptr, t, &sat);
        TOCK(11, tk);
        bar_named(1 + pair);
        TOCK(12, tk);
        if (issuer) {
          tc_fence_after();
          if (elect_one()) {
            issue_hidden<true>(tm_acc1, tm_acc0, 0, W1_addr, kChunkB, kIdesc128);
            umma_commit(bar_mma);
          }
          __syncwarp();
        }
        TOCK(13, tk);
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        TOCK(14, tk);
        hidden_epilogue<kBF16, true, kCheck>(tm_acc1 + tm_lane, nullptr, t, &sat);
        TOCK(15, tk);
        bar_named(1 + pair);
        TOCK(16, tk);
        if (issuer) {
          tc_fence_after();
          if (elect_one()) {
            issue_hidden<true>(tm_acc0, tm_acc1, 0, WH_addr, kChunkBH, kIdescH);
            umma_commit(bar_mma);
          }
          __syncwarp();
        }
        TOCK(17, tk);
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        TOCK(18, tk);
        // ---- composite (utils/nerf_util.py:28-73): cols 64 = sigma, 65..67 = rgb logits, 0..63 = features
        {
          uint32_t h[4];
          HAV_TMEM_LD4(h, tm_acc0 + tm_lane + kRgbFeat);
          tmem_wait_ld();
          const float w = cs.step<true>(__uint_as_float(h[0]), nz, dist * ray.dnorm, z);
          if (pass == 0 && npass == 2) wcol[s * kRaysPerBlock] = w;
#pragma unroll
          for (int j = 0; j < 3; ++j) sums[j] = fmaf(w, sigmoidf_fast(__uint_as_float(h[1 + j])), sums[j]);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t r[32];
            HAV_TMEM_LD32(r, tm_acc0 + tm_lane + q * 32);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) sums[3 + q * 32 + c] = fmaf(w, __uint_as_float(r[c]), sums[3 + q * 32 + c]);
          }
        }
        TOCK(19, tk);
        tc_fence_before();
        bar_named(1 + pair);   // every row has read the head accumulator before the next L0 overwrites acc0
        TOCK(20, tk);
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
        __threadfence_block();
        mbar_arrive(bar_zfine);   // the producer may start the fine pass of this block
      }
    }
  }
  if (kCheck && sat != 0 && P.status != nullptr) atomicOr(P.status, 2);   // a hidden activation was clipped to 65504
}

template <bool kBF16, bool kCheck>
__global__ void __launch_bounds__(kThreads2, 1) render_tc2_kernel(const RenderDev P, int num_ray_blocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const uint32_t smem_base = smem_u32(smem);
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSmBar + 96);

  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kSmBar + 96), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int p = 0; p < kPairs; ++p) {
      const uint32_t b = smem_base + kSmBar + p * kBarsPerPair * 8;
      mbar_init(b + kBarXFull * 8, 128);   // every producer thread arrives after its own stores + proxy fence
      mbar_init(b + kBarXFree * 8, 1);     // tcgen05.commit
      mbar_init(b + kBarMma * 8, 1);       // tcgen05.commit
      mbar_init(b + kBarZFine * 8, 128);   // every consumer thread arrives after writing its fine depths
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.wimg);
    uint4 *dst = reinterpret_cast<uint4 *>(smem);
    for (int i = tid; i < kWImgBytes / 16; i += kThreads2) dst[i] = __ldg(src + i);
    if (tid < 128) {
      const uint32_t one = kBF16 ? 0x3F80u : 0x3C00u;
      *reinterpret_cast<uint4 *>(smem + kSmConst + tid * 16) = make_uint4(one, 0u, 0u, 0u);
      *reinterpret_cast<uint4 *>(smem + kSmConst + kChunkA + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform in the compiler's eyes
  const int pair = (warp_u >> 2) & 1, t = tid & 127;
  if (warp_u < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsumerRegs));
    consumer_loop<kBF16, kCheck>(P, num_ray_blocks, smem_base, tmem_base, pair, warp_u & 3, t);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
    producer_loop<kBF16>(P, num_ray_blocks, smem, smem_base, pair, t);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
}

}  // namespace tc2

int tc2_num_ctas(int num_ray_blocks) {
  int want = (num_ray_blocks + tc2::kPairs - 1) / tc2::kPairs;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  return want < sms ? (want > 0 ? want : 1) : sms;
}

template <bool kBF16, bool kCheck = false>
static cudaError_t launch_tc2(const RenderDev &P, int num_ray_blocks, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tc2::render_tc2_kernel<kBF16, kCheck>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmemBytes2);
  if (e != cudaSuccess) return e;
  tc2::render_tc2_kernel<kBF16, kCheck><<<tc2_num_ctas(num_ray_blocks), tc2::kThreads2, tc2::kSmemBytes2, st>>>(P, num_ray_blocks);
  return cudaGetLastError();
}

#ifdef HAV_TC_TIMING
extern "C" int hav_debug_phase_cycles(unsigned long long *out32, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out32, tc2::g_phase_cycles, sizeof(unsigned long long) * 32);
  if (e == cudaSuccess && reset) {
    unsigned long long z[32] = {0};
    e = cudaMemcpyToSymbol(tc2::g_phase_cycles, z, sizeof(z));
  }
  return (int)e;
}
#endif

cudaError_t launch_render_16_v2(const RenderDev &P, int num_ray_blocks, bool bf16, cudaStream_t st) {
  if (!bf16 && P.status != nullptr) return launch_tc2<false, true>(P, num_ray_blocks, st);   // HAV_RENDER_CHECK_RANGE
  return bf16 ? launch_tc2<true>(P, num_ray_blocks, st) : launch_tc2<false>(P, num_ray_blocks, st);
}

}  // namespace hav
