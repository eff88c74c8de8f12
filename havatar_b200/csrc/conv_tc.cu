// Tensor-core 2-D convolution for the StyleUNet blocks (model/styleUnet.py): ModulatedConv2d (:165-297), EqualConv2d
// (:88-123), with the modulation / demodulation / noise / bias / leaky-relu of StyledConv (:565-599), ToRGB (:602-628)
// and ConvLayer (:326-368) fused around one implicit GEMM on tcgen05.
//
// Formulation (the reference's own non-fused branch, styleUnet.py:225-251): the modulated convolution
//     y[b] = conv(x[b], scale * W * s[b]) * demod[b]          is computed as
//     y[b, co] = demod[b, co] * sum_{ci,kh,kw} (scale * W[co,ci,kh,kw]) * (s[b,ci] * x[b,ci,...])
// i.e. ONE shared fp16 weight matrix for the whole batch (no per-sample [B*Cout,Cin,k,k] weights, no grouped conv),
// s applied while the input tile is staged, demod applied in the epilogue.
//
// Implicit GEMM: M = 128 output positions (a 16 x 8 patch), N = up to 128 output channels, K = Cin * k * k.  The input
// patch with its halo is staged ONCE per 64-channel block into shared memory as [Cin/8][18 x 10 pixels][8] fp16; the
// A operand of tap (kh,kw) is that same buffer seen through a UMMA descriptor whose start address is shifted by
// (kh * 10 + kw) pixels and whose 8-row groups (one image row of the patch) are 10 pixels apart -- no im2col copy.
// Weights arrive per (tap, channel block) as 16 KB bulk copies (cp.async.bulk + mbarrier) from a pre-packed image.
// up = 2 (conv_transpose2d stride 2, :264-277) is polyphase: a tile is 128 INPUT positions (i,j); the four output
// parities out[2i+a, 2j+b] are four accumulators (TMEM column blocks) fed by the taps with kh = a (mod 2), kw = b (mod 2)
// reading x[i - kh/2, j - kw/2] -- 9 tap-GEMMs per tile like a plain 3x3 conv, no multiplications by inserted zeros.
// down = 2 (:279-287, after the blur; also the data gradient of the transposed convolution) is polyphase too: a tile is 128
// OUTPUT positions, the input patch is staged as its four parity planes x[2i+a, 2j+b] and tap (kh,kw) reads plane
// (kh&1, kw&1) shifted by (kh>>1, kw>>1) -- 9 tap-GEMMs per tile, no stride-1 positions computed and thrown away.
// Split-K on thread-block clusters: the 8 x 8 ... 32 x 32 layers of the network have 512 - 1024 input channels and a handful of
// output tiles, i.e. a few CTAs that each stream megabytes of weights while most of the chip idles.  There the channel blocks
// of one output tile are divided over the 2 / 4 / 8 CTAs of a cluster (gridDim.z): every CTA accumulates its share in TMEM,
// writes the partial tile to its own shared memory, and after a cluster barrier CTA r sums column slice r of all partial
// tiles through distributed shared memory, applies the fused tail and stores -- no global workspace, no atomics, a fixed
// summation order.
#include <map>
#include <mutex>
#include <tuple>

#include "tc_common.cuh"

namespace hav {
namespace conv {

using namespace tc;

constexpr int kTileH = 16, kTileW = 8;            // output patch = 128 GEMM rows, row m = (m / 8, m % 8)
constexpr int kCinBlk = 64;                       // channels per staged block (4 K-steps of 16)
constexpr int kNTileMax = 128;
constexpr int kBStagesMax = 16;                   // weight ring: slots of one (tap, channel block) slice each, as many as the budget holds
constexpr int kBRingBytes = 4 * kNTileMax * kCinBlk * 2;       // 64 KB: two CTAs per SM
constexpr int kBRingBytesDeep = 8 * kNTileMax * kCinBlk * 2;   // 128 KB: launches that cannot fill the chip with two CTAs per SM anyway
constexpr int kMaxHaloPx = (kTileH + 2) * (kTileW + 2);   // 180
constexpr int kASlotBytes = (kCinBlk / 8) * kMaxHaloPx * 16;   // 23040
// stride-2 (polyphase) variant: the input patch is staged as its four parity planes x[2i+a, 2j+b] of (16+1) x (8+1) pixels
constexpr int kDnHaloPx = (kTileH + 1) * (kTileW + 1);    // 153
constexpr int kASlotBytesDN = 4 * (kCinBlk / 8) * kDnHaloPx * 16;   // 78336
// shared memory: [weight ring][barriers 512 B][epilogue tables: out_scale, bias of the CTA's output channels, 1 KB][modulation of
// the two staged channel blocks 512 B][two A buffers]; offsets relative to the end of the ring (its size is a launch parameter)
constexpr int kSmB = 0;
constexpr int kOffBar = 0, kOffEpi = 512, kOffScale = kOffEpi + 2 * kNTileMax * 4, kOffA = kOffScale + 2 * kCinBlk * 4;
constexpr int kOffDump = kOffScale;                        // split-K partial tile: over the drained scale + A buffers (and beyond)
constexpr int kSmemTail = kOffA + 2 * kASlotBytes;         // 48128: + 64 KB ring = 113664, two CTAs per SM
constexpr int kSmemTailDN = kOffA + 2 * kASlotBytesDN;     // + 64 KB ring = 224256: one CTA per SM
constexpr int kThreads = 128;

struct ConvDev {
  int B, Cin, Cout, H, W, Ho, Wo, k, up, down, pad, act;
  int n_tile, n_tiles, kblocks, tiles_x, tiles_y, tmem_cols;
  int b_stages, ring_bytes;   // weight ring geometry (slot = one slice of n_tile x 64 x 2 bytes)
  int ksplit;          // CTAs per cluster sharing one output tile (split over the channel blocks); 1 = no cluster
  const float *x, *in_scale, *out_scale, *noise, *bias;
  const uint8_t *wpack;
  float *out;
  const float *residual;   // out's shape and layout, added after the activation (or nullptr)
  float noise_weight;
  int noise_bstride;   // 0 = one noise image broadcast over the batch
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_conv(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// weights [Cout,Cin,k,k] fp32 -> per (n tile, channel block, tap): [8 chunks][n_tile rows][8] fp16, value scale * w
// (flipped in kh,kw when flip != 0: the transposed convolution).  Channels / rows beyond Cin / Cout are zero.
// One thread per (output row n, 8-channel K chunk): it reads the 8 x k*k source values (a contiguous 8*k*k run when the weight
// is [Cout,Cin,k,k]; k*k-float runs that neighbouring threads continue when it is [Cin,Cout,k,k]) and writes one 16-byte unit
// per tap, consecutive threads -> consecutive units.
template <bool kBF16, int taps>
__global__ void __launch_bounds__(128) pack_conv_weights_kernel(const float *__restrict__ w, uint16_t *__restrict__ out, int Cout, int Cin,
                                                                int n_tile, int n_tiles, int kblocks, float scale, int flip,
                                                                int transpose_io) {
  const long total = (long)n_tiles * kblocks * (kCinBlk / 8) * n_tile;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i;
    const int n = r % n_tile; r /= n_tile;
    const int chunk = r % (kCinBlk / 8); r /= (kCinBlk / 8);
    const int kb = r % kblocks; r /= kblocks;
    const int nt = (int)r;
    const int co = nt * n_tile + n, ci0 = kb * kCinBlk + chunk * 8;
    uint4 *dst = reinterpret_cast<uint4 *>(out) + (((long)nt * kblocks + kb) * taps * (kCinBlk / 8) + chunk) * n_tile + n;
    float v[taps][8];      // all 8 x taps loads are issued before the first conversion
#pragma unroll
    for (int tap = 0; tap < taps; ++tap) {
      const int src_tap = flip ? taps - 1 - tap : tap;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ci = ci0 + e;
        v[tap][e] = 0.0f;
        if (co < Cout && ci < Cin)
          v[tap][e] = __ldg(transpose_io ? w + ((long)ci * Cout + co) * taps + src_tap : w + ((long)co * Cin + ci) * taps + src_tap);
      }
    }
#pragma unroll
    for (int tap = 0; tap < taps; ++tap)
      dst[(long)tap * (kCinBlk / 8) * n_tile] =
          make_uint4(pack2<kBF16>(scale * v[tap][0], scale * v[tap][1]), pack2<kBF16>(scale * v[tap][2], scale * v[tap][3]),
                     pack2<kBF16>(scale * v[tap][4], scale * v[tap][5]), pack2<kBF16>(scale * v[tap][6], scale * v[tap][7]));
  }
}

// demod[b,co] = rsqrt(sum_{ci,kh,kw} (scale * W[co,ci,kh,kw] * s[b,ci])^2 + eps)   (styleUnet.py:256-258)
__global__ void __launch_bounds__(128) modconv_demod_kernel(const float *__restrict__ w, const float *__restrict__ s,
                                                            float *__restrict__ demod, int Cout, int Cin, int kk, float scale,
                                                            float eps) {
  const int co = blockIdx.x, b = blockIdx.y;
  float acc = 0.0f;
  for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
    const float *wp = w + ((long)co * Cin + ci) * kk;
    float q = 0.0f;
    for (int t = 0; t < kk; ++t) q = fmaf(wp[t], wp[t], q);
    const float sv = s[(long)b * Cin + ci] * scale;
    acc = fmaf(q, sv * sv, acc);
  }
  __shared__ float red[4];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) demod[(long)b * Cout + co] = rsqrtf(red[0] + red[1] + red[2] + red[3] + eps);
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  float v;
  // not volatile: the loads of one reduction are independent and may be scheduled together (ordering against the writers
  // comes from the cluster barrier, which is volatile with a memory clobber)
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
  return v;
}

// The fused tail of one output pixel x 8 consecutive output channels (co0 .. co0+7 of the whole layer): demodulation, noise, bias,
// leaky-relu * sqrt(2), then the store in the requested layout.
// epi = the CTA's shared-memory tables [n_tile] out_scale | [n_tile] bias, cl = co0's column inside the CTA's channel tile.
template <bool OUTCL>
__device__ __forceinline__ void conv_tail_store8(const ConvDev &P, const float *epi, int cl, int b, int oy, int ox, int co0, float nz,
                                                 float (&v)[8]) {
  const float4 s0 = *reinterpret_cast<const float4 *>(epi + cl), s1 = *reinterpret_cast<const float4 *>(epi + cl + 4);
  const float4 b0 = *reinterpret_cast<const float4 *>(epi + P.n_tile + cl), b1 = *reinterpret_cast<const float4 *>(epi + P.n_tile + cl + 4);
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, bi[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t = fmaf(v[j], sc[j], nz) + bi[j];
    if (P.act) t = (t > 0.0f ? t : 0.2f * t) * 1.41421356237309515f;
    v[j] = t;
  }
  if (OUTCL) {
    // channels-last fp16: 8 channels of this pixel = 16 contiguous bytes (Cout % 8 == 0 is checked on the host)
    if (co0 + 8 <= P.Cout) {
      const size_t off = (((size_t)b * P.Ho + oy) * P.Wo + ox) * P.Cout + co0;
      if (P.residual != nullptr) {
        const uint4 r = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint16_t *>(P.residual) + off));
        const __half2 *rh = reinterpret_cast<const __half2 *>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(rh[j]);
          v[2 * j] += f.x, v[2 * j + 1] += f.y;
        }
      }
      uint16_t *oc = reinterpret_cast<uint16_t *>(P.out) + off;
      *reinterpret_cast<uint4 *>(oc) = make_uint4(pack2<false>(v[0], v[1]), pack2<false>(v[2], v[3]), pack2<false>(v[4], v[5]), pack2<false>(v[6], v[7]));
    }
  } else {
    const size_t plane = (size_t)P.Ho * P.Wo;
    const size_t off = ((size_t)b * P.Cout) * plane + (size_t)oy * P.Wo + ox;
    float *ob = P.out + off;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (co0 + j < P.Cout) ob[(size_t)(co0 + j) * plane] = v[j] + (P.residual != nullptr ? __ldg(P.residual + off + (size_t)(co0 + j) * plane) : 0.0f);
  }
}

#ifdef HAV_CONV_TIMING      // phase stamps of CTA 0 (debug builds: HAV_NVCC_DEFS=-DHAV_CONV_TIMING, scripts/time_conv_phases.py)
__device__ unsigned long long g_conv_t[16];
__device__ __forceinline__ void conv_stamp(int i) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_conv_t[i] = t;
  }
}
#define HAV_CONV_STAMP(i, cond) do { if (cond) conv_stamp(i); } while (0)
#else
#define HAV_CONV_STAMP(i, cond) do { } while (0)
#endif

// CTA = 8 staging / epilogue warps + 1 control warp (weight ring + MMA issue).  The control warp never stages, so the
// halo patch of channel block kb+1 is being written while the tensor core works through the 9 taps of block kb.
constexpr int kStageThreads = 256;
constexpr int kThreadsV2 = kStageThreads + 32;

// INCL / OUTCL: input / output tensor is channels-last fp16 ([B,H,W,C] half) instead of NCHW fp32
template <bool kBF16, int KS, bool UPP, bool INCL, bool OUTCL, bool DN = false>
__global__ void __launch_bounds__(kThreadsV2) conv_tc_kernel(const ConvDev P) {
  static_assert(!DN || (KS == 3 && !UPP), "the polyphase stride-2 variant is 3x3 only");
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem_base = smem_u32(smem);
  const int kSmBar = P.ring_bytes + kOffBar, kSmA = P.ring_bytes + kOffA, kBStages = P.b_stages;
  const uint32_t bar_bfull = smem_base + kSmBar, bar_bfree = bar_bfull + kBStagesMax * 8, bar_afree = bar_bfree + kBStagesMax * 8,
                 bar_afull = bar_afree + 16, bar_acc = bar_afull + 16, bar_tmem = bar_acc + 32;    // tmem slot sits between them
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSmBar + 304);
  float *sscale = reinterpret_cast<float *>(smem + P.ring_bytes + kOffScale);   // [2][64] modulation of the staged channel block
  float *epi = reinterpret_cast<float *>(smem + P.ring_bytes + kOffEpi);        // [n_tile] out_scale | [n_tile] bias

  const int nt = blockIdx.y;
  int sp = blockIdx.x;
  const int tx = sp % P.tiles_x; sp /= P.tiles_x;
  const int ty = sp % P.tiles_y;
  const int b = sp / P.tiles_y;
  const int vy0 = ty * kTileH, vx0 = tx * kTileW;
  // staged patch: plain conv = tile + (KS-1) halo; polyphase transposed conv = tile + one row above / column left
  constexpr int hw = (UPP || DN) ? kTileW + 1 : kTileW + KS - 1, hh = (UPP || DN) ? kTileH + 1 : kTileH + KS - 1;
  constexpr int halo_px = hh * hw, chunk_bytes = halo_px * 16, taps = KS * KS;
  constexpr int a_slot = DN ? kASlotBytesDN : kASlotBytes;
  constexpr int n_chunks = DN ? 4 * (kCinBlk / 8) : kCinBlk / 8;   // DN: chunk index = parity plane * 8 + channel chunk
  // split-K: this CTA owns channel blocks [kb_begin, kb_end) of the tile (cluster rank = blockIdx.z)
  const int kb_per = (P.kblocks + P.ksplit - 1) / P.ksplit;
  const int kb_begin = min((int)blockIdx.z * kb_per, P.kblocks), kb_end = min(kb_begin + kb_per, P.kblocks), my_kblocks = kb_end - kb_begin;
  const int total_steps = my_kblocks * taps;
  const uint32_t b_bytes = (uint32_t)P.n_tile * kCinBlk * 2;     // = ring slot size

  HAV_CONV_STAMP(0, tid == 0);
  if (tid == 32) {
    for (int i = 0; i < kBStages; ++i) mbar_init(bar_bfull + i * 8, 1), mbar_init(bar_bfree + i * 8, 1);
    mbar_init(bar_afree, 1), mbar_init(bar_afree + 8, 1), mbar_init(bar_acc, 1), mbar_init(bar_tmem, 1);
    mbar_init(bar_afull, kStageThreads), mbar_init(bar_afull + 8, kStageThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid >= 64 && tid < 64 + P.n_tile) {      // epilogue tables of this CTA's output channels (read from shared memory per column)
    const int j = tid - 64, co = blockIdx.y * P.n_tile + j;
    epi[j] = (P.out_scale != nullptr && co < P.Cout) ? __ldg(P.out_scale + (size_t)(blockIdx.x / (P.tiles_x * P.tiles_y)) * P.Cout + co) : 1.0f;
    epi[P.n_tile + j] = (P.bias != nullptr && co < P.Cout) ? __ldg(P.bias + co) : 0.0f;
  }
  __syncthreads();
  HAV_CONV_STAMP(1, tid == 0);
  const uint8_t *wsrc = P.wpack + ((size_t)nt * P.kblocks + kb_begin) * taps * b_bytes;
  uint32_t tmem_acc = 0;

  if (warp_u == kStageThreads / 32) {
    // ================= control warp =================
    if (elect_one()) {
      for (int i = 0; i < kBStages && i < total_steps; ++i) {
        mbar_expect_tx(bar_bfull + i * 8, b_bytes);
        bulk_g2s(smem_base + kSmB + i * b_bytes, wsrc + (size_t)i * b_bytes, b_bytes, bar_bfull + i * 8);
      }
    }
    __syncwarp();
    // the accumulator columns are allocated here, off the staging warps' path (they need the address only for the epilogue and
    // pick it up through bar_tmem); the first weight slices are already in flight
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kSmBar + 304), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    tc_fence_before();
    __syncwarp();
    tc_fence_after();
    tmem_acc = *tmem_slot;
    if (elect_one()) {
      mbar_arrive_conv(bar_tmem);
      const uint32_t idesc = instr_desc(P.n_tile, kBF16);
      // ring position of the current step / of the previous step, kept incrementally (the ring depth is a launch parameter)
      int slot = 0, sphase = 0, ps = 0, pphase = 0, step = 0;
      for (int kb = 0; kb < my_kblocks; ++kb) {      // kb counts this CTA's channel blocks
        mbar_wait_spin(bar_afull + (kb & 1) * 8, (kb >> 1) & 1);
        tc_fence_after();
        const uint32_t A_addr = smem_base + kSmA + (kb & 1) * a_slot;
#pragma unroll 1
        for (int tap = 0; tap < taps; ++tap, ++step) {
          mbar_wait_spin(bar_bfull + slot * 8, sphase);
          HAV_CONV_STAMP(4, step == 0);
          tc_fence_after();
          const int kh = tap / KS, kw = tap - kh * KS;
          const uint32_t b0 = smem_base + kSmB + slot * b_bytes;
          uint32_t a0, d0;
          bool fresh;   // first contribution to its accumulator
          if (UPP) {    // out[2i+a, 2j+b] += x[i - kh/2, j - kw/2] * W[kh,kw],  a = kh & 1, b = kw & 1
            a0 = A_addr + ((1 - (kh >> 1)) * hw + (1 - (kw >> 1))) * 16;
            d0 = tmem_acc + (((kh & 1) << 1) | (kw & 1)) * P.n_tile;
            fresh = kb == 0 && kh < 2 && kw < 2;
          } else if (DN) {   // out[i, j] += x[2i + kh, 2j + kw] * W[kh,kw]: parity plane (kh&1, kw&1), shift (kh>>1, kw>>1)
            a0 = A_addr + ((((kh & 1) << 1) | (kw & 1)) * (kCinBlk / 8)) * chunk_bytes + ((kh >> 1) * hw + (kw >> 1)) * 16;
            d0 = tmem_acc, fresh = step == 0;
          } else {
            a0 = A_addr + (kh * hw + kw) * 16, d0 = tmem_acc, fresh = step == 0;
          }
#pragma unroll
          for (int j = 0; j < kCinBlk / 16; ++j)
            umma_ss(d0, smem_desc(a0 + 2 * j * chunk_bytes, chunk_bytes, hw * 16),
                    smem_desc(b0 + 2 * j * P.n_tile * 16, P.n_tile * 16, 128), idesc, !(fresh && j == 0));
          umma_commit(bar_bfree + slot * 8);
          // refill the slot of the PREVIOUS step (drained, or about to be) with the slice kBStages steps after it
          const int prev = step - 1, nxt = prev + kBStages;
          if (prev >= 0) {
            if (nxt < total_steps) {
              mbar_wait_spin(bar_bfree + ps * 8, pphase);
              mbar_expect_tx(bar_bfull + ps * 8, b_bytes);
              bulk_g2s(smem_base + kSmB + ps * b_bytes, wsrc + (size_t)nxt * b_bytes, b_bytes, bar_bfull + ps * 8);
            }
            if (++ps == kBStages) ps = 0, pphase ^= 1;
          }
          if (++slot == kBStages) slot = 0, sphase ^= 1;
        }
        umma_commit(bar_afree + (kb & 1) * 8);
      }
      if (my_kblocks > 0) umma_commit(bar_acc);
      HAV_CONV_STAMP(5, true);
    }
    __syncwarp();
  } else {
    // ================= staging warps =================
    const float *xb = P.x + (size_t)b * P.Cin * P.H * P.W;
    const uint16_t *xcl = reinterpret_cast<const uint16_t *>(P.x);
    const size_t cstride = (size_t)P.H * P.W;
    for (int kl = 0; kl < my_kblocks; ++kl) {
      const int kb = kb_begin + kl;        // channel block of the layer; buffers and barrier phases follow the local count kl
      uint8_t *A = smem + kSmA + (kl & 1) * a_slot;
      float *sc = sscale + (kl & 1) * kCinBlk;
      if (kl >= 2) mbar_wait_spin(bar_afree + (kl & 1) * 8, ((kl - 2) >> 1) & 1);   // MMAs of block kl-2 are done with this buffer
      if (INCL) {
        if (tid < kCinBlk / 2 && P.in_scale != nullptr) {   // packed fp16 pairs (c, c+1)
          const int c = kb * kCinBlk + 2 * tid;
          const float s0 = c < P.Cin ? __ldg(P.in_scale + (size_t)b * P.Cin + c) : 0.0f;
          const float s1 = c + 1 < P.Cin ? __ldg(P.in_scale + (size_t)b * P.Cin + c + 1) : 0.0f;
          reinterpret_cast<uint32_t *>(sc)[tid] = pack2<false>(s0, s1);
        }
      } else if (tid < kCinBlk) {
        const int c = kb * kCinBlk + tid;
        sc[tid] = (c < P.Cin) ? (P.in_scale != nullptr ? __ldg(P.in_scale + (size_t)b * P.Cin + c) : 1.0f) : 0.0f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kStageThreads) : "memory");
      HAV_CONV_STAMP(2, tid == 0 && kl == 0);
      // unit = (8-channel chunk, halo pixel) = 16 bytes of the A operand; consecutive threads take consecutive halo pixels of one
      // chunk.  The loads of kGroup units per thread are issued back to back before the first one is consumed: a thread's
      // staging time is a few load latencies per channel block, not one per unit.
      constexpr int n_units = n_chunks * halo_px, n_iters = (n_units + kStageThreads - 1) / kStageThreads;
      constexpr int kGroup = INCL ? 7 : 3;
#pragma unroll 1
      for (int i0 = 0; i0 < n_iters; i0 += kGroup) {
        uint4 u[kGroup];
        float v[INCL ? 1 : kGroup][8];
        int dst[kGroup], cqs[kGroup];
#pragma unroll
        for (int g = 0; g < kGroup; ++g) {
          const int unit = tid + (i0 + g) * kStageThreads;
          const int chunk = unit / halo_px, hp = unit - chunk * halo_px;
          const int py = hp / hw, px = hp - py * hw;
          const int Y = DN ? 2 * (vy0 + py) + (chunk >> 4) : vy0 + py - P.pad;
          const int X = DN ? 2 * (vx0 + px) + ((chunk >> 3) & 1) : vx0 + px - P.pad;
          const bool live = i0 + g < n_iters && unit < n_units;
          const bool ok = live && Y >= 0 && X >= 0 && Y < P.H && X < P.W;
          const int cq = DN ? (chunk & 7) : chunk;      // channel chunk of the staged 64-channel block
          const int c0 = kb * kCinBlk + cq * 8;
          dst[g] = live ? chunk * chunk_bytes + hp * 16 : -1, cqs[g] = cq;
          u[g] = make_uint4(0u, 0u, 0u, 0u);
          if (INCL) {
            // channels-last fp16: the 8 channels of this unit are one aligned 16-byte load
            if (ok && c0 < P.Cin) u[g] = __ldg(reinterpret_cast<const uint4 *>(xcl + (((size_t)b * P.H + Y) * P.W + X) * P.Cin + c0));
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[INCL ? 0 : g][e] = (ok && c0 + e < P.Cin) ? __ldg(xb + (size_t)(c0 + e) * cstride + (size_t)Y * P.W + X) : 0.0f;
          }
        }
#pragma unroll
        for (int g = 0; g < kGroup; ++g) {
          if (dst[g] < 0) continue;
          if (INCL) {
            if (P.in_scale != nullptr) {     // modulate as packed pairs
              const uint32_t *sh = reinterpret_cast<const uint32_t *>(sc) + cqs[g] * 4;   // 4 packed pairs of this chunk
              u[g].x = mul2<false>(u[g].x, sh[0]), u[g].y = mul2<false>(u[g].y, sh[1]);
              u[g].z = mul2<false>(u[g].z, sh[2]), u[g].w = mul2<false>(u[g].w, sh[3]);
            }
          } else {
            float (&w)[8] = v[INCL ? 0 : g];
#pragma unroll
            for (int e = 0; e < 8; ++e) w[e] *= sc[cqs[g] * 8 + e];
            u[g] = make_uint4(pack2<kBF16>(w[0], w[1]), pack2<kBF16>(w[2], w[3]), pack2<kBF16>(w[4], w[5]), pack2<kBF16>(w[6], w[7]));
          }
          *reinterpret_cast<uint4 *>(A + dst[g]) = u[g];
        }
      }
      fence_async_smem();
      mbar_arrive_conv(bar_afull + (kl & 1) * 8);
      HAV_CONV_STAMP(3, tid == 0 && kl == 0);
      HAV_CONV_STAMP(6, tid == 0 && kl == my_kblocks - 1);
    }
    // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., column half w/4.  Row m = tile position (vy0 + m/8, vx0 + m%8);
    //      polyphase: accumulator block ph = 2a+b holds output (2i+a, 2j+b)
    mbar_wait_spin(bar_tmem, 0);
    tc_fence_after();
    tmem_acc = *tmem_slot;
    if (my_kblocks > 0) {
      mbar_wait_spin(bar_acc, 0);
      tc_fence_after();
    }
    HAV_CONV_STAMP(7, tid == 0);
    const int wq = warp_u & 3, half = warp_u >> 2;
    const int m = wq * 32 + (tid & 31), vy = vy0 + (m >> 3), vx = vx0 + (m & 7);
    const uint32_t trow = tmem_acc + ((uint32_t)(wq * 32) << 16);
    // output pixel of row m in accumulator block ph, its validity and its noise term
    auto out_pixel = [&](int ph, int &oy, int &ox, float &nz) -> bool {
      bool ok;
      if (UPP) ok = true, oy = 2 * vy + (ph >> 1), ox = 2 * vx + (ph & 1);
      else if (!DN && P.down == 2) ok = !(vy & 1) && !(vx & 1), oy = vy >> 1, ox = vx >> 1;   // small layers, 1x1 stride 2
      else ok = true, oy = vy, ox = vx;
      ok = ok && oy < P.Ho && ox < P.Wo;
      nz = 0.0f;
      if (ok && P.noise != nullptr) nz = P.noise_weight * __ldg(P.noise + (size_t)b * P.noise_bstride + (size_t)oy * P.Wo + ox);
      return ok;
    };
    if (P.ksplit == 1) {
      const int cols_half = ((P.n_tile / 16 + 1) / 2) * 16;
      const int c_beg = half * cols_half, c_end = min(P.n_tile, c_beg + cols_half);
#pragma unroll 1
      for (int ph = 0; ph < (UPP ? 4 : 1); ++ph) {
        int oy, ox;
        float nz;
        const bool ok = out_pixel(ph, oy, ox, nz);
        for (int c0 = c_beg; c0 < c_end; c0 += 16) {
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(trow + ph * P.n_tile + c0));
          tmem_wait_ld();
          if (ok) {
            float v[8];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[hf * 8 + j]);
              conv_tail_store8<OUTCL>(P, epi, c0 + hf * 8, b, oy, ox, nt * P.n_tile + c0 + hf * 8, nz, v);
            }
          }
        }
      }
    } else {
      // ---- split-K: partial tile -> own shared memory (column-major [column][128 rows] fp32 over the drained operand
      //      buffers), cluster barrier, then this CTA reduces its slice of the columns over all ranks
      const int n_cols = (UPP ? 4 : 1) * P.n_tile;
      float *dump = reinterpret_cast<float *>(smem + P.ring_bytes + kOffDump);   // behind the barriers and the epilogue tables
      if (my_kblocks > 0) {
        const int cols_half = ((n_cols / 16 + 1) / 2) * 16;
        const int c_beg = half * cols_half, c_end = min(n_cols, c_beg + cols_half);
        for (int c0 = c_beg; c0 < c_end; c0 += 16) {
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(trow + c0));
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) dump[(c0 + j) * 128 + m] = __uint_as_float(r[j]);
        }
      }
    }
  }
  HAV_CONV_STAMP(8, tid == 0);
  if (P.ksplit > 1) {
    cluster_sync_all();          // every CTA of the cluster (all warps, the control warp included) has written its partial tile
    if (warp_u < kStageThreads / 32) {
      const int n_cols = (UPP ? 4 : 1) * P.n_tile, slice = n_cols / P.ksplit;       // slice % 8 == 0 (host)
      const int m = tid & 127, vy = vy0 + (m >> 3), vx = vx0 + (m & 7);
      const int live_ranks = (P.kblocks + kb_per - 1) / kb_per;                      // ranks beyond it own no channel block
      for (int g = tid >> 7; g < slice / 8; g += kStageThreads / 128) {
        const int c = (int)blockIdx.z * slice + g * 8, ph = c / P.n_tile, cl = c - ph * P.n_tile;
        float v[8], part[8][8];      // all remote loads of the group are in flight together; summed in rank order
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            part[q][j] = q < live_ranks ? ld_dsmem_f32(smem_base + P.ring_bytes + kOffDump + (uint32_t)((c + j) * 128 + m) * 4u, (uint32_t)q) : 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = part[0][j];
#pragma unroll
          for (int q = 1; q < 8; ++q) v[j] += part[q][j];
        }
        bool ok;
        int oy, ox;
        if (UPP) ok = true, oy = 2 * vy + (ph >> 1), ox = 2 * vx + (ph & 1);
        else if (!DN && P.down == 2) ok = !(vy & 1) && !(vx & 1), oy = vy >> 1, ox = vx >> 1;
        else ok = true, oy = vy, ox = vx;
        ok = ok && oy < P.Ho && ox < P.Wo;
        if (ok) {
          float nz = 0.0f;
          if (P.noise != nullptr) nz = P.noise_weight * __ldg(P.noise + (size_t)b * P.noise_bstride + (size_t)oy * P.Wo + ox);
          conv_tail_store8<OUTCL>(P, epi, cl, b, oy, ox, nt * P.n_tile + cl, nz, v);
        }
      }
    }
    cluster_sync_all();          // no CTA leaves while its partial tile is still being read
  }
  HAV_CONV_STAMP(9, tid == 0);
  tc_fence_before();
  __syncthreads();
  if (warp_u == kStageThreads / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(P.tmem_cols));
  HAV_CONV_STAMP(10, tid == 0);
}

}  // namespace conv
}  // namespace hav

using namespace hav;

// output channels per CTA: 128 for the plain convolution, 64 for the polyphase transposed convolution (4 accumulators)
static int conv_n_tile(int cout, int up) {
  const int cap = up == 2 ? 64 : conv::kNTileMax;
  int n = (cout + 15) / 16 * 16;
  return n < cap ? n : cap;
}

extern "C" uint64_t hav_conv_wpack_bytes(int cout, int cin, int ksize, int up) {
  if (cout < 1 || cin < 1 || (ksize != 1 && ksize != 3) || (up != 1 && up != 2)) return 0;
  const int n_tile = conv_n_tile(cout, up), n_tiles = (cout + n_tile - 1) / n_tile, kblocks = (cin + conv::kCinBlk - 1) / conv::kCinBlk;
  return (uint64_t)n_tiles * kblocks * ksize * ksize * n_tile * conv::kCinBlk * 2;
}

extern "C" int hav_conv_pack_weights(void *wpack, const float *w, int cout, int cin, int ksize, float scale, int up,
                                     int transpose_io, int precision, void *stream) {
  if (wpack == nullptr || w == nullptr) return HAV_E_NULL;
  if (cout < 1 || cin < 1 || (ksize != 1 && ksize != 3) || (up != 1 && up != 2)) return HAV_E_SHAPE;
  if (precision != HAV_PREC_FP16 && precision != HAV_PREC_BF16) return HAV_E_VALUE;
  const int flip = (transpose_io >> 1) & 1;
  transpose_io &= 1;
  const int n_tile = conv_n_tile(cout, up), n_tiles = (cout + n_tile - 1) / n_tile, kblocks = (cin + conv::kCinBlk - 1) / conv::kCinBlk;
  const long units = (long)n_tiles * kblocks * (conv::kCinBlk / 8) * n_tile;
  const int grid = (int)((units + 127) / 128 < 148 * 16 ? (units + 127) / 128 : 148 * 16);
  auto go = [&](auto kern) {
    kern<<<grid, 128, 0, (cudaStream_t)stream>>>(w, (uint16_t *)wpack, cout, cin, n_tile, n_tiles, kblocks, scale, flip, transpose_io);
  };
  const bool bf = precision == HAV_PREC_BF16;
  if (ksize == 3) bf ? go(conv::pack_conv_weights_kernel<true, 9>) : go(conv::pack_conv_weights_kernel<false, 9>);
  else bf ? go(conv::pack_conv_weights_kernel<true, 1>) : go(conv::pack_conv_weights_kernel<false, 1>);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_modconv_demod(float *demod, const float *w, const float *style, int batch, int cout, int cin, int ksize,
                                 float scale, float eps, void *stream) {
  if (demod == nullptr || w == nullptr || style == nullptr) return HAV_E_NULL;
  if (batch < 1 || cout < 1 || cin < 1 || ksize < 1 || batch > 65535) return HAV_E_SHAPE;
  conv::modconv_demod_kernel<<<dim3(cout, batch), 128, 0, (cudaStream_t)stream>>>(w, style, demod, cout, cin, ksize * ksize, scale, eps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

#ifdef HAV_CONV_TIMING
extern "C" int hav_conv_debug_stamps(unsigned long long *out16) {
  return (int)cudaMemcpyFromSymbol(out16, conv::g_conv_t, sizeof(unsigned long long) * 16);
}
#endif

extern "C" int hav_conv2d_forward(const hav_conv_args *a, void *stream) {
  if (a == nullptr) return HAV_E_NULL;
  if (a->struct_bytes != sizeof(hav_conv_args)) return HAV_E_VALUE;
  if (a->batch < 0 || a->cin < 1 || a->cout < 1 || a->in_h < 1 || a->in_w < 1) return HAV_E_SHAPE;
  if ((a->ksize != 1 && a->ksize != 3) || (a->up != 1 && a->up != 2) || (a->down != 1 && a->down != 2)) return HAV_E_VALUE;
  if (a->up == 2 && (a->down == 2 || a->ksize != 3)) return HAV_E_VALUE;
  if (a->precision != HAV_PREC_FP16 && a->precision != HAV_PREC_BF16) return HAV_E_VALUE;
  if ((a->in_layout | a->out_layout) & ~1) return HAV_E_VALUE;
  if ((a->in_layout || a->out_layout) && a->precision != HAV_PREC_FP16) return HAV_E_VALUE;
  if (a->in_layout && (a->cin % 8) != 0) return HAV_E_SHAPE;
  if (a->out_layout && (a->cout % 8) != 0) return HAV_E_SHAPE;
  if (a->batch == 0) return HAV_OK;
  if (a->x == nullptr || a->wpack == nullptr || a->out == nullptr) return HAV_E_NULL;
  conv::ConvDev P;
  memset(&P, 0, sizeof(P));
  P.B = a->batch, P.Cin = a->cin, P.Cout = a->cout, P.H = a->in_h, P.W = a->in_w, P.k = a->ksize, P.up = a->up, P.down = a->down;
  P.act = a->act;
  int vh, vw;   // tile grid: stride-1 output positions (output positions for the polyphase stride-2 form)
  bool dn = false;
  if (a->up == 2) {
    P.pad = 1, vh = a->in_h + 1, vw = a->in_w + 1;      // polyphase tiles run over input positions 0..H (conv_transpose2d, stride 2, pad 0)
    P.Ho = 2 * a->in_h + 1, P.Wo = 2 * a->in_w + 1;
  } else if (a->down == 2) {
    P.pad = 0, vh = a->in_h - a->ksize + 1, vw = a->in_w - a->ksize + 1;                                // stride 2, pad 0
    if (vh < 1 || vw < 1) return HAV_E_SHAPE;
    P.Ho = (vh + 1) / 2, P.Wo = (vw + 1) / 2;
    // polyphase (tiles over the OUTPUT positions, 4 staged parity planes, one CTA per SM) once the grid fills the chip twice
    // over; small layers keep the light-weight form (stride-1 tiles, even positions stored, two CTAs per SM): measured --
    // the polyphase form is 7-27 % faster on the large layers and slower on the 8x8 ... 32x32 ones
    if (a->ksize == 3) {
      const long out_tiles = (long)a->batch * ((P.Wo + conv::kTileW - 1) / conv::kTileW) * ((P.Ho + conv::kTileH - 1) / conv::kTileH);
      const int nt = (a->cout + conv_n_tile(a->cout, 1) - 1) / conv_n_tile(a->cout, 1);
      dn = out_tiles * nt >= 2 * 148;
      if (dn) vh = P.Ho, vw = P.Wo;
    }
  } else {
    P.pad = a->ksize / 2, vh = a->in_h, vw = a->in_w, P.Ho = vh, P.Wo = vw;
  }
  P.n_tile = conv_n_tile(a->cout, a->up), P.n_tiles = (a->cout + P.n_tile - 1) / P.n_tile;
  P.kblocks = (a->cin + conv::kCinBlk - 1) / conv::kCinBlk;
  {
    const int need = (a->up == 2 ? 4 : 1) * P.n_tile;
    P.tmem_cols = 32;
    while (P.tmem_cols < need) P.tmem_cols *= 2;
  }
  P.tiles_x = (vw + conv::kTileW - 1) / conv::kTileW, P.tiles_y = (vh + conv::kTileH - 1) / conv::kTileH;
  P.x = (const float *)a->x, P.in_scale = a->in_scale, P.out_scale = a->out_scale, P.noise = a->noise, P.bias = a->bias;
  P.wpack = (const uint8_t *)a->wpack, P.out = (float *)a->out, P.noise_weight = a->noise_weight;
  P.residual = (const float *)a->residual;
  P.noise_bstride = a->noise_per_sample ? P.Ho * P.Wo : 0;
  const long sp_tiles = (long)a->batch * P.tiles_x * P.tiles_y;
  if (sp_tiles > 2147483647L || P.n_tiles > 65535) return HAV_E_SHAPE;
  // split-K over a cluster when the layer has many channel blocks and too few output tiles to occupy the chip
  const long ctas = sp_tiles * P.n_tiles;
  const int n_cols = (a->up == 2 ? 4 : 1) * P.n_tile;
  int ks_first = 1;
  {
    static const bool no_split = getenv("HAV_CONV_NO_SPLITK") != nullptr;      // A/B aids
    static const int max_split = getenv("HAV_CONV_MAX_SPLIT") ? atoi(getenv("HAV_CONV_MAX_SPLIT")) : 8;
    if (!no_split && !dn && P.kblocks >= 2 && ctas * 2 <= 148) {
      int ks = max_split < 8 ? (max_split < 1 ? 1 : max_split) : 8;
      while (ks > 1 && (ks > P.kblocks || ctas * ks > 148 || n_cols % (8 * ks) != 0)) ks >>= 1;
      ks_first = ks;
    }
  }
  int smem_bytes = 0;
  // ring geometry and shared-memory size for a given split
  auto geometry = [&](int ks) -> bool {
    P.ksplit = ks;
    // weight ring: 64 KB keeps two CTAs per SM; a launch of at most one CTA per SM takes 128 KB
    const long total_ctas = ctas * ks;
    const int slot = P.n_tile * conv::kCinBlk * 2;
    static const int deep_max = getenv("HAV_CONV_DEEP_MAX_CTAS") ? atoi(getenv("HAV_CONV_DEEP_MAX_CTAS")) : 148;   // A/B aid
    int budget = (!dn && total_ctas <= deep_max) ? conv::kBRingBytesDeep : conv::kBRingBytes;
    if (ks > 1) {     // the partial-tile dump shares the 227 KB with the ring
      const int room = 227 * 1024 - conv::kOffDump - n_cols * 128 * 4;
      if (room < budget) budget = room;
    }
    P.b_stages = budget / slot < conv::kBStagesMax ? budget / slot : conv::kBStagesMax;
    if (P.b_stages < 2) return false;
    P.ring_bytes = P.b_stages * slot;
    smem_bytes = P.ring_bytes + (dn ? conv::kSmemTailDN : conv::kSmemTail);
    if (ks > 1) {
      const int need = P.ring_bytes + conv::kOffDump + n_cols * 128 * 4;    // partial-tile dump
      if (need > smem_bytes) smem_bytes = need;
    }
    return true;
  };
  cudaError_t e;
  auto launch = [&](auto kern) -> cudaError_t {
    // every cluster of a split launch must be resident at once (a second wave of clusters would double the layer's time):
    // halve the split until the device can hold them (clusters are placed inside one GPC, so this is less than SMs / size)
    int ks = ks_first;
    for (;; ks >>= 1) {
      if (!geometry(ks)) return cudaErrorInvalidValue;
      cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
      if (er != cudaSuccess) return er;
      if (ks == 1) break;
      static std::mutex mu;
      static std::map<std::tuple<const void *, int, int, int>, int> cache;      // (kernel, device, split, smem) -> resident clusters
      int dev = 0;
      cudaGetDevice(&dev);
      const auto key = std::make_tuple((const void *)kern, dev, ks, smem_bytes);
      int resident = -1;
      {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) resident = it->second;
      }
      if (resident < 0) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(1, 1, ks), q.blockDim = dim3(conv::kThreadsV2), q.dynamicSmemBytes = smem_bytes;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = 1, qa[0].val.clusterDim.y = 1, qa[0].val.clusterDim.z = ks;
        q.attrs = qa, q.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&resident, kern, &q) != cudaSuccess) resident = 0, (void)cudaGetLastError();
        std::lock_guard<std::mutex> lk(mu);
        cache[key] = resident;
      }
      if (ctas <= resident) break;
    }
    dim3 grid((unsigned)sp_tiles, P.n_tiles, P.ksplit);
    if (P.ksplit == 1) {
      kern<<<grid, conv::kThreadsV2, smem_bytes, (cudaStream_t)stream>>>(P);
      return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = dim3(conv::kThreadsV2), cfg.dynamicSmemBytes = smem_bytes, cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = P.ksplit;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, P);
  };
  const bool bf = a->precision == HAV_PREC_BF16;
  const int lay = (a->in_layout ? 2 : 0) | (a->out_layout ? 1 : 0);
  if (bf) {   // bf16 operands: NCHW fp32 tensors only
    if (a->up == 2) e = launch(conv::conv_tc_kernel<true, 3, true, false, false>);
    else if (dn) e = launch(conv::conv_tc_kernel<true, 3, false, false, false, true>);
    else if (a->ksize == 3) e = launch(conv::conv_tc_kernel<true, 3, false, false, false>);
    else e = launch(conv::conv_tc_kernel<true, 1, false, false, false>);
  } else {
#define HAV_CONV_DISPATCH(KS_, UPP_, DN_)                                                        \
  switch (lay) {                                                                                 \
    case 0: e = launch(conv::conv_tc_kernel<false, KS_, UPP_, false, false, DN_>); break;        \
    case 1: e = launch(conv::conv_tc_kernel<false, KS_, UPP_, false, true, DN_>); break;         \
    case 2: e = launch(conv::conv_tc_kernel<false, KS_, UPP_, true, false, DN_>); break;         \
    default: e = launch(conv::conv_tc_kernel<false, KS_, UPP_, true, true, DN_>); break;         \
  }
    if (a->up == 2) { HAV_CONV_DISPATCH(3, true, false) }
    else if (dn) { HAV_CONV_DISPATCH(3, false, true) }
    else if (a->ksize == 3) { HAV_CONV_DISPATCH(3, false, false) }
    else { HAV_CONV_DISPATCH(1, false, false) }
#undef HAV_CONV_DISPATCH
  }
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}
