// Weight gradient of the StyleUNet convolutions (the backward of conv_tc.cu) on tcgen05, plus the small row kernel that
// finishes the data gradient of a modulated convolution.
//
// Reference: autograd through ModulatedConv2d.forward / EqualConv2d.forward (model/styleUnet.py:222-297, :108-118), which the
// reference routes through cuDNN's convolution_backward (model/op/conv2d_gradfix.py:94-227).
//
// With the shared-weight formulation of conv_tc.cu,  y[b,co] = d[b,co] * sum_{ci,kh,kw} (scale W[co,ci,kh,kw]) (s[b,ci] x[b,ci,..]),
//     dW[co,ci,kh,kw] = scale * sum_{b,p} (d[b,co] g[b,co,p]) * (s[b,ci] x[b,ci,p + (kh,kw) - pad])
// is ONE GEMM per tap with M = Cout, N = Cin and K = every output position of the batch.  Both operands are read where they
// lie in their NCHW fp32 tensors: K (positions) is the contiguous axis, so
//   A = d*g   tile [128 co][8 rows x 16 px]  is staged K-major   ([px/8][co][8 px],  16 B = 8 consecutive pixels of a channel),
//   B = s*x   halo [48 ci][10 x 18 px]       is staged MN-major  ([ci/8][halo px][8 ci], 16 B = 8 channels of a pixel) --
// the layout conv_tc.cu stages its input in -- so the (kh,kw) tap is the SAME buffer seen through a descriptor whose start
// is shifted by (kh*18 + kw) pixels: one K = 16 step = one 16-pixel row of the tile, 9 taps = 9 accumulators of 48 columns in
// TMEM (432 of 512 columns), no im2col, no transposed copies.  A CTA owns a (128 co) x (48 ci) x 9 block of dW and a strided
// share of the position tiles (split-K); partial sums leave TMEM as vectorised fp32 reductions (red.global.add.v4.f32).
//
// The strided convolutions are polyphase: with q the tensor that is read with stride 2,
//   down = 2 (stride 2, pad 0):  dW[kh,kw] = sum_p g[p] x[2p + (kh,kw)]     A = g, B = x
//   up   = 2 (conv_transpose2d): dW[kh,kw] = sum_p x[p] g[2p + (kh,kw)]     A = x, B = g  (accumulators hold dW transposed)
// the B side is staged as its four parity planes q[2i+a, 2j+b] ((8+1) x (16+1) pixels each) and tap (kh,kw) reads plane
// (kh&1, kw&1) shifted by (kh>>1, kw>>1) -- the same 72 MMAs per tile as a stride-1 layer, no multiplications by zeros.
#include "tc_common.cuh"

namespace hav {
namespace wg {

using namespace tc;

constexpr int kTH = 8, kTW = 16;                   // position tile: 8 rows x 16 px = 8 K-steps of 16
constexpr int kNb = 48, kMa = 128;                 // B-side channels (TMEM columns per tap) and A-side channels (TMEM lanes) per CTA
constexpr int kAChunk = kMa * 16 + 16;              // K-chunk stride of the A tile, +16 B so row-major stores spread over the banks
constexpr int kABytes = (kTH * kTW / 8) * kAChunk;    // 33024
constexpr int kStageThreads = 512, kThreads = kStageThreads + 32;   // 16 staging warps: the loads are latency-bound, not issue-bound
constexpr int kTmemCols = 512;
constexpr int kTmemA = 9 * kNb;                    // columns [432, 496): the A tile, 8 K-steps x 8 columns (two bf16 per column)
constexpr uint32_t kBMajorMN = 1u << 16;
constexpr int kSmScale = 128, kSmA = 1024;         // [0,128): mbarriers + TMEM slot; [128,1024): B-side scale tables

// B-side staging geometry.  STRIDE 1: one halo of (8+2) x (16+2) pixels, tap (kh,kw) = shift (kh,kw).
// STRIDE 2: four parity planes q[2i+a, 2j+b] of (8+1) x (16+1) pixels, tap (kh,kw) = plane (kh&1, kw&1) shifted by (kh>>1, kw>>1).
template <int STRIDE>
struct Geo {
  static constexpr int kW = STRIDE == 1 ? kTW + 2 : kTW + 1, kH = STRIDE == 1 ? kTH + 2 : kTH + 1;
  static constexpr int kPx = kW * kH;                          // 180 / 153
  static constexpr int kPlanes = STRIDE == 1 ? 1 : 4;
  static constexpr int kChunk = kPx * 16;                      // bytes of one 8-channel chunk of one plane
  static constexpr int kPlaneBytes = (kNb / 8) * kChunk;
  static constexpr int kBBytes = kPlanes * kPlaneBytes;        // 17280 / 58752
  static constexpr int kUnits = kPlanes * (kNb / 8) * kPx;     // 1080 / 3672
  static constexpr int kStages = STRIDE == 1 ? 3 : 2;
  static constexpr int kSmB = kSmA + kStages * kABytes;
  static constexpr int kSmemBytes = kSmB + kStages * kBBytes;  // 151168 / 184064
};

struct WgDev {
  int B, Ca, Cb;                      // channels of the A tensor (TMEM lanes) and of the B tensor (TMEM columns)
  int a_h, a_w, b_h, b_w;             // spatial sizes; position tiles run over the A tensor
  int pad;                            // stride 1: B coordinate = position + tap - pad;  stride 2: 2 * position + tap
  int out_t;                          // 0: dw[(a_ch * Cb + b_ch) * taps + tap]   1: dw[(b_ch * Ca + a_ch) * taps + tap]
  int dbg;                            // HAV_WG_DEBUG: 1 = skip operand staging after the first ring fill (timing experiments only)
  int a_vec, dw_vec;                  // 16-byte vector loads of the A tensor / vector reductions into dw are aligned
  int tiles_x, tiles_y, ntiles, nsplit, a_tiles, b_tiles;
  const float *pa, *pb, *sa, *sb;     // tensors [B,C,h,w] and their per-(sample, channel) scales [B,C] (NULL = 1)
  float *dw;
  float wscale;
};

__device__ __forceinline__ void mbar_arrive_wg(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float *p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// 128 rows x 256 bits (one K = 16 step of a 16-bit A operand) from a K-major shared-memory matrix into TMEM lanes 0..127,
// 8 columns: the layout tcgen05.mma reads an A operand from when it is given a TMEM address
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

#define HAV_TMEM_LD8(r, taddr)                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                 \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(taddr))

template <int KS, int STRIDE>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_kernel(const WgDev P) {
  using G = Geo<STRIDE>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_base, bar_free = bar_full + G::kStages * 8, bar_acc = bar_free + G::kStages * 8;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + 120);
  constexpr int taps = KS * KS;

  const int at = blockIdx.x % P.a_tiles, bt = blockIdx.x / P.a_tiles;
  const int split = blockIdx.y;
  const int ca0 = at * kMa, cb0 = bt * kNb;

  if (warp_u == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + 120), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int i = 0; i < G::kStages; ++i) mbar_init(bar_full + i * 8, kStageThreads), mbar_init(bar_free + i * 8, 1);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp_u == kStageThreads / 32) {
    // ================= control warp: 8 rows x taps MMAs per position tile =================
    if (elect_one()) {
      constexpr uint32_t idesc = instr_desc(kNb, true) | kBMajorMN;
      int it = 0;
      for (int t = split; t < P.ntiles; t += P.nsplit, ++it) {
        const int st = it % G::kStages;
        mbar_wait_spin(bar_full + st * 8, (it / G::kStages) & 1);
        tc_fence_after();
        const uint32_t A0 = smem_base + kSmA + st * kABytes, B0 = smem_base + G::kSmB + st * G::kBBytes;
        // the A tile goes to TMEM once (8 copies of 128 x 32 B): each of its K-steps is then read by `taps` MMAs without
        // touching shared memory again (an SS-mode M128 x N48 MMA is bound by its 4 KB A-operand fetch, not by the math).
        // tcgen05.cp and tcgen05.mma of one thread execute in issue order, so the copies of tile t+1 cannot overtake the
        // MMAs of tile t that still read these columns.
#pragma unroll
        for (int r = 0; r < kTH; ++r) tmem_cp_128x256b(tmem_acc + kTmemA + r * 8, smem_desc(A0 + 2 * r * kAChunk, kAChunk, 128));
#pragma unroll 1
        for (int r = 0; r < kTH; ++r) {
#pragma unroll
          for (int tap = 0; tap < taps; ++tap) {
            const int kh = tap / KS, kw = tap - kh * KS;
            uint32_t b_addr;
            if (STRIDE == 1) b_addr = B0 + ((r + kh) * G::kW + kw) * 16;
            else b_addr = B0 + (((kh & 1) << 1) | (kw & 1)) * G::kPlaneBytes + ((r + (kh >> 1)) * G::kW + (kw >> 1)) * 16;
            // MN-major: LBO = distance between 8-pixel K groups (128 B), SBO = distance between 8-channel groups (one chunk)
            umma_ts(tmem_acc + tap * kNb, tmem_acc + kTmemA + r * 8, smem_desc(b_addr, 128, G::kChunk), idesc, (it > 0 || r > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_free + st * 8);
      }
      umma_commit(bar_acc);
    }
    __syncwarp();
  } else {
    // ================= staging warps =================
    const size_t a_plane = (size_t)P.a_h * P.a_w, b_plane = (size_t)P.b_h * P.b_w;
    // A tile: one warp instruction = one channel x 8 rows x 64 B (lane = (row, float4 of the row)): 8 cache lines per 512 B
    const int a_f4 = tid & 3, a_row = (tid >> 2) & 7, a_ch0 = (tid >> 5) * (kMa / 16);   // 16 warps x 8 channels
    const bool a_vec = P.a_vec != 0;
    int it = 0;
    for (int t = split; t < P.ntiles; t += P.nsplit, ++it) {
      const int st = it % G::kStages;
      int sp = t;
      const int tx = sp % P.tiles_x; sp /= P.tiles_x;
      const int ty = sp % P.tiles_y;
      const int b = sp / P.tiles_y;
      const int Y0 = ty * kTH, X0 = tx * kTW;
      if (P.dbg == 1 && it >= G::kStages) {   // experiment: the MMA side alone (operands of the first ring fill are reused)
        mbar_wait_spin(bar_free + st * 8, ((it / G::kStages) - 1) & 1);
        mbar_arrive_wg(bar_full + st * 8);
        continue;
      }
      // ---- A loads first (they do not depend on the ring slot): 8 channels of this warp, this lane's row and 4-px quarter
      float4 va[kMa / 16];
      float as[kMa / 16];
      {
        const int Y = Y0 + a_row, X = X0 + a_f4 * 4;
#pragma unroll
        for (int j = 0; j < kMa / 16; ++j) {
          const int ca = ca0 + a_ch0 + j;
          const bool ok = ca < P.Ca && Y < P.a_h;
          const float *src = P.pa + (((size_t)b * P.Ca + (ok ? ca : 0)) * P.a_h + (ok ? Y : 0)) * P.a_w + X;
          if (a_vec && ok && X + 4 <= P.a_w) {
            va[j] = __ldg(reinterpret_cast<const float4 *>(src));
          } else {
            va[j].x = (ok && X + 0 < P.a_w) ? __ldg(src + 0) : 0.0f;
            va[j].y = (ok && X + 1 < P.a_w) ? __ldg(src + 1) : 0.0f;
            va[j].z = (ok && X + 2 < P.a_w) ? __ldg(src + 2) : 0.0f;
            va[j].w = (ok && X + 3 < P.a_w) ? __ldg(src + 3) : 0.0f;
          }
          as[j] = (ca < P.Ca && P.sa != nullptr) ? __ldg(P.sa + (size_t)b * P.Ca + ca) : 1.0f;
        }
      }
      if (it >= G::kStages) mbar_wait_spin(bar_free + st * 8, ((it / G::kStages) - 1) & 1);
      float *sc = reinterpret_cast<float *>(smem + kSmScale) + st * kNb;
      if (tid < kNb) {
        const int cb = cb0 + tid;
        sc[tid] = cb < P.Cb ? (P.sb != nullptr ? __ldg(P.sb + (size_t)b * P.Cb + cb) : 1.0f) : 0.0f;
      }
      {
        uint8_t *A = smem + kSmA + st * kABytes + (2 * a_row + (a_f4 >> 1)) * kAChunk + (a_f4 & 1) * 8;
#pragma unroll
        for (int j = 0; j < kMa / 16; ++j)
          *reinterpret_cast<uint2 *>(A + (a_ch0 + j) * 16) =
              make_uint2(pack2<true>(va[j].x * as[j], va[j].y * as[j]), pack2<true>(va[j].z * as[j], va[j].w * as[j]));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kStageThreads) : "memory");   // the scale table of this stage is complete
      // ---- B: MN-major planes.  unit = (plane, 8-channel chunk, pixel); consecutive threads take consecutive pixels.
      //      Batches of 3-4 units: 24-32 independent loads in flight per thread before the first conversion.
      {
        uint8_t *Bm = smem + G::kSmB + st * G::kBBytes;
        const float *bb = P.pb + (size_t)b * P.Cb * b_plane;
        constexpr int kBatch = STRIDE == 1 ? 3 : 4, kIters = (G::kUnits + kStageThreads * kBatch - 1) / (kStageThreads * kBatch);
#pragma unroll 1
        for (int bi = 0; bi < kIters; ++bi) {
          float vb[kBatch][8];
          int dst[kBatch], chs[kBatch];
#pragma unroll
          for (int j = 0; j < kBatch; ++j) {
            const int u = tid + (bi * kBatch + j) * kStageThreads;
            const bool in = u < G::kUnits;
            const int plane = STRIDE == 1 ? 0 : u / ((kNb / 8) * G::kPx);
            const int rem = STRIDE == 1 ? u : u - plane * ((kNb / 8) * G::kPx);
            const int chunk = rem / G::kPx, hp = rem - chunk * G::kPx;
            const int py = hp / G::kW, px = hp - py * G::kW;
            int Y, X;
            if (STRIDE == 1) Y = Y0 + py - P.pad, X = X0 + px - P.pad;
            else Y = 2 * (Y0 + py) + (plane >> 1), X = 2 * (X0 + px) + (plane & 1);
            const bool ok = in && Y >= 0 && X >= 0 && Y < P.b_h && X < P.b_w;
            const int c0 = cb0 + chunk * 8;
            const float *src = bb + (size_t)c0 * b_plane + (size_t)Y * P.b_w + X;
#pragma unroll
            for (int e = 0; e < 8; ++e) vb[j][e] = (ok && c0 + e < P.Cb) ? __ldg(src + (size_t)e * b_plane) : 0.0f;
            dst[j] = in ? plane * G::kPlaneBytes + chunk * G::kChunk + hp * 16 : -1;
            chs[j] = chunk * 8;
          }
#pragma unroll
          for (int j = 0; j < kBatch; ++j) {
            if (dst[j] >= 0) {
              const float *s8 = sc + chs[j];
              *reinterpret_cast<uint4 *>(Bm + dst[j]) =
                  make_uint4(pack2<true>(vb[j][0] * s8[0], vb[j][1] * s8[1]), pack2<true>(vb[j][2] * s8[2], vb[j][3] * s8[3]),
                             pack2<true>(vb[j][4] * s8[4], vb[j][5] * s8[5]), pack2<true>(vb[j][6] * s8[6], vb[j][7] * s8[7]));
            }
          }
        }
      }
      fence_async_smem();
      mbar_arrive_wg(bar_full + st * 8);
    }
    // ---- epilogue: accumulator block `tap` holds [A channel = TMEM lane][B channel = column]; dw is [Cout][Cin][kh][kw]
    if (it > 0) {
      mbar_wait_spin(bar_acc, 0);
      tc_fence_after();
      const int wq = warp_u & 3, part = warp_u >> 2;          // 4 warps per lane quadrant share the column groups
      const int row = wq * 32 + (tid & 31), ca_e = ca0 + row;
      const uint32_t trow = tmem_acc + ((uint32_t)(wq * 32) << 16);
      const bool vec_ok = P.dw_vec != 0 && P.out_t == 0;
#pragma unroll 1
      for (int gi = part; gi < kNb / 4; gi += kStageThreads / 128) {
        uint32_t r[taps][4];
#pragma unroll
        for (int tap = 0; tap < taps; ++tap) HAV_TMEM_LD4(r[tap], trow + tap * kNb + gi * 4);
        tmem_wait_ld();
        const int cb = cb0 + gi * 4;
        if (ca_e < P.Ca && cb < P.Cb) {
          if (vec_ok && cb + 4 <= P.Cb) {
            float *dst = P.dw + ((size_t)ca_e * P.Cb + cb) * taps;
            float f[taps * 4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
              for (int tap = 0; tap < taps; ++tap) f[e * taps + tap] = __uint_as_float(r[tap][e]) * P.wscale;
#pragma unroll
            for (int j = 0; j < taps * 4; j += 4) red_add_v4(dst + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (cb + e < P.Cb) {
                float *dst = P.out_t ? P.dw + ((size_t)(cb + e) * P.Ca + ca_e) * taps : P.dw + ((size_t)ca_e * P.Cb + cb + e) * taps;
#pragma unroll
                for (int tap = 0; tap < taps; ++tap) red_add(dst + tap, __uint_as_float(r[tap][e]) * P.wscale);
              }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_u == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols));
}

// out[row, i] = a[row, i] * scale[row]  and  dot[row] = sum_i a[row, i] * x[row, i]  in one pass over the rows of two [rows, n]
// fp32 matrices (a row = one (sample, channel) image).  Finishes the data gradient of a modulated convolution:
// dx = s * dxs and ds = sum_p x * dxs from the un-modulated data gradient dxs; also the demodulation gradient sum_p g * y.
__global__ void __launch_bounds__(256) rowscale_dot_kernel(float *__restrict__ out, float *__restrict__ dot, const float *__restrict__ a,
                                                           const float *__restrict__ x, const float *__restrict__ scale, long rows, long n) {
  __shared__ float red[8];
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float *ar = a + row * n, *xr = x != nullptr ? x + row * n : nullptr;
    float *orow = out != nullptr ? out + row * n : nullptr;
    const float s = scale != nullptr ? __ldg(scale + row) : 1.0f;
    float acc = 0.0f;
    if ((n & 3) == 0) {
      for (long i = threadIdx.x * 4L; i < n; i += 1024) {
        const float4 av = *reinterpret_cast<const float4 *>(ar + i);
        if (xr != nullptr) {
          const float4 xv = __ldg(reinterpret_cast<const float4 *>(xr + i));
          acc = fmaf(av.x, xv.x, fmaf(av.y, xv.y, fmaf(av.z, xv.z, fmaf(av.w, xv.w, acc))));
        }
        if (orow != nullptr) *reinterpret_cast<float4 *>(orow + i) = make_float4(av.x * s, av.y * s, av.z * s, av.w * s);
      }
    } else {
      for (long i = threadIdx.x; i < n; i += 256) {
        const float av = ar[i];
        if (xr != nullptr) acc = fmaf(av, __ldg(xr + i), acc);
        if (orow != nullptr) orow[i] = av * s;
      }
    }
    if (dot != nullptr) {
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) dot[row] = red[0] + red[1] + red[2] + red[3] + red[4] + red[5] + red[6] + red[7];
      __syncthreads();
    }
  }
}

}  // namespace wg
}  // namespace hav

using namespace hav;

extern "C" int hav_rowscale_dot(float *out, float *dot, const float *a, const float *x, const float *scale, int64_t rows, int64_t n,
                                void *stream) {
  if (a == nullptr || (out == nullptr && dot == nullptr) || (dot != nullptr && x == nullptr)) return HAV_E_NULL;
  if (rows < 0 || n < 0) return HAV_E_SHAPE;
  if (rows == 0 || n == 0) {
    if (dot != nullptr && rows > 0) cudaMemsetAsync(dot, 0, rows * sizeof(float), (cudaStream_t)stream);
    return HAV_OK;
  }
  const int grid = (int)(rows < 148L * 16 ? rows : 148L * 16);
  wg::rowscale_dot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, dot, a, x, scale, rows, n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_conv2d_wgrad(const hav_conv_wgrad_args *a, void *stream) {
  if (a == nullptr) return HAV_E_NULL;
  if (a->struct_bytes != sizeof(hav_conv_wgrad_args)) return HAV_E_VALUE;
  if (a->batch < 0 || a->cin < 1 || a->cout < 1 || a->in_h < 1 || a->in_w < 1) return HAV_E_SHAPE;
  if ((a->ksize != 1 && a->ksize != 3) || (a->up != 1 && a->up != 2) || (a->down != 1 && a->down != 2)) return HAV_E_VALUE;
  if ((a->up == 2 || a->down == 2) && (a->up == a->down || a->ksize != 3)) return HAV_E_VALUE;
  if (a->dw == nullptr) return HAV_E_NULL;
  const int k = a->ksize, taps = k * k;
  if (!a->accumulate) {
    cudaError_t e0 = cudaMemsetAsync(a->dw, 0, (size_t)a->cout * a->cin * taps * sizeof(float), (cudaStream_t)stream);
    if (e0 != cudaSuccess) return (int)e0;
  }
  if (a->batch == 0) return HAV_OK;
  if (a->g == nullptr || a->x == nullptr) return HAV_E_NULL;
  wg::WgDev P;
  memset(&P, 0, sizeof(P));
  P.B = a->batch;
  int stride = 1;
  if (a->up == 2) {           // y = conv_transpose2d(x, stride 2, pad 0), g is (2H+1) x (2W+1):  dw = sum_p x[p] g[2p + tap]
    P.Ca = a->cin, P.Cb = a->cout, P.pa = a->x, P.pb = a->g, P.sa = a->in_scale, P.sb = a->out_scale;
    P.a_h = a->in_h, P.a_w = a->in_w, P.b_h = 2 * a->in_h + 1, P.b_w = 2 * a->in_w + 1, P.out_t = 1, stride = 2;
  } else {
    P.Ca = a->cout, P.Cb = a->cin, P.pa = a->g, P.pb = a->x, P.sa = a->out_scale, P.sb = a->in_scale;
    P.b_h = a->in_h, P.b_w = a->in_w, P.out_t = 0;
    if (a->down == 2) {       // y = conv2d(x, stride 2, pad 0), g is ((H-3)/2+1) x ((W-3)/2+1):  dw = sum_p g[p] x[2p + tap]
      if (a->in_h < 3 || a->in_w < 3) return HAV_E_SHAPE;
      P.a_h = (a->in_h - 3) / 2 + 1, P.a_w = (a->in_w - 3) / 2 + 1, stride = 2;
    } else {                  // stride 1, pad k/2:  dw = sum_p g[p] x[p + tap - pad]
      P.a_h = a->in_h, P.a_w = a->in_w, P.pad = k / 2;
    }
  }
  P.tiles_x = (P.a_w + wg::kTW - 1) / wg::kTW, P.tiles_y = (P.a_h + wg::kTH - 1) / wg::kTH;
  const long ntiles = (long)a->batch * P.tiles_x * P.tiles_y;
  if (ntiles > 2147483647L) return HAV_E_SHAPE;
  P.ntiles = (int)ntiles;
  P.a_tiles = (P.Ca + wg::kMa - 1) / wg::kMa, P.b_tiles = (P.Cb + wg::kNb - 1) / wg::kNb;
  const long blocks = (long)P.a_tiles * P.b_tiles;
  if (blocks > 2147483647L) return HAV_E_SHAPE;
  // split-K so that the grid is at most two FULL waves of one CTA per SM (512 TMEM columns each): a third, nearly empty
  // wave would cost as much as a full one; one wave when a CTA would otherwise see only a handful of tiles
  long nsplit = (2 * 148) / blocks;
  if (nsplit < 1) nsplit = 1;
  if (ntiles / nsplit < 6 && 148 / blocks >= 1) nsplit = 148 / blocks;
  if (nsplit > ntiles) nsplit = ntiles;
  if (nsplit > 65535) nsplit = 65535;
  if (nsplit < 1) nsplit = 1;
  P.nsplit = (int)nsplit;
  P.dw = a->dw, P.wscale = a->wscale;
  {
    static const int dbg = getenv("HAV_WG_DEBUG") != nullptr ? atoi(getenv("HAV_WG_DEBUG")) : 0;
    P.dbg = dbg;
  }
  P.a_vec = (P.a_w & 3) == 0 && ((uintptr_t)P.pa & 15) == 0;   // every 4-px quarter of a tile row is a 16-byte aligned float4
  P.dw_vec = (((size_t)a->cin * taps) & 3) == 0 && ((uintptr_t)a->dw & 15) == 0;
  dim3 grid((unsigned)blocks, (unsigned)nsplit);
  cudaError_t e;
  auto launch = [&](auto kern, int smem_bytes) -> cudaError_t {
    cudaError_t er = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (er != cudaSuccess) return er;
    kern<<<grid, wg::kThreads, smem_bytes, (cudaStream_t)stream>>>(P);
    return cudaGetLastError();
  };
  if (k == 1) e = launch(wg::conv_wgrad_kernel<1, 1>, wg::Geo<1>::kSmemBytes);
  else if (stride == 1) e = launch(wg::conv_wgrad_kernel<3, 1>, wg::Geo<1>::kSmemBytes);
  else e = launch(wg::conv_wgrad_kernel<3, 2>, wg::Geo<2>::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}
