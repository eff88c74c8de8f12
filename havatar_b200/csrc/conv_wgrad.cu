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
// The strided convolutions reuse the kernel through zero insertion while staging:
//   down = 2 (stride 2, pad 0):  dW = sum_P G2[P] x[P + (kh,kw)],   G2 = g with zeros inserted between its pixels (a_step = 2)
//   up   = 2 (conv_transpose2d): dW = sum_P g[P] X2[P - (kh,kw)],   X2 = x with zeros inserted (x_step = 2), taps mirrored
// (4x the necessary MMA work on those few layers, none on the stride-1 layers that dominate).
#include "tc_common.cuh"

namespace hav {
namespace wg {

using namespace tc;

constexpr int kTH = 8, kTW = 16;                   // position tile: 8 rows x 16 px = 8 K-steps of 16
constexpr int kHW = kTW + 2, kHH = kTH + 2, kHaloPx = kHH * kHW;   // 180
constexpr int kNci = 48, kMco = 128;
constexpr int kABytes = (kTH * kTW / 8) * kMco * 16;   // 32768
constexpr int kBChunk = kHaloPx * 16;                  // 2880
constexpr int kBBytes = (kNci / 8) * kBChunk;          // 17280
constexpr int kStages = 3;
constexpr int kSmA = 0;
constexpr int kSmB = kSmA + kStages * kABytes;
constexpr int kSmScale = kSmB + kStages * kBBytes;     // [kStages][48] float
constexpr int kSmBar = kSmScale + kStages * kNci * 4;
constexpr int kSmemBytes = kSmBar + 128;
constexpr int kStageThreads = 256, kThreads = kStageThreads + 32;
constexpr int kTmemCols = 512;
constexpr uint32_t kBMajorMN = 1u << 16;

struct WgDev {
  int B, Cin, Cout, k, taps;
  int GH, GW, g_h, g_w, a_step;       // virtual position grid, real size of g, zero-insertion step of g
  int x_h, x_w, x_step, x_off;        // real size of x, zero-insertion step of x, halo origin = tile origin + x_off
  int a_vec, dw_vec;                  // 16-byte vector loads of g / vector reductions into dw are aligned
  int mirror;                         // tap (kh,kw) reads the halo at (2-kh, 2-kw) instead of (kh,kw)
  int tiles_x, tiles_y, ntiles, nsplit, co_tiles, ci_tiles;
  const float *g, *x, *in_scale, *out_scale;
  float *dw;
  float wscale;
};

__device__ __forceinline__ void mbar_arrive_wg(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

#define HAV_TMEM_LD8(r, taddr)                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                 \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(taddr))

template <int KS>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_kernel(const WgDev P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_base + kSmBar, bar_free = bar_full + kStages * 8, bar_acc = bar_free + kStages * 8;
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSmBar + 120);
  constexpr int taps = KS * KS;

  const int cot = blockIdx.x % P.co_tiles, cit = blockIdx.x / P.co_tiles;
  const int split = blockIdx.y;
  const int co0 = cot * kMco, ci0 = cit * kNci;

  if (warp_u == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + kSmBar + 120), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int i = 0; i < kStages; ++i) mbar_init(bar_full + i * 8, kStageThreads), mbar_init(bar_free + i * 8, 1);
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
      constexpr uint32_t idesc = instr_desc(kNci, true) | kBMajorMN;
      int it = 0;
      for (int t = split; t < P.ntiles; t += P.nsplit, ++it) {
        const int st = it % kStages;
        mbar_wait_spin(bar_full + st * 8, (it / kStages) & 1);
        tc_fence_after();
        const uint32_t A0 = smem_base + kSmA + st * kABytes, B0 = smem_base + kSmB + st * kBBytes;
#pragma unroll 1
        for (int r = 0; r < kTH; ++r) {
          const uint64_t adesc = smem_desc(A0 + 2 * r * (kMco * 16), kMco * 16, 128);
#pragma unroll
          for (int tap = 0; tap < taps; ++tap) {
            int kh = tap / KS, kw = tap - kh * KS;
            if (P.mirror) kh = KS - 1 - kh, kw = KS - 1 - kw;
            // MN-major: LBO = distance between 8-pixel K groups (128 B), SBO = distance between 8-channel groups (one chunk)
            const uint64_t bdesc = smem_desc(B0 + ((r + kh) * kHW + kw) * 16, 128, kBChunk);
            umma_ss(tmem_acc + tap * kNci, adesc, bdesc, idesc, (it > 0 || r > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_free + st * 8);
      }
      umma_commit(bar_acc);
    }
    __syncwarp();
  } else {
    // ================= staging warps =================
    const size_t g_plane = (size_t)P.g_h * P.g_w, x_plane = (size_t)P.x_h * P.x_w;
    const int co_l = tid & (kMco - 1), co = co0 + co_l;
    const bool a_vec = P.a_vec != 0;
    int it = 0;
    for (int t = split; t < P.ntiles; t += P.nsplit, ++it) {
      const int st = it % kStages;
      int sp = t;
      const int tx = sp % P.tiles_x; sp /= P.tiles_x;
      const int ty = sp % P.tiles_y;
      const int b = sp / P.tiles_y;
      const int Y0 = ty * kTH, X0 = tx * kTW;
      if (it >= kStages) mbar_wait_spin(bar_free + st * 8, ((it / kStages) - 1) & 1);
      float *sc = reinterpret_cast<float *>(smem + kSmScale) + st * kNci;
      if (tid < kNci) {
        const int ci = ci0 + tid;
        sc[tid] = ci < P.Cin ? (P.in_scale != nullptr ? __ldg(P.in_scale + (size_t)b * P.Cin + ci) : 1.0f) : 0.0f;
      }
      // ---- A: d*g, K-major.  unit = (8-px chunk c = 2*row + half, channel); this thread's channel is fixed
      {
        uint8_t *A = smem + kSmA + st * kABytes;
        const bool co_ok = co < P.Cout;
        const float as = co_ok ? (P.out_scale != nullptr ? __ldg(P.out_scale + (size_t)b * P.Cout + co) : 1.0f) : 0.0f;
        const float *gb = P.g + ((size_t)b * P.Cout + (co_ok ? co : 0)) * g_plane;
#pragma unroll 2
        for (int c = tid >> 7; c < 16; c += 2) {
          const int Y = Y0 + (c >> 1), X = X0 + (c & 1) * 8;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.0f;
          if (co_ok && Y < P.GH) {
            if (a_vec && X + 8 <= P.GW) {
              const float4 p = __ldg(reinterpret_cast<const float4 *>(gb + (size_t)Y * P.g_w + X));
              const float4 q = __ldg(reinterpret_cast<const float4 *>(gb + (size_t)Y * P.g_w + X + 4));
              v[0] = p.x, v[1] = p.y, v[2] = p.z, v[3] = p.w, v[4] = q.x, v[5] = q.y, v[6] = q.z, v[7] = q.w;
            } else if (P.a_step == 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (X + e < P.GW) v[e] = __ldg(gb + (size_t)Y * P.g_w + X + e);
            } else if (!(Y & 1)) {      // zero-inserted g: virtual (Y, X) holds g[Y/2, X/2] when both are even (X0 is even)
#pragma unroll
              for (int e = 0; e < 8; e += 2)
                if (X + e < P.GW) v[e] = __ldg(gb + (size_t)(Y >> 1) * P.g_w + ((X + e) >> 1));
            }
          }
          *reinterpret_cast<uint4 *>(A + c * (kMco * 16) + co_l * 16) =
              make_uint4(pack2<true>(v[0] * as, v[1] * as), pack2<true>(v[2] * as, v[3] * as), pack2<true>(v[4] * as, v[5] * as),
                         pack2<true>(v[6] * as, v[7] * as));
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kStageThreads) : "memory");   // the scale table of this stage is complete
      // ---- B: s*x halo, MN-major.  unit = (8-channel chunk, halo pixel); consecutive threads take consecutive pixels
      {
        uint8_t *Bm = smem + kSmB + st * kBBytes;
        const float *xb = P.x + (size_t)b * P.Cin * x_plane;
        int hp = tid, chunk = 0;
        while (hp >= kHaloPx) hp -= kHaloPx, ++chunk;
#pragma unroll 1
        for (; chunk < kNci / 8;) {
          const int py = hp / kHW, px = hp - py * kHW;
          int Y = Y0 + py + P.x_off, X = X0 + px + P.x_off;
          bool ok = Y >= 0 && X >= 0;
          if (P.x_step == 2) ok = ok && !(Y & 1) && !(X & 1), Y >>= 1, X >>= 1;
          ok = ok && Y < P.x_h && X < P.x_w;
          const int c0 = ci0 + chunk * 8;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = (ok && c0 + e < P.Cin) ? __ldg(xb + (size_t)(c0 + e) * x_plane + (size_t)Y * P.x_w + X) : 0.0f;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= sc[chunk * 8 + e];
          *reinterpret_cast<uint4 *>(Bm + chunk * kBChunk + hp * 16) =
              make_uint4(pack2<true>(v[0], v[1]), pack2<true>(v[2], v[3]), pack2<true>(v[4], v[5]), pack2<true>(v[6], v[7]));
          hp += kStageThreads;
          while (hp >= kHaloPx) hp -= kHaloPx, ++chunk;
        }
      }
      fence_async_smem();
      mbar_arrive_wg(bar_full + st * 8);
    }
    // ---- epilogue: accumulator block `tap` holds dW[co lane][ci column]; memory order is [co][ci][kh][kw]
    if (it > 0) {
      mbar_wait_spin(bar_acc, 0);
      tc_fence_after();
      const int wq = warp_u & 3, half = warp_u >> 2;
      const int row = wq * 32 + (tid & 31), co_e = co0 + row;
      const uint32_t trow = tmem_acc + ((uint32_t)(wq * 32) << 16);
      const bool vec_ok = P.dw_vec != 0;
#pragma unroll 1
      for (int gi = half; gi < kNci / 8; gi += 2) {
        uint32_t r[taps][8];
#pragma unroll
        for (int tap = 0; tap < taps; ++tap) HAV_TMEM_LD8(r[tap], trow + tap * kNci + gi * 8);
        tmem_wait_ld();
        const int cb = ci0 + gi * 8;
        if (co_e < P.Cout && cb < P.Cin) {
          float *dst = P.dw + ((size_t)co_e * P.Cin + cb) * taps;
          if (vec_ok && cb + 8 <= P.Cin) {
            float f[taps * 8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
#pragma unroll
              for (int tap = 0; tap < taps; ++tap) f[e * taps + tap] = __uint_as_float(r[tap][e]) * P.wscale;
#pragma unroll
            for (int j = 0; j < taps * 8; j += 4) red_add_v4(dst + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (cb + e < P.Cin) {
#pragma unroll
                for (int tap = 0; tap < taps; ++tap) atomicAdd(dst + e * taps + tap, __uint_as_float(r[tap][e]) * P.wscale);
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
  P.B = a->batch, P.Cin = a->cin, P.Cout = a->cout, P.k = k, P.taps = taps;
  P.x_h = a->in_h, P.x_w = a->in_w;
  if (a->up == 2) {           // y = conv_transpose2d(x, stride 2, pad 0): g is (2H+1) x (2W+1)
    P.g_h = 2 * a->in_h + 1, P.g_w = 2 * a->in_w + 1;
    P.GH = P.g_h, P.GW = P.g_w, P.a_step = 1, P.x_step = 2, P.x_off = -2, P.mirror = 1;
  } else if (a->down == 2) {  // y = conv2d(x, stride 2, pad 0): g is ((H-3)/2+1) x ((W-3)/2+1)
    if (a->in_h < 3 || a->in_w < 3) return HAV_E_SHAPE;
    P.g_h = (a->in_h - 3) / 2 + 1, P.g_w = (a->in_w - 3) / 2 + 1;
    P.GH = 2 * P.g_h - 1, P.GW = 2 * P.g_w - 1, P.a_step = 2, P.x_step = 1, P.x_off = 0, P.mirror = 0;
  } else {
    P.g_h = a->in_h, P.g_w = a->in_w;
    P.GH = P.g_h, P.GW = P.g_w, P.a_step = 1, P.x_step = 1, P.x_off = -(k / 2), P.mirror = 0;
  }
  P.tiles_x = (P.GW + wg::kTW - 1) / wg::kTW, P.tiles_y = (P.GH + wg::kTH - 1) / wg::kTH;
  const long ntiles = (long)a->batch * P.tiles_x * P.tiles_y;
  if (ntiles > 2147483647L) return HAV_E_SHAPE;
  P.ntiles = (int)ntiles;
  P.co_tiles = (a->cout + wg::kMco - 1) / wg::kMco, P.ci_tiles = (a->cin + wg::kNci - 1) / wg::kNci;
  const long blocks = (long)P.co_tiles * P.ci_tiles;
  if (blocks > 2147483647L) return HAV_E_SHAPE;
  long nsplit = (2 * 148 + blocks - 1) / blocks;      // about two waves of CTAs (one CTA per SM: 512 TMEM columns each)
  if (nsplit > ntiles) nsplit = ntiles;
  if (nsplit > 65535) nsplit = 65535;
  if (nsplit < 1) nsplit = 1;
  P.nsplit = (int)nsplit;
  P.g = a->g, P.x = a->x, P.in_scale = a->in_scale, P.out_scale = a->out_scale, P.dw = a->dw, P.wscale = a->wscale;
  P.a_vec = P.a_step == 1 && (P.g_w & 3) == 0 && ((uintptr_t)a->g & 15) == 0;
  P.dw_vec = (((size_t)a->cin * taps) & 3) == 0 && ((uintptr_t)a->dw & 15) == 0;
  dim3 grid((unsigned)blocks, (unsigned)nsplit);
  cudaError_t e;
  if (k == 3) {
    e = cudaFuncSetAttribute(wg::conv_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    wg::conv_wgrad_kernel<3><<<grid, wg::kThreads, wg::kSmemBytes, (cudaStream_t)stream>>>(P);
  } else {
    e = cudaFuncSetAttribute(wg::conv_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    wg::conv_wgrad_kernel<1><<<grid, wg::kThreads, wg::kSmemBytes, (cudaStream_t)stream>>>(P);
  }
  e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}
