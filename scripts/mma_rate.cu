// Micro-benchmark: sustained rate of one SM's tcgen05.mma stream for the operand placements and tile widths the kernels of this
// repo use, alone and beside shared-memory traffic from other warps.  One CTA per SM on every SM; one thread issues `iters`
// K = 16 MMAs (M = 128) into one accumulator and waits for the last commit.  Operands are whatever the buffers hold.
//   SS: A and B through shared-memory descriptors (canonical no-swizzle K-major layout, as render_tc2 L0 / conv_tc use)
//   TS: A from TMEM (as render_tc2 L1 / head and conv_wgrad use)
//   sm128 / TM128: the same with 128-byte-swizzled K-major shared-memory operands
// Build + run (B200):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I havatar_b200/csrc scripts/mma_rate.cu -o /tmp/mma_rate && /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace hav::tc;

constexpr int kKSteps = 4;                  // operand buffers hold K = 64; the stream cycles through the 4 K-steps
constexpr int kASz = 8 * 128 * 16;          // 16 KB
constexpr int kBSzMax = 8 * 256 * 16;       // 32 KB
constexpr int kTraffic = 64 * 1024;         // region the traffic warps read / write
constexpr int kSmem = kASz + kBSzMax + kTraffic + 256;

// K-major SWIZZLE_128B descriptor: rows of 64 16-bit elements (128 B), 8-row groups 1024 B apart, K = 16 step = +32 B
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(32 + 256, 1) mma_rate_kernel(int mode, int N, int iters, int traffic_warps, int traffic_store,
                                                                unsigned long long *out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + kASz + kBSzMax + kTraffic;
  volatile uint32_t *slot = reinterpret_cast<volatile uint32_t *>(smem + kASz + kBSzMax + kTraffic + 64);
  volatile uint32_t *stop = reinterpret_cast<volatile uint32_t *>(smem + kASz + kBSzMax + kTraffic + 128);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kASz + kBSzMax + kTraffic) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + kASz + kBSzMax + kTraffic + 64), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    *stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  unsigned long long bytes = 0;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = instr_desc(N, false);
      const uint32_t A = sb, B = sb + kASz;
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int k = i & (kKSteps - 1);
        const uint64_t bd = smem_desc(B + 2 * k * N * 16, N * 16, 128);
        if (mode == 0) umma_ss(tmem, smem_desc(A + 2 * k * 2048, 2048, 128), bd, idesc, 1);
        else if (mode == 1) umma_ts(tmem, tmem + 256 + k * 8, bd, idesc, 1);
        else if (mode == 2) umma_ss(tmem, desc_sw128(A + k * 32), desc_sw128(B + k * 32), idesc, 1);      // both operands 128B-swizzled
        else umma_ts(tmem, tmem + 256 + k * 8, desc_sw128(B + k * 32), idesc, 1);
      }
      umma_commit(bar);
      mbar_wait_spin(bar, 0);
      const long long t1 = clock64();
      *stop = 1;
      if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    }
    __syncwarp();
  } else if (warp <= traffic_warps) {
    // shared-memory traffic beside the MMA stream: conflict-free 16-byte loads (or stores) over a 64 KB region
    uint4 acc = make_uint4(0, 0, 0, 0);
    uint4 *base = reinterpret_cast<uint4 *>(smem + kASz + kBSzMax);
    int off = (warp - 1) * 32 + (tid & 31);
    while (*stop == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (traffic_store) base[off] = acc;
        else {
          const uint4 v = base[off];
          acc.x ^= v.x, acc.y ^= v.y, acc.z ^= v.z, acc.w ^= v.w;
        }
        off = (off + 256) & (kTraffic / 16 - 1);
      }
      bytes += 8 * 16;
    }
    if (acc.x == 0x12345u) out[3] = acc.y;   // keep the loads alive
    if (blockIdx.x == 0) atomicAdd(&out[1], bytes);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// Second measurement: what one SM can pull from L2 with cp.async.bulk when every SM streams the same weight-like buffer (each
// CTA walks `n` slices of `bytes` through a ring of `stages` slots; no MMAs).  This is the supply side of conv_tc's weight ring.
__global__ void __launch_bounds__(32, 1) bulk_rate_kernel(const uint8_t *src, int src_slices, int bytes, int n, int stages, unsigned long long *out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sb = smem_u32(smem), bar = sb + 160 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(bar + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long t0 = clock64();
    int first = (blockIdx.x * 7) % src_slices;
    for (int i = 0; i < n + stages; ++i) {
      const int slot = i % stages;
      if (i >= stages) mbar_wait_spin(bar + slot * 8, ((i - stages) / stages) & 1);     // slice i - stages has landed
      if (i < n) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + slot * 8), "r"((uint32_t)bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb + slot * bytes),
                     "l"(src + (size_t)((first + i) % src_slices) * bytes), "r"((uint32_t)bytes), "r"(bar + slot * 8) : "memory");
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  unsigned long long *out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  const int iters = 1 << 15;
  printf("tcgen05.mma kind::f16, M = 128, K = 16 per instruction, %d SMs busy, %d MMAs per SM; dense fp16 peak = 8192 FLOP/clk/SM\n", sms, iters);
  printf("%-4s %4s %-22s %12s %12s %14s\n", "A", "N", "smem traffic beside", "clk/MMA", "% of peak", "traffic B/clk");
  struct Cfg { int mode, N, tw, st; };
  const Cfg cfgs[] = {{2, 256, 0, 0}, {2, 128, 0, 0}, {2, 64, 0, 0}, {3, 128, 0, 0}, {2, 128, 4, 0}, {2, 128, 4, 1}, {0, 256, 0, 0}, {0, 128, 0, 0}, {0, 80, 0, 0}, {0, 64, 0, 0}, {1, 256, 0, 0}, {1, 128, 0, 0}, {1, 80, 0, 0},
                      {0, 128, 4, 0}, {0, 128, 8, 0}, {0, 128, 4, 1}, {1, 128, 4, 0}, {1, 128, 8, 0}, {1, 128, 4, 1}};
  for (int w = 0; w < 200; ++w) mma_rate_kernel<<<sms, 32 + 256, kSmem>>>(1, 256, iters, 0, 0, out);   // ~0.5 s: clocks and power state settle
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int pass = 0; pass < 2; ++pass) {
    printf("pass %d\n", pass);
    for (const Cfg &c : cfgs) {
      unsigned long long h[4] = {0, 0, 0, 0};
      double best_clk = 1e30, best_ms = 1e30, traffic = 0;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(out, 0, 64);
        cudaEventRecord(e0);
        mma_rate_kernel<<<sms, 32 + 256, kSmem>>>(c.mode, c.N, iters, c.tw, c.st, out);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if ((double)h[0] / iters < best_clk) best_clk = (double)h[0] / iters, traffic = (double)h[1] / (double)h[0];
        if (ms < best_ms) best_ms = ms;
      }
      const double ideal = 128.0 * c.N * 16 * 2 / 8192.0;
      char tr[32];
      snprintf(tr, sizeof tr, c.tw ? "%d warps %s" : "none", c.tw, c.st ? "STS.128" : "LDS.128");
      printf("%-5s %4d %-20s %10.1f %9.1f%% %12.1f %10.0f TFLOP/s (whole kernel, events)\n",
             c.mode == 0 ? "smem" : c.mode == 1 ? "TMEM" : c.mode == 2 ? "sm128" : "TM128", c.N, tr, best_clk, 100.0 * ideal / best_clk, traffic,
             (double)sms * iters * 128.0 * c.N * 16 * 2 / (best_ms * 1e-3) / 1e12);
    }
  }
  {
    const int bytes = 16384, src_slices = 288;       // 4.7 MB: one 512 x 512 x 3 x 3 fp16 weight tensor, L2 resident
    uint8_t *src;
    cudaMalloc(&src, (size_t)bytes * src_slices);
    cudaMemset(src, 0, (size_t)bytes * src_slices);
    cudaFuncSetAttribute(bulk_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024 + 256);
    printf("cp.async.bulk L2 -> shared memory, every SM streaming the same 4.7 MB buffer in 16 KB slices\n%8s %8s %12s %12s\n", "CTAs", "in flight", "B/clk/SM", "chip TB/s");
    for (int ctas : {sms, 128, 64, 16})
      for (int stages : {2, 4, 8}) {
        unsigned long long h[4];
        const int n = 2048;
        for (int rep = 0; rep < 2; ++rep) {
          bulk_rate_kernel<<<ctas, 32, 160 * 1024 + 256>>>(src, src_slices, bytes, n, stages, out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("bulk kernel failed\n"); return 1; }
        }
        cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
        const double bpc = (double)n * bytes / (double)h[0];
        printf("%8d %8d %12.1f %12.2f\n", ctas, stages, bpc, bpc * ctas * khz * 1e3 / 1e12);
      }
  }
  return 0;
}
