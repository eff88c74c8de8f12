// Internal (non-ABI) declarations shared by the render translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/havatar_b200.h"

namespace hav {

struct RenderDev;

int render_check_args(const hav_render_args *a);              // render_api.cu
void render_fill_dev(const hav_render_args *a, RenderDev &P);  // everything except the packed-weight pointers

// ---- fp32 packed-weight blob (float offsets), built by pack_mlp_fp32_kernel ----
constexpr int kOffW0t = 0;                       // [176][128]
constexpr int kOffW1t = kOffW0t + 176 * 128;     // [128][128]
constexpr int kOffWht = kOffW1t + 128 * 128;     // [128][68]
constexpr int kOffB0 = kOffWht + 128 * 68;       // [128]
constexpr int kOffB1 = kOffB0 + 128;             // [128]
constexpr int kOffBh = kOffB1 + 128;             // [68]
constexpr int kOffWr = kOffBh + 68;              // [3][64]
constexpr int kOffBr = kOffWr + 192;             // [3] (+1 pad)
constexpr int kPackF32Floats = kOffBr + 4;

void launch_pack_mlp_fp32(const hav_render_args *a, float *out, cudaStream_t st);
cudaError_t launch_render_fp32(const RenderDev &P, int num_blocks, cudaStream_t st);

// ---- 16-bit tcgen05 path (render_tc.cu) ----
uint64_t tc_weight_image_bytes();
uint64_t tc_planes_bytes(int nimg, int H, int W);
int tc_num_ctas(int num_ray_blocks);
int tc_scratch_slots(int num_ray_blocks);
void launch_pack_mlp_16(const hav_render_args *a, uint8_t *wimg, cudaStream_t st, int32_t *status = nullptr);
cudaError_t launch_pack_planes_16(const float *planes, uint16_t *out, int nimg, int H, int W, bool bf16, cudaStream_t st,
                                  int32_t *status = nullptr);
cudaError_t launch_render_16(const RenderDev &P, int num_ray_blocks, bool bf16, cudaStream_t st);
cudaError_t launch_render_16_v2(const RenderDev &P, int num_ray_blocks, bool bf16, cudaStream_t st);  // render_tc2.cu
// render_tc3.cu: CTA pairs (cta_group::2); mode 0 fp16, 1 bf16, 2 fp16 hi + lo split
cudaError_t launch_render_16_v3(const RenderDev &P, int num_ray_blocks, int mode, cudaStream_t st);
int tc3_scratch_slots(int num_ray_blocks);
void launch_pack_mlp_16_lo(const hav_render_args *a, uint8_t *wimg_lo, cudaStream_t st);   // fp16(w - fp16(w)) image
cudaError_t launch_pack_planes_f32(const float *planes, float *out, int nimg, int H, int W, cudaStream_t st);

}  // namespace hav
