// placeholder until the tcgen05 kernel lands: the 16-bit precisions report HAV_E_VALUE.
#include "render_common.cuh"
#include "render_internal.h"

namespace hav {
uint64_t tc_weight_image_bytes() { return 0; }
void launch_pack_mlp_bf16(const hav_render_args *, uint8_t *, cudaStream_t) {}
void launch_pack_planes_bf16(const float *, uint16_t *, int, int, int, int, cudaStream_t) {}
cudaError_t launch_render_bf16(const RenderDev &, int, cudaStream_t) { return cudaErrorNotSupported; }
int tc_num_ctas(int n) { return n; }
}  // namespace hav
