// Adam over ONE flat fp32 parameter buffer (torch.optim.Adam semantics: train_avatar.py:68-71, train_avatarHD.py:117-122 build
// torch Adam optimisers; torch runs them as ~14 multi-tensor launches over several hundred tensors per step).
//
// The data-parallel layout already keeps every gradient in a few flat buckets (havatar_b200/parallel.py); with parameters and
// both moments flat as well, an optimiser step is a single HBM-bound pass:  read p, g, m, v  ->  write p, m, v (and g = 0, so
// the separate zero_grad pass disappears): 32 B per parameter, 16-byte vector accesses, grid = a few CTAs per SM.
// The step counter and the learning rate live in device memory (state[0], state[1]) so the launch can be captured in a CUDA
// graph and the host can change the learning rate between replays without re-capturing.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/havatar_b200.h"

namespace hav {
namespace optim {

__global__ void adam_tick_kernel(float *state) { state[0] += 1.0f; }

__global__ void __launch_bounds__(256) adam_flat_kernel(float4 *__restrict__ p, float4 *__restrict__ g, float4 *__restrict__ m,
                                                        float4 *__restrict__ v, long n4, const float *__restrict__ state, float beta1,
                                                        float beta2, float eps, float grad_scale, int zero_grad) {
  const float step = state[0], lr = state[1];
  // torch/optim/adam.py (_multi_tensor_adam): step_size = lr / (1 - beta1^t); denom = sqrt(v) / sqrt(1 - beta2^t) + eps
  const float bc1 = 1.0f - powf(beta1, step), bc2s = sqrtf(1.0f - powf(beta2, step));
  const float step_size = lr / bc1, inv_bc2s = 1.0f / bc2s;
  const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 pv = p[i], gv = g[i], mv = m[i], vv = v[i];
#define HAV_ADAM1(c)                                                 \
  {                                                                  \
    const float gg = gv.c * grad_scale;                              \
    mv.c = mv.c + (gg - mv.c) * (1.0f - beta1);                      \
    vv.c = vv.c * beta2 + (1.0f - beta2) * gg * gg;                  \
    pv.c = pv.c - step_size * (mv.c / (sqrtf(vv.c) * inv_bc2s + eps)); \
  }
    HAV_ADAM1(x) HAV_ADAM1(y) HAV_ADAM1(z) HAV_ADAM1(w)
#undef HAV_ADAM1
    p[i] = pv, m[i] = mv, v[i] = vv;
    if (zero_grad) g[i] = z;
  }
}

}  // namespace optim
}  // namespace hav

extern "C" int hav_adam_flat(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float *state, float beta1,
                             float beta2, float eps, float grad_scale, int zero_grad, void *stream) {
  if (param == nullptr || grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || state == nullptr) return HAV_E_NULL;
  if (n < 0 || (n & 3) != 0) return HAV_E_SHAPE;        // the flat buffers are padded to a multiple of 4 elements
  if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) return HAV_E_VALUE;
  hav::optim::adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
  if (n > 0) {
    const long n4 = n / 4;
    long want = (n4 + 255) / 256;
    const int grid = (int)(want < 148L * 8 ? want : 148L * 8);
    hav::optim::adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4 *)param, (float4 *)grad, (float4 *)exp_avg,
                                                                        (float4 *)exp_avg_sq, n4, state, beta1, beta2, eps, grad_scale,
                                                                        zero_grad);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? HAV_OK : (int)e;
}
