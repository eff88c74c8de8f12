// C-ABI entry points of the fused render (include/havatar_b200.h): argument checks, workspace
// carving, weight/plane packing and kernel dispatch.  No allocation, no synchronisation.
#include <stdlib.h>
#include <string.h>

#include "render_common.cuh"
#include "render_internal.h"

namespace hav {

static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
  uint64_t pack_f32, wimg, planes_cl, zbuf, wbuf, total;
  int num_blocks, scratch_blocks;
};

int render_check_args(const hav_render_args *a) {
  if (a == nullptr) return HAV_E_NULL;
  if (a->struct_bytes != sizeof(hav_render_args)) return HAV_E_VALUE;
  if (a->precision != HAV_PREC_FP32 && a->precision != HAV_PREC_BF16 && a->precision != HAV_PREC_FP16 &&
      a->precision != HAV_PREC_FP16X3)
    return HAV_E_VALUE;
  if ((a->flags & ~(HAV_RENDER_REUSE_PACKED | HAV_RENDER_CHECK_RANGE | HAV_RENDER_CTA_PAIRS)) != 0) return HAV_E_VALUE;
  if (a->batch < 0 || a->rays < 0) return HAV_E_SHAPE;
  if ((int64_t)a->batch * a->rays > (int64_t)1 << 30) return HAV_E_SHAPE;
  if (a->num_coarse < 2 || a->num_coarse > kMaxSamples) return HAV_E_SHAPE;
  if (a->num_fine < 0 || a->num_fine > kMaxFine) return HAV_E_SHAPE;
  if (a->num_fine > 0 && a->num_coarse < 3) return HAV_E_SHAPE;
  if (a->num_fine > 0 && (a->num_coarse + 1) / 2 + a->num_fine > kMaxSamples) return HAV_E_SHAPE;
  if (a->plane_c != kPlaneC) return HAV_E_SHAPE;
  if (a->plane_h < 1 || a->plane_w < 1 || a->plane_h > 4096 || a->plane_w > 4096) return HAV_E_SHAPE;
  if (a->vol_d < 1 || a->vol_h < 1 || a->vol_w < 1) return HAV_E_SHAPE;
  if ((int64_t)2 * a->batch * (a->plane_h + 3) * (a->plane_w + 3) >= ((int64_t)1 << 31) / 64) return HAV_E_SHAPE;
  if ((int64_t)a->batch * a->rays == 0) return HAV_OK;
  if (a->camera != nullptr) {   // in-kernel ray generation
    if (a->img_h < 1 || a->img_w < 1 || (int64_t)a->img_h * a->img_w > (int64_t)1 << 30) return HAV_E_SHAPE;
    if (a->pixel_index == nullptr && (int64_t)a->rays != (int64_t)a->img_h * a->img_w) return HAV_E_SHAPE;
  } else if (a->ray_batch == nullptr) {
    return HAV_E_NULL;
  }
  const void *req[] = {a->inv_head_T, a->planes, a->wvol, a->w0, a->b0, a->w1, a->b1, a->w_alpha,
                       a->b_alpha, a->w_feat, a->b_feat, a->w_rgb, a->b_rgb, a->rgb_coarse, a->depth_coarse,
                       a->acc_coarse, a->weights_max};
  for (const void *p : req)
    if (p == nullptr) return HAV_E_NULL;
  if (a->num_fine > 0 && (a->rgb_fine == nullptr || a->depth_fine == nullptr || a->acc_fine == nullptr)) return HAV_E_NULL;
  if ((a->flags & HAV_RENDER_CHECK_RANGE) != 0 && a->range_status == nullptr) return HAV_E_NULL;
  return HAV_OK;
}

static WsLayout layout(const hav_render_args *a) {
  WsLayout L;
  memset(&L, 0, sizeof(L));
  const int64_t total = (int64_t)a->batch * a->rays;
  L.num_blocks = (int)((total + kRaysPerBlock - 1) / kRaysPerBlock);
  const int Sf = a->num_fine > 0 ? (a->num_coarse + 1) / 2 + a->num_fine : 0;
  uint64_t off = 0;
  L.pack_f32 = off, off = align_up(off + (uint64_t)kPackF32Floats * 4, 256);
  if (a->precision != HAV_PREC_FP32) {
    const int x3 = a->precision == HAV_PREC_FP16X3 ? 2 : 1;   // split mode: hi + lo weight images, fp32 channels-last planes
    L.wimg = off, off = align_up(off + x3 * tc_weight_image_bytes(), 256);
    L.planes_cl = off, off = align_up(off + x3 * tc_planes_bytes(2 * a->batch, a->plane_h, a->plane_w), 256);
    const int s2 = tc_scratch_slots(L.num_blocks), s3 = tc3_scratch_slots(L.num_blocks);
    L.scratch_blocks = s2 > s3 ? s2 : s3;
  } else {
    L.scratch_blocks = L.num_blocks;
  }
  if (a->num_fine > 0) {
    L.zbuf = off, off = align_up(off + (uint64_t)L.scratch_blocks * Sf * kRaysPerBlock * 4, 256);
    L.wbuf = off, off = align_up(off + (uint64_t)L.scratch_blocks * a->num_coarse * kRaysPerBlock * 4, 256);
  }
  L.total = off;
  return L;
}

void render_fill_dev(const hav_render_args *a, RenderDev &P) {
  const int64_t total = (int64_t)a->batch * a->rays;
  memset(&P, 0, sizeof(P));
  P.B = a->batch, P.R = a->rays, P.total_rays = (int)total;
  P.Sc = a->num_coarse, P.nfine = a->num_fine;
  P.Sf = a->num_fine > 0 ? (a->num_coarse + 1) / 2 + a->num_fine : 0;
  P.PH = a->plane_h, P.PW = a->plane_w, P.VD = a->vol_d, P.VH = a->vol_h, P.VW = a->vol_w;
  for (int i = 0; i < 3; ++i) {
    P.ps[i] = a->plane_scale[i], P.pt[i] = a->plane_trans[i];
    P.ss[i] = a->skin_scale[i], P.st[i] = a->skin_trans[i];
  }
  P.rays = a->ray_batch, P.bg = a->background, P.invT = a->inv_head_T, P.planes = a->planes, P.wvol = a->wvol;
  P.t_rand = a->t_rand, P.noise_c = a->noise_coarse, P.u_rand = a->u_rand, P.noise_f = a->noise_fine;
  P.rgb_c = a->rgb_coarse, P.depth_c = a->depth_coarse, P.acc_c = a->acc_coarse, P.wmax = a->weights_max;
  P.rgb_f = a->rgb_fine, P.depth_f = a->depth_fine, P.acc_f = a->acc_fine, P.z_fine = a->z_fine;
  P.camera = a->camera, P.pixel_index = a->camera != nullptr ? a->pixel_index : nullptr;
  P.img_h = a->img_h, P.img_w = a->img_w, P.pdf_inds = a->num_fine > 0 ? a->pdf_inds : nullptr;
  P.status = ((a->flags & HAV_RENDER_CHECK_RANGE) != 0 && a->precision == HAV_PREC_FP16) ? a->range_status : nullptr;
}

}  // namespace hav

using namespace hav;

extern "C" uint64_t hav_render_workspace_bytes(const hav_render_args *a) {
  if (render_check_args(a) != HAV_OK) return 0;
  return layout(a).total;
}

extern "C" int hav_render_forward(const hav_render_args *a, void *stream) {
  int rc = render_check_args(a);
  if (rc != HAV_OK) return rc;
  const int64_t total = (int64_t)a->batch * a->rays;
  if (total == 0) return HAV_OK;  // empty ray batch: nothing to write
  WsLayout L = layout(a);
  if (a->workspace == nullptr) return HAV_E_NULL;
  if (a->workspace_bytes < L.total || ((uintptr_t)a->workspace & 255) != 0) return HAV_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t *ws = (uint8_t *)a->workspace;

  RenderDev P;
  render_fill_dev(a, P);
  float *pk = (float *)(ws + L.pack_f32);
  P.W0t = pk + kOffW0t, P.W1t = pk + kOffW1t, P.Wht = pk + kOffWht;
  P.b0 = pk + kOffB0, P.b1 = pk + kOffB1, P.bh = pk + kOffBh, P.Wr = pk + kOffWr, P.br = pk + kOffBr;
  if (a->num_fine > 0) P.zbuf = (float *)(ws + L.zbuf), P.wbuf = (float *)(ws + L.wbuf);

  const bool reuse = (a->flags & HAV_RENDER_REUSE_PACKED) != 0;
  if (!reuse) launch_pack_mlp_fp32(a, pk, st);
  cudaError_t e = cudaSuccess;
  if (a->precision == HAV_PREC_FP32) {
    e = launch_render_fp32(P, L.num_blocks, st);
  } else {
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return HAV_E_ARCH;
    P.wimg = ws + L.wimg;
    P.planes_cl = (const uint16_t *)(ws + L.planes_cl);
    const bool bf16 = a->precision == HAV_PREC_BF16;
    if (a->precision == HAV_PREC_FP16X3) {
      if (!reuse) {
        hav_render_args h = *a;
        h.precision = HAV_PREC_FP16;
        launch_pack_mlp_16(&h, ws + L.wimg, st, nullptr);
        launch_pack_mlp_16_lo(&h, ws + L.wimg + tc_weight_image_bytes(), st);
        e = launch_pack_planes_f32(a->planes, (float *)(ws + L.planes_cl), 2 * a->batch, a->plane_h, a->plane_w, st);
        if (e != cudaSuccess) return (int)e;
      }
      e = launch_render_16_v3(P, L.num_blocks, 2, st);
      return e == cudaSuccess ? HAV_OK : (int)e;
    }
    if ((a->flags & HAV_RENDER_CHECK_RANGE) != 0) {
      e = cudaMemsetAsync(a->range_status, 0, sizeof(int32_t), st);
      if (e != cudaSuccess) return (int)e;
    }
    if (!reuse) {
      launch_pack_mlp_16(a, ws + L.wimg, st, P.status);
      e = launch_pack_planes_16(a->planes, (uint16_t *)(ws + L.planes_cl), 2 * a->batch, a->plane_h, a->plane_w, bf16, st, P.status);
      if (e != cudaSuccess) return (int)e;
    }
    const bool v3 = (a->flags & HAV_RENDER_CTA_PAIRS) != 0 || getenv("HAV_TC3") != nullptr;   // CTA-pair kernel (render_tc3.cu)
    e = v3 ? launch_render_16_v3(P, L.num_blocks, bf16 ? 1 : 0, st) : launch_render_16(P, L.num_blocks, bf16, st);
  }
  return e == cudaSuccess ? HAV_OK : (int)e;
}

extern "C" int hav_abi_version(void) { return HAV_ABI_VERSION; }

extern "C" const char *hav_error_string(int code) {
  switch (code) {
    case HAV_OK: return "ok";
    case HAV_E_NULL: return "a required pointer is NULL";
    case HAV_E_SHAPE: return "unsupported or inconsistent sizes";
    case HAV_E_WORKSPACE: return "workspace too small or misaligned";
    case HAV_E_ARCH: return "device is not sm_100 (B200)";
    case HAV_E_VALUE: return "bad enum/flag value or struct size mismatch";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}
