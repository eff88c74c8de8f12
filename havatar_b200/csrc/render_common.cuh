// Shared device code of the fused render kernels (fp32 CUDA-core kernel and bf16 tcgen05 kernel).
// Each function cites the reference file:line it implements (paths relative to XChenZ/havatar).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/havatar_b200.h"

namespace hav {

constexpr int kPlaneC = 64;              // channels per plane (model/nerf_trainer.py:22)
constexpr int kFeat = 2 * kPlaneC;       // bi-plane feature vector (model/nerf_model.py:99)
constexpr int kFreqs = 8;                // model/nerf_model.py:11,16
constexpr int kPE = kFreqs * 2 * 3;      // 48
constexpr int kIn = kFeat + kPE;         // 176
constexpr int kHid = 128;                // model/nerf_model.py:46-47
constexpr int kRgbFeat = 64;             // fc_rgbFeat out (model/nerf_model.py:50)
constexpr int kOut = 3 + kRgbFeat;       // 67 composited channels
constexpr int kRaysPerBlock = 128;       // one ray per thread / per TMEM lane
constexpr int kMaxSamples = 256;
constexpr int kMaxFine = 64;

// Kernel-side view of hav_render_args (+ derived sizes and packed-weight pointers).
struct RenderDev {
  int B, R, total_rays;
  int Sc, nfine, Sf;
  int PH, PW, VD, VH, VW;
  float ps[3], pt[3], ss[3], st[3];
  const float *rays, *bg, *invT, *planes, *wvol;
  const float *t_rand, *noise_c, *u_rand, *noise_f;
  float *rgb_c, *depth_c, *acc_c, *wmax, *rgb_f, *depth_f, *acc_f, *z_fine;
  // in-kernel ray generation (ABI 2): camera [B,18] = fx fy cx cy | c2w [3,4] | near far; pixel_index [B,R] or NULL
  const float *camera;
  const int32_t *pixel_index;
  int img_h, img_w;
  int32_t *pdf_inds;  // optional [B,R,nfine] searchsorted indices (utils/nerf_util.py:102)
  int32_t *status;    // HAV_RENDER_CHECK_RANGE: fp16 range report, or NULL
  // fp32 path: transposed weights [K][N] + biases (built by pack_mlp_fp32_kernel in the workspace)
  const float *W0t, *W1t, *Wht, *b0, *b1, *bh, *Wr, *br;
  // bf16 path: pre-swizzled UMMA smem image of the weights + channels-last bf16 planes
  const uint8_t *wimg;
  const uint16_t *planes_cl;  // [2,B,H,W,64] bf16
  // per-block scratch in the workspace: zbuf [num_blocks][Sf][128], wbuf [num_blocks][Sc][128]
  float *zbuf, *wbuf;
};

struct Ray {
  float o[3], d[3], near, far, dnorm;
  int b;       // batch element
  bool valid;
};

// dataloader/data_util.py:28-56 get_rays for one pixel: d = normalize(R_c2w . K^-1 [i, j, 1]^T) with
// K = [[fx,0,cx*W],[0,fy,cy*H],[0,0,1]] inverted in closed form.  `cam` = fx fy cx cy | c2w row-major [3,4].
// Used by hav_get_rays' kernel and by the in-kernel ray generation, so both produce the same bits.
template <class F>
__device__ __forceinline__ void pixel_ray(F cam, int W, int H, int pix, float o[3], float d[3]) {
  const float fx = cam(0), fy = cam(1), cx = cam(2) * (float)W, cy = cam(3) * (float)H;
  const float i = (float)(pix % W), j = (float)(pix / W);   // pixel (x = i, y = j): dataloader/dataloader.py:72
  // every operation spelled out (no compiler-chosen contraction): the kernel of hav_get_rays and the render kernels must agree
  float c[3], v[3];
  c[0] = __fmaf_rn(i, __fdiv_rn(1.0f, fx), __fdiv_rn(-cx, fx));
  c[1] = __fmaf_rn(j, __fdiv_rn(1.0f, fy), __fdiv_rn(-cy, fy));
  c[2] = 1.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
    v[a] = __fmaf_rn(cam(4 + a * 4 + 2), c[2], __fmaf_rn(cam(4 + a * 4 + 1), c[1], __fmul_rn(cam(4 + a * 4), c[0])));
  const float nrm = __fsqrt_rn(__fmaf_rn(v[2], v[2], __fmaf_rn(v[1], v[1], __fmul_rn(v[0], v[0]))));
#pragma unroll
  for (int a = 0; a < 3; ++a) d[a] = __fdiv_rn(v[a], nrm), o[a] = cam(4 + a * 4 + 3);
}

__device__ __forceinline__ Ray load_ray(const RenderDev &P, int g) {
  Ray r;
  r.valid = g < P.total_rays;
  int gi = r.valid ? g : 0;
  r.b = gi / P.R;
  if (P.camera != nullptr) {
    const float *cam = P.camera + (size_t)r.b * 18;
    const int pix = P.pixel_index != nullptr ? __ldg(P.pixel_index + gi) : gi - r.b * P.R;
    pixel_ray([&](int k) { return __ldg(cam + k); }, P.img_w, P.img_h, pix, r.o, r.d);
    r.near = __ldg(cam + 16), r.far = __ldg(cam + 17);
  } else {
    const float4 *p = reinterpret_cast<const float4 *>(P.rays + (size_t)gi * 8);
    float4 a = __ldg(p), c = __ldg(p + 1);
    r.o[0] = a.x, r.o[1] = a.y, r.o[2] = a.z;
    r.d[0] = a.w, r.d[1] = c.x, r.d[2] = c.y;
    r.near = c.z, r.far = c.w;
  }
  // utils/nerf_util.py:38  ray_directions[..., None, :].norm(p=2, dim=-1)
  r.dnorm = sqrtf(r.d[0] * r.d[0] + r.d[1] * r.d[1] + r.d[2] * r.d[2]);
  return r;
}

// torch.linspace(0, 1, S)[s] in fp32 (model/nerf_trainer.py:129): ATen fills symmetrically,
// start + step*i in the lower half and end - step*(S-1-i) in the upper half.
__device__ __forceinline__ float linspace01(int s, int S) {
  float step = 1.0f / (float)(S - 1);
  return (s < S / 2) ? step * (float)s : 1.0f - step * (float)(S - 1 - s);
}

// model/nerf_trainer.py:129-139: coarse depth of sample s (stratified jitter when t_rand != NULL).
__device__ __forceinline__ float coarse_z_plain(const Ray &r, int s, int S) {
  float t = linspace01(s, S);
  // separately rounded products, then one add -- what the ATen elementwise kernels produce (no FMA contraction)
  return __fadd_rn(__fmul_rn(r.near, 1.0f - t), __fmul_rn(r.far, t));
}
__device__ __forceinline__ float coarse_z(const RenderDev &P, const Ray &r, int g, int s) {
  float z = coarse_z_plain(r, s, P.Sc);
  if (P.t_rand == nullptr) return z;
  float zl = (s > 0) ? coarse_z_plain(r, s - 1, P.Sc) : z;
  float zu = (s < P.Sc - 1) ? coarse_z_plain(r, s + 1, P.Sc) : z;
  float lower = (s > 0) ? 0.5f * (z + zl) : z;
  float upper = (s < P.Sc - 1) ? 0.5f * (zu + z) : z;
  float tr = __ldg(P.t_rand + (size_t)g * P.Sc + s);
  return __fadd_rn(lower, __fmul_rn(upper - lower, tr));
}

// align_corners=True un-normalisation of ATen grid_sampler: ((x + 1) / 2) * (size - 1)
__device__ __forceinline__ float unnorm(float c, int size) { return ((c + 1.0f) * 0.5f) * (float)(size - 1); }

// utils/util.py:409-418 voxel_feature: F.grid_sample 5-D, trilinear, padding_mode='border', align_corners=True.
// Corner order and weight products follow ATen's grid_sampler_3d (tnw,tne,tsw,tse,bnw,bne,bsw,bse).
__device__ __forceinline__ float trilinear_border(const float *__restrict__ vol, int D, int H, int W, float x,
                                                  float y, float z) {
  float ix = fminf(fmaxf(unnorm(x, W), 0.0f), (float)(W - 1));
  float iy = fminf(fmaxf(unnorm(y, H), 0.0f), (float)(H - 1));
  float iz = fminf(fmaxf(unnorm(z, D), 0.0f), (float)(D - 1));
  float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
  int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
  float wx1 = ix - x0f, wy1 = iy - y0f, wz1 = iz - z0f;
  float wx0 = (x0f + 1.0f) - ix, wy0 = (y0f + 1.0f) - iy, wz0 = (z0f + 1.0f) - iz;
  bool xe = x0 + 1 <= W - 1, ye = y0 + 1 <= H - 1, ze = z0 + 1 <= D - 1;
  int x1 = xe ? x0 + 1 : x0, y1 = ye ? y0 + 1 : y0, z1 = ze ? z0 + 1 : z0;
  const float *p0 = vol + ((size_t)z0 * H) * W, *p1 = vol + ((size_t)z1 * H) * W;
  float v000 = __ldg(p0 + y0 * W + x0), v001 = __ldg(p0 + y0 * W + x1);
  float v010 = __ldg(p0 + y1 * W + x0), v011 = __ldg(p0 + y1 * W + x1);
  float v100 = __ldg(p1 + y0 * W + x0), v101 = __ldg(p1 + y0 * W + x1);
  float v110 = __ldg(p1 + y1 * W + x0), v111 = __ldg(p1 + y1 * W + x1);
  float out = v000 * ((wx0 * wy0) * wz0);
  if (xe) out += v001 * ((wx1 * wy0) * wz0);
  if (ye) out += v010 * ((wx0 * wy1) * wz0);
  if (xe && ye) out += v011 * ((wx1 * wy1) * wz0);
  if (ze) out += v100 * ((wx0 * wy0) * wz1);
  if (ze && xe) out += v101 * ((wx1 * wy0) * wz1);
  if (ze && ye) out += v110 * ((wx0 * wy1) * wz1);
  if (ze && xe && ye) out += v111 * ((wx1 * wy1) * wz1);
  return out;
}

// model/Skinning_Field.py:70-98 Deformation_Field_new.forward for one point.  Bone 0 = identity
// (Skinning_Field.py:50), bone 1 = inv_head_T[b] ([4,3]: p1 = (p + T[3]) @ T[:3,:3], :83).
__device__ __forceinline__ void skin_warp(const RenderDev &P, const float *__restrict__ T, const float p[3],
                                          float pc[3]) {
  float q0 = p[0] + T[9], q1 = p[1] + T[10], q2 = p[2] + T[11];
  float p1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) p1[j] = q0 * T[j] + q1 * T[3 + j] + q2 * T[6 + j];
  const size_t vs = (size_t)P.VD * P.VH * P.VW;
  float w0 = trilinear_border(P.wvol, P.VD, P.VH, P.VW, p[0] * P.ss[0] + P.st[0], p[1] * P.ss[1] + P.st[1],
                              p[2] * P.ss[2] + P.st[2]);                                 // :85, bone 0
  float w1 = trilinear_border(P.wvol + vs, P.VD, P.VH, P.VW, p1[0] * P.ss[0] + P.st[0],
                              p1[1] * P.ss[1] + P.st[1], p1[2] * P.ss[2] + P.st[2]);     // :85, bone 1
  float den = (w0 + w1) + 1e-8f;                                                         // :87
  float a0 = w0 / den, a1 = w1 / den;
#pragma unroll
  for (int j = 0; j < 3; ++j) pc[j] = a0 * p[j] + a1 * p1[j];                            // :90,95
}

// One bilinear tap set of F.grid_sample(4-D, bilinear, zeros, align_corners=True) (utils/util.py:395-406).
// off[] are texel indices (y*W+x) clamped in range; w[] are the ATen corner weights, 0 when the corner is
// outside (zeros padding).  Corner order nw, ne, sw, se.
struct Taps {
  int off[4];
  float w[4];
};
__device__ __forceinline__ Taps bilinear_taps(float gx, float gy, int H, int W) {
  float ix = unnorm(gx, W), iy = unnorm(gy, H);
  float x0f = floorf(ix), y0f = floorf(iy);
  float wx1 = ix - x0f, wy1 = iy - y0f;
  float wx0 = (x0f + 1.0f) - ix, wy0 = (y0f + 1.0f) - iy;
  // clamp before the int conversion: far-away points only produce masked taps
  int x0 = (int)fminf(fmaxf(x0f, -2.0f), (float)(W + 1));
  int y0 = (int)fminf(fmaxf(y0f, -2.0f), (float)(H + 1));
  int x1 = x0 + 1, y1 = y0 + 1;
  bool vx0 = x0 >= 0 && x0 <= W - 1, vx1 = x1 >= 0 && x1 <= W - 1;
  bool vy0 = y0 >= 0 && y0 <= H - 1, vy1 = y1 >= 0 && y1 <= H - 1;
  int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
  int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  Taps t;
  t.off[0] = cy0 * W + cx0, t.w[0] = (vx0 && vy0) ? wx0 * wy0 : 0.0f;
  t.off[1] = cy0 * W + cx1, t.w[1] = (vx1 && vy0) ? wx1 * wy0 : 0.0f;
  t.off[2] = cy1 * W + cx0, t.w[2] = (vx0 && vy1) ? wx0 * wy1 : 0.0f;
  t.off[3] = cy1 * W + cx1, t.w[3] = (vx1 && vy1) ? wx1 * wy1 : 0.0f;
  return t;
}

// Running state of utils/nerf_util.py:28-73 volume_render_radiance_field for one ray.
struct Composite {
  float T, acc, depth, wmax;
  __device__ __forceinline__ void reset() { T = 1.0f, acc = 0.0f, depth = 0.0f, wmax = 0.0f; }
  // returns the sample weight w_i = alpha_i * T_i  (:59-60; cumprod_exclusive :4-25)
  template <bool kFast = false>
  __device__ __forceinline__ float step(float alpha_raw, float noise, float dist, float z) {
    float sigma = fmaxf(alpha_raw + noise, 0.0f);   // :58
    float alpha = 1.0f - (kFast ? __expf(-sigma * dist) : expf(-sigma * dist));       // :59
    float w = alpha * T;
    T *= (1.0f - alpha) + 1e-10f;
    acc += w;                                       // :67
    depth += w * z;                                 // :64-65
    wmax = fmaxf(wmax, w);                          // model/nerf_trainer.py:195
    return w;
  }
};

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.0f / (1.0f + expf(-x)); }  // nerf_util.py:45
// MUFU.EX2 + MUFU.RCP version for the 16-bit tensor-core modes (rel. error ~1e-6, far below the operand rounding)
__device__ __forceinline__ float sigmoidf_fast(float x) { return __frcp_rn(1.0f + __expf(-x)); }

// utils/nerf_util.py:76-117 sample_pdf for one ray (one thread).
//   zmid(j)    : bin edge j (z_vals_mid, j = 0..Sc-2)
//   wcol[j*ld] : in  = coarse weights w_j (j = 0..Sc-1; w_0 and w_{Sc-1} are not used, :79), used as cdf scratch (overwritten)
//   zs[k]      : out = the nfine samples (unsorted order of u);  inds_out[k] = searchsorted index (optional)
template <class MidFn>
__device__ __forceinline__ void sample_pdf_core(MidFn zmid, int Sc, int nfine, float *wcol, int ld, const float *u_rand,
                                                float *zs, int32_t *inds_out) {
  const int M = Sc - 1;  // number of bins edges (z_mid) == cdf entries
  // pdf over weights[1:-1] + 1e-5 (:79-80)
  float sum = 0.0f;
  for (int j = 1; j <= Sc - 2; ++j) sum += wcol[j * ld] + 1e-5f;
  // cdf[0] = 0, cdf[j] = cdf[j-1] + pdf[j]  (:81-84); stored at wcol[j], j = 0..M-1
  float run = 0.0f;
  wcol[0] = 0.0f;
  for (int j = 1; j <= Sc - 2; ++j) {
    run += (wcol[j * ld] + 1e-5f) / sum;
    wcol[j * ld] = run;
  }
  for (int k = 0; k < nfine; ++k) {
    float u;
    if (u_rand == nullptr) {
      // torch.linspace(0, 1, nfine) (:87-91)
      u = (nfine == 1) ? 0.0f : linspace01(k, nfine);
    } else {
      // (arange(n) * s) + rand * (s - 1e-6) (:93-95); s is a python double there
      u = __fadd_rn(__fmul_rn((float)k, (float)(1.0 / (double)nfine)), __fmul_rn(u_rand[k], (float)(1.0 / (double)nfine - 1e-6)));
    }
    // inds = searchsorted(cdf, u, right=True): first index with cdf[i] > u, M if none (:102)
    int lo = 0, hi = M;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (wcol[mid * ld] > u) hi = mid; else lo = mid + 1;
    }
    if (inds_out != nullptr) inds_out[k] = lo;
    int below = max(lo - 1, 0), above = min(lo, M - 1);                              // :103-104
    float c0 = wcol[below * ld], c1 = wcol[above * ld];
    float b0 = zmid(below), b1 = zmid(above);                                        // z_vals_mid
    float denom = c1 - c0;
    denom = (denom < 1e-5f) ? 1.0f : denom;                                          // :112-113
    float t = (u - c0) / denom;
    zs[k] = __fadd_rn(b0, __fmul_rn(t, b1 - b0));                                                      // :114-115
  }
}

// sample_pdf + model/nerf_trainer.py:166-170 merge, for one ray (one thread).
//   zc(s)      : coarse depth of sample s (0..Sc-1)
//   zout[j*ld] : out = sorted(cat(z[::2], z_samples)), Sf = (Sc+1)/2 + nfine entries
template <class ZFn>
__device__ __forceinline__ void sample_pdf_merge(ZFn zc, int Sc, int nfine, float *wcol, int ld, const float *u_rand,
                                                 float *zout, int32_t *inds_out = nullptr) {
  float zs[kMaxFine];
  sample_pdf_core([&](int j) { return 0.5f * (zc(j + 1) + zc(j)); }, Sc, nfine, wcol, ld, u_rand, zs, inds_out);
  // torch.sort semantics even if rounding ever produced an inversion: insertion sort (normally a no-op)
  for (int k = 1; k < nfine; ++k) {
    float v = zs[k];
    int j = k - 1;
    while (j >= 0 && zs[j] > v) { zs[j + 1] = zs[j]; --j; }
    zs[j + 1] = v;
  }
  // merge with z_vals[:, ::2] (model/nerf_trainer.py:170)
  const int nh = (Sc + 1) / 2;
  int a = 0, bq = 0;
  for (int j = 0; j < nh + nfine; ++j) {
    float za = (a < nh) ? zc(2 * a) : 0.0f;
    bool take_a = (a < nh) && (bq >= nfine || za <= zs[bq]);
    zout[j * ld] = take_a ? za : zs[bq];
    if (take_a) ++a; else ++bq;
  }
}

}  // namespace hav
