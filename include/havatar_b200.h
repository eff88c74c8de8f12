/*
 * havatar_b200 -- C ABI of the B200-native (sm_100a) HAvatar render hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 *   - takes DEVICE pointers to contiguous float32 tensors in the reference's own layouts,
 *   - never allocates (the caller passes outputs and, where needed, a workspace),
 *   - launches asynchronously on the CUDA stream passed as `stream` (a cudaStream_t cast to void*;
 *     NULL = legacy default stream) on the current device, and is re-entrant,
 *   - returns 0 on success, a negative HAV_E_* code for an argument error, or a positive
 *     cudaError_t when the launch failed.  No exception crosses the boundary.
 *
 * Reference interfaces replaced (paths relative to the XChenZ/havatar tree):
 *   hav_fused_bias_act   <- model/op/fused_bias_act.cpp:18-32   (pybind module `fused`)
 *   hav_upfirdn2d        <- model/op/upfirdn2d.cpp:17-31        (pybind module `upfirdn2d`)
 *   hav_render_forward   <- model/nerf_trainer.py:120-201       (Trainer.predict_and_render_radiance; the
 *                            reference has no native boundary here -- it is ~140 ATen launches per chunk)
 *   hav_render_backward  <- loss.backward() through model/nerf_trainer.py:120-201 (train_avatar.py:149; ATen autograd)
 *   hav_get_rays         <- dataloader/data_util.py:28-56 + dataloader/dataloader.py:174-180
 *   hav_make_render_cond <- dataloader/dataloader.py:218-229 (make_render_cond_)
 *   hav_sample_pdf       <- utils/nerf_util.py:76-117 (sample_pdf)
 *   hav_conv2d_forward   <- model/styleUnet.py:222-297 (ModulatedConv2d.forward) and :108-118 (EqualConv2d.forward): the
 *                            reference calls cuDNN grouped conv2d / conv_transpose2d through model/op/conv2d_gradfix.py:22-75
 *   hav_pack_planes      <- model/nerf_model.py:85 (plane stacking; layout change for the bf16 path)
 *   hav_conv2d_wgrad     <- autograd's convolution_backward under the same modules (model/op/conv2d_gradfix.py:94-227)
 *   hav_adam_flat        <- torch.optim.Adam.step() in the training loops (train_avatar.py:151, train_avatarHD.py:231,279-280)
 * INTEGRATION.md shows the reference-side binding for each.
 */
#ifndef HAVATAR_B200_H_
#define HAVATAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HAV_ABI_VERSION 3

/* argument errors (negative); positive return values are cudaError_t */
#define HAV_OK 0
#define HAV_E_NULL (-1)       /* a required pointer is NULL */
#define HAV_E_SHAPE (-2)      /* unsupported or inconsistent sizes */
#define HAV_E_WORKSPACE (-3)  /* workspace too small (see hav_render_workspace_bytes) */
#define HAV_E_ARCH (-4)       /* device is not sm_100 */
#define HAV_E_VALUE (-5)      /* bad enum / flag value */

/* arithmetic of the MLP inside hav_render_forward */
#define HAV_PREC_FP32 0 /* CUDA-core fp32 everywhere: reference-exact mode (1e-5 class parity) */
#define HAV_PREC_BF16 1 /* tcgen05 bf16 operands, fp32 accumulate in TMEM (fast path, fp32 exponent range) */
#define HAV_PREC_FP16 2 /* tcgen05 fp16 operands (saturating converts), fp32 accumulate: fast path, 8x finer rounding */
#define HAV_PREC_FP16X3 3 /* tcgen05 split precision: activations and weights as fp16 hi + lo pairs, three MMAs per product
                             (hi*hi + lo*hi + hi*lo, fp32 accumulate), fp32 planes / blends / positional encoding / composite:
                             fp32-class results (1e-5 class parity) at tensor-core speed.  Operand range is fp16's. */

/* hav_render_args.flags */
#define HAV_RENDER_REUSE_PACKED 1 /* the workspace still holds the packed MLP weights and planes written by a previous
                                     hav_render_forward with the same weights/planes/precision/batch: skip re-packing
                                     (weights change once per optimiser step, planes once per frame; the reference
                                     re-renders the same frame in 4096-ray groups, train_avatar.py:182-218) */
#define HAV_RENDER_CHECK_RANGE 2  /* HAV_PREC_FP16 only: report operands that leave the fp16 range in *range_status (below).
                                     fp16 conversions saturate, so without this an out-of-range model renders finite but
                                     wrong values; callers switch to HAV_PREC_BF16 when the status comes back non-zero */

#define HAV_RENDER_CTA_PAIRS 4    /* 16-bit modes: run the CTA-pair kernel (tcgen05 cta_group::2, render_tc3.cu) instead of the
                                     single-CTA kernel (render_tc2.cu); HAV_PREC_FP16X3 always runs on CTA pairs */

int hav_abi_version(void);
const char *hav_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * fused bias + activation.  Replaces fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 * (model/op/fused_bias_act.cpp:18-32, kernel model/op/fused_bias_act_kernel.cu:18-65).
 *   y[i] = f(x[i] + bias[(i / step_b) % size_b]) * scale,   selected by act*10+grad:
 *     30: leaky-relu(alpha) fwd   31: its gradient gated by sign(ref[i])   32: 0
 *     10/11: linear               12: 0
 *   bias == NULL <=> "empty bias tensor"; ref == NULL <=> "empty refer tensor" (kernel.cu:79-80).
 *   step_b = product of dims after the channel dim (kernel.cu:86-88), size_b = channels.
 */
int hav_fused_bias_act(float *out, const float *x, const float *bias, const float *ref, int64_t numel,
                       int64_t step_b, int64_t size_b, int act, int grad, float alpha, float scale,
                       void *stream);

/* ------------------------------------------------------------------------------------------------
 * upfirdn2d.  Replaces upfirdn2d.upfirdn2d(input[major,in_h,in_w,minor], kernel[kh,kw], up_x, up_y,
 * down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1) (model/op/upfirdn2d.cpp:17-31, kernels
 * model/op/upfirdn2d_kernel.cu:49-207): zero-insert upsample, pad (negative = crop), correlate with
 * the FLIPPED kernel, decimate.  out is [major, out_h, out_w, minor] with
 *   out_h = (in_h*up_y + pad_y0 + pad_y1 - kh + down_y) / down_y   (kernel.cu:236-241), same for w.
 */
int hav_upfirdn2d(float *out, const float *x, const float *kernel, int major, int in_h, int in_w,
                  int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0,
                  int pad_x1, int pad_y0, int pad_y1, void *stream);

/* FusedLeakyReLUFunctionBackward in one pass (model/op/fused_act.py:23-47): for x [batch, channels, inner]
 *   grad_input = (ref > 0 ? grad_out : grad_out * alpha) * scale        (== hav_fused_bias_act act 3, grad 1)
 *   partials[s, c] = sum of grad_input over split s of channel c's batch * inner elements;   grad_bias[c] = sum_s partials[s, c]
 * (fixed summation order; replaces the separate full-tensor reduction).  splits = hav_bias_act_backward_splits(...) <= 64. */
int hav_bias_act_backward_splits(int batch, int channels, int64_t inner);
int hav_bias_act_backward(float *grad_input, float *partials, const float *grad_out, const float *ref, int batch, int channels,
                          int64_t inner, int splits, float alpha, float scale, void *stream);

/* StyledConv's tail in one pass (model/styleUnet.py:596-598 = NoiseInjection :300-310 followed by FusedLeakyReLU): for x [batch,
 * channels, inner] and noise [batch or 1, 1, inner],  out = lrelu(x + *noise_weight * noise + bias[c], alpha) * scale.  noise_weight
 * points to DEVICE memory (it is a parameter).  Backward: grad_input and partials as hav_bias_act_backward, plus
 * noise_partials[s, c] = sum(grad_input * noise) over the split: grad of the noise weight = sum over (s, c). */
int hav_noise_bias_act(float *out, const float *x, const float *bias, const float *noise, const float *noise_weight, int batch, int channels,
                       int64_t inner, int noise_per_sample, float alpha, float scale, void *stream);
int hav_noise_bias_act_backward(float *grad_input, float *partials, float *noise_partials, const float *grad_out, const float *ref,
                                const float *noise, int batch, int channels, int64_t inner, int noise_per_sample, int splits, float alpha,
                                float scale, void *stream);

/* upfirdn2d on channels-last fp16 tensors (HAV_LAYOUT_NHWC_F16: x [B,H,W,C] -> out [B,Ho,Wo,C], C % 8 == 0), square up / down
 * factors, with the tail of StyledConv fused in (model/styleUnet.py:593-599 after the blur of :264-277):
 *   out = act( fir(x) + noise_weight * noise + bias[c] ),  act 0: none, 1: leaky-relu(0.2) * sqrt(2); noise / bias may be NULL. */
int hav_upfirdn2d_cl(void *out, const void *x, const float *kernel, int batch, int in_h, int in_w, int channels, int kh, int kw,
                     int up, int down, int pad_x0, int pad_x1, int pad_y0, int pad_y1, const float *noise, float noise_weight,
                     int noise_per_sample, const float *bias, int act, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused volumetric render: per-ray depth sampling -> 2-bone skinning warp -> bi-plane bilinear fetch
 * -> positional encoding -> 5-linear MLP -> alpha composite (-> sample_pdf -> second pass).
 * One call == Trainer.predict_and_render_radiance over ALL B*R rays (model/nerf_trainer.py:120-201);
 * the reference's 4096-ray python chunk loop (nerf_trainer.py:65-71) is not needed because no
 * per-sample tensor ever reaches HBM.
 */
typedef struct hav_render_args {
  uint32_t struct_bytes; /* = sizeof(hav_render_args): ABI guard */
  int32_t precision;     /* HAV_PREC_* */
  int32_t batch;         /* B */
  int32_t rays;          /* R rays per batch element */
  int32_t num_coarse;    /* S_c: nerf.<mode>.num_coarse (2..256) */
  int32_t num_fine;      /* nerf.<mode>.num_fine; 0 = coarse only; fine pass has (S_c+1)/2 + num_fine samples */
  int32_t plane_c;       /* feature channels per plane (64) */
  int32_t plane_h, plane_w;
  int32_t vol_d, vol_h, vol_w;
  int32_t flags;         /* HAV_RENDER_* bits, 0 by default */
  float plane_scale[3], plane_trans[3]; /* model_coarse.gridwarper (utils/util.py:214-236) */
  float skin_scale[3], skin_trans[3];   /* headpose_skin_net.gridwarper (model/nerf_trainer.py:29-34) */

  /* inputs, float32, reference layouts */
  const float *ray_batch;   /* [B,R,8]  o3 d3 near far   (dataloader/dataloader.py:179-180) */
  const float *background;  /* [B,R,3]  or NULL          (utils/nerf_util.py:70-71) */
  const float *inv_head_T;  /* [B,4,3]  rows 0-2 R^-1, row 3 -t (dataloader/dataloader.py:215-216) */
  const float *planes;      /* [2,B,C,H,W] NCHW          (model/nerf_model.py:85) */
  const float *wvol;        /* [1,2,D,H,W] skinning weights (model/Skinning_Field.py:79) */
  const float *w0, *b0;     /* layers_xyz.0  [128,176],[128]  (model/nerf_model.py:46) */
  const float *w1, *b1;     /* layers_xyz.1  [128,128],[128] */
  const float *w_alpha, *b_alpha; /* fc_alpha   [1,128],[1] */
  const float *w_feat, *b_feat;   /* fc_rgbFeat [64,128],[64] */
  const float *w_rgb, *b_rgb;     /* fc_rgb     [3,64],[3] */

  /* the reference's random draws as explicit inputs (all optional; NULL = that randomness is off) */
  const float *t_rand;       /* [B,R,S_c] U[0,1): stratified jitter (model/nerf_trainer.py:132-139) */
  const float *noise_coarse; /* [B,R,S_c] N(0,1)*std: sigma noise (utils/nerf_util.py:47-57) */
  const float *u_rand;       /* [B,R,num_fine] U[0,1): sample_pdf jitter (utils/nerf_util.py:93-96); NULL = det */
  const float *noise_fine;   /* [B,R,S_f] */

  /* outputs, float32 (fine outputs may be NULL when num_fine == 0) */
  float *rgb_coarse;   /* [B,R,67]  rgb3 | feature64 */
  float *depth_coarse; /* [B,R] */
  float *acc_coarse;   /* [B,R] */
  float *weights_max;  /* [B,R]  max_s w of the LAST pass (model/nerf_trainer.py:195,200) */
  float *rgb_fine, *depth_fine, *acc_fine;
  float *z_fine;       /* optional [B,R,S_f]: merged+sorted fine depths (debug / tests), or NULL */

  void *workspace;          /* >= hav_render_workspace_bytes(args) bytes, 256-byte aligned */
  uint64_t workspace_bytes;

  /* ---- ABI 2 ---- */
  /* In-kernel ray generation (dataloader/data_util.py:28-56 get_rays + the near / far fill of dataloader/dataloader.py:174-180):
   * when `camera` is non-NULL, ray_batch may be NULL and the rays of batch element b are generated inside the render kernel
   * from camera[b] = { fx, fy, cx, cy | c2w row-major [3,4] | near, far } (18 floats, DEVICE memory; focal lengths in pixels,
   * principal point as a fraction of the image size) for an img_h x img_w image: ray r <-> pixel (p / img_w, p % img_w) with
   * p = pixel_index[b*R + r], or p = r when pixel_index is NULL (dataloader.py:72; then R must equal img_h * img_w).
   * Same arithmetic as hav_get_rays, so a render from `camera` equals the render of hav_get_rays' output bit for bit. */
  const float *camera;        /* [B,18] or NULL */
  const int32_t *pixel_index; /* [B,R] or NULL (the dataloader's select_inds as y * img_w + x, dataloader.py:160-170) */
  int32_t img_h, img_w;
  int32_t *pdf_inds;          /* optional [B,R,num_fine]: sample_pdf's searchsorted(cdf, u, right=True) indices
                                 (utils/nerf_util.py:102) -- integer bookkeeping exposed for bit-exact parity tests, or NULL */
  int32_t *range_status;      /* DEVICE int32, required with HAV_RENDER_CHECK_RANGE: zeroed by the call, then bit 0 = a plane texel
                                 or MLP weight exceeds the fp16 range, bit 1 = a hidden activation saturated at +-65504 */
} hav_render_args;

uint64_t hav_render_workspace_bytes(const hav_render_args *args);
int hav_render_forward(const hav_render_args *args, void *stream);

/* sample_pdf alone (utils/nerf_util.py:76-117): the same device function the render kernels run between their two passes,
 * one thread per row -- exists so that the integer bookkeeping (searchsorted indices, :102) can be compared bit for bit on
 * identical inputs.  bins [n,m] (z_vals_mid), weights [n,m-1], u [n,nfine] uniform draws or NULL (det=True, :87-91);
 * samples [n,nfine] (unsorted, as sample_pdf returns them), inds [n,nfine] int32 or NULL; scratch: n*(m+1) floats. */
int hav_sample_pdf(const float *bins, const float *weights, const float *u, int n, int m, int nfine, float *samples,
                   int32_t *inds, float *scratch, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Backward of hav_render_forward: what autograd produces for Trainer.predict_and_render_radiance in the reference
 * (model/nerf_trainer.py:120-201 under loss.backward(), train_avatar.py:149) -- gradients with respect to the bi-plane
 * features, the skinning-weight volume and the ten MLP tensors.  `fwd` is the argument block of the forward call being
 * differentiated: same inputs and random draws, its OUTPUT buffers still holding that call's results (rgb / depth / acc of
 * each pass are read; z_fine is required when num_fine > 0: the resampled depths are constants of the backward because the
 * reference detaches them, nerf_trainer.py:167), precision HAV_PREC_FP16 or HAV_PREC_BF16.  fwd->workspace is not used.
 * weights_max has no gradient path (the reference never differentiates it, train_avatar.py:131-146).
 * Upstream gradients may be NULL (= zero).  All outputs are OVERWRITTEN (not accumulated), float32, reference layouts.
 * grad_scale: power-of-two loss scale for the 16-bit gradient operands; <= 0 = chosen on the device from max|upstream|.
 */
typedef struct hav_render_bwd_args {
  uint32_t struct_bytes; /* = sizeof(hav_render_bwd_args) */
  float grad_scale;
  const hav_render_args *fwd;
  const float *g_rgb_coarse;   /* [B,R,67] */
  const float *g_depth_coarse; /* [B,R] */
  const float *g_acc_coarse;   /* [B,R] */
  const float *g_rgb_fine, *g_depth_fine, *g_acc_fine;
  float *g_planes;             /* [2,B,C,H,W] */
  float *g_wvol;               /* [1,2,D,H,W] */
  float *g_w0, *g_b0, *g_w1, *g_b1, *g_w_alpha, *g_b_alpha, *g_w_feat, *g_b_feat, *g_w_rgb, *g_b_rgb;
  void *workspace;             /* >= hav_render_backward_workspace_bytes(args) bytes, 256-byte aligned */
  uint64_t workspace_bytes;
} hav_render_bwd_args;

uint64_t hav_render_backward_workspace_bytes(const hav_render_bwd_args *args);
int hav_render_backward(const hav_render_bwd_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Ray generation (dataloader/data_util.py:28-56 get_rays + near/far of dataloader/dataloader.py:174-180).
 * intr = {fx, fy, cx, cy} (focal in pixels, principal point as a fraction of the image size);
 * c2w = row-major [3,4].  Writes ray_batch [H*W, 8] = o3 d3 near far, ray r <-> pixel (r / W, r % W).
 */
int hav_get_rays(float *ray_batch, int height, int width, const float intr[4], const float c2w[12],
                 float near, float far, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Condition renderings (dataloader/dataloader.py:218-229 make_render_cond_): render / normal are the decoded
 * ortho_{front,left,right}_{render,normal}_256_baseGama.png pairs as [n, pixels, 3] uint8 RGB in DEVICE memory; out is
 * [n, 7, pixels] float32 = render / 255 | normal / 255 | (normal != 0) -- channels first, the layout the plane generators take
 * (train_avatar.py:121-123).  Uploading uint8 and converting here moves 7x fewer bytes than the reference's float32 [H,W,7].
 */
int hav_make_render_cond(float *out, const uint8_t *render, const uint8_t *normal, int n, int pixels, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core convolution of the StyleUNet blocks: ModulatedConv2d (model/styleUnet.py:165-297) in the shared-weight
 * formulation of its non-fused branch (:225-251), EqualConv2d (:88-123), with the surrounding per-layer elementwise work
 * fused in: out = act( conv(x * in_scale[b,ci], W) * out_scale[b,co] + noise_weight * noise + bias[co] ).
 *   ksize 1 or 3;  up = 2: conv_transpose2d stride 2, padding 0 (:264-270), output (2H+1) x (2W+1) for ksize 3;
 *   down = 2: conv2d stride 2, padding 0 (:281-283);  otherwise stride 1, padding ksize/2 (:289-291).
 *   x [B,Cin,H,W], out [B,Cout,Ho,Wo] float32 NCHW; operands are rounded to 16 bit, accumulation is fp32 (TMEM).
 * Weights are packed once per weight update with hav_conv_pack_weights for the SAME `up` they will be used with
 * (w is [Cout,Cin,k,k], or [Cin,Cout,k,k] -- conv_transpose2d's own layout -- when bit 0 of transpose_io is set; bit 1
 * (HAV_CONV_PACK_FLIP) mirrors the taps, w[.., k-1-kh, k-1-kw]: with both bits set the packed image is the weight of the
 * DATA-GRADIENT convolution of a stride-1 layer, so hav_conv2d_forward also serves as that layer's backward-data kernel).
 */
#define HAV_CONV_PACK_TRANSPOSE_IO 1
#define HAV_CONV_PACK_FLIP 2
#define HAV_LAYOUT_NCHW_F32 0 /* [B,C,H,W] float32 */
#define HAV_LAYOUT_NHWC_F16 1 /* [B,H,W,C] IEEE half, C % 8 == 0 (fp16 precision only) */

typedef struct hav_conv_args {
  uint32_t struct_bytes; /* = sizeof(hav_conv_args) */
  int32_t precision;     /* HAV_PREC_FP16 or HAV_PREC_BF16: must match the packed weights */
  int32_t batch, cin, cout, in_h, in_w;
  int32_t ksize, up, down;
  int32_t act;              /* 0: none, 1: leaky-relu(0.2) * sqrt(2) (model/op/fused_act.py:103-122) */
  int32_t noise_per_sample; /* 0: noise is [1,1,Ho,Wo] broadcast over the batch, 1: [B,1,Ho,Wo] */
  float noise_weight;       /* NoiseInjection.weight (model/styleUnet.py:300-310) */
  int32_t in_layout;        /* HAV_LAYOUT_NCHW_F32 (reference layout) or HAV_LAYOUT_NHWC_F16 (internal hand-over between layers) */
  int32_t out_layout;
  const void *x;
  const void *wpack;        /* hav_conv_wpack_bytes(cout, cin, ksize, up) bytes written by hav_conv_pack_weights */
  const float *in_scale;    /* [B,Cin] modulation s, or NULL */
  const float *out_scale;   /* [B,Cout] demodulation, or NULL */
  const float *noise;       /* or NULL */
  const float *bias;        /* [Cout] or NULL */
  void *out;
  const void *residual;     /* NULL, or a tensor of out's shape and layout added AFTER the activation: out = act(...) + residual
                               (FromRGB's `out + skip`, model/styleUnet.py:464-465; ToRGB's, :625-626) */
} hav_conv_args;

uint64_t hav_conv_wpack_bytes(int cout, int cin, int ksize, int up);
int hav_conv_pack_weights(void *wpack, const float *w, int cout, int cin, int ksize, float scale, int up, int transpose_io,
                          int precision, void *stream);
/* demod[b,co] = rsqrt(sum_{ci,kh,kw} (scale * w[co,ci,kh,kw] * style[b,ci])^2 + eps)   (model/styleUnet.py:256-258) */
int hav_modconv_demod(float *demod, const float *w, const float *style, int batch, int cout, int cin, int ksize, float scale,
                      float eps, void *stream);
int hav_conv2d_forward(const hav_conv_args *args, void *stream);

/* Modulation vectors and demodulation factors of every modulated convolution of one network in two launches (the per-layer
 * form above costs one small linear + one reduction per layer on the critical path of a frame; model/styleUnet.py:237-258).
 * `layers` is a DEVICE array of n_layers descriptors; layer l computes
 *   s[b,ci]   = (sum_d mod_w[ci,d] * latent[b, latent_index, d]) * mod_scale + mod_b[ci] * mod_lr_mul      (EqualLinear, :126-162)
 *   d[b,co]   = rsqrt(conv_scale^2 * sum_ci s[b,ci]^2 * wsq[co,ci] + eps)    when wsq != NULL            (demodulate=True)
 * into s_all + s_off ([batch,cin] block) and d_all + d_off ([batch,cout] block).  wsq[co,ci] = sum over taps of W^2
 * (hav_conv_tap_squares; input independent, cache it per weight version).  s_prefix / d_prefix: DEVICE int arrays of
 * n_layers + 1 running sums of cin / of cout-or-0 (layers without demodulation contribute 0 rows).  batch <= 8. */
typedef struct hav_style_layer {
  const float *mod_w;  /* [cin, style_dim] EqualLinear weight (unscaled parameter) */
  const float *mod_b;  /* [cin] EqualLinear bias or NULL */
  const float *wsq;    /* [cout, cin] tap-summed squares of the convolution weight, or NULL (no demodulation) */
  int32_t cin, cout, latent_index, s_off, d_off;
  float mod_scale, mod_lr_mul, conv_scale;
} hav_style_layer;
int hav_conv_tap_squares(float *wsq, const float *w, int cout, int cin, int ksize, void *stream);
int hav_style_plan_run(float *s_all, float *d_all, const float *latent, int batch, int n_latent, int style_dim,
                       const hav_style_layer *layers, const int *s_prefix, const int *d_prefix, int n_layers, int s_rows, int d_rows,
                       float eps, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Backward of hav_conv2d_forward (replaces cuDNN's convolution_backward under model/op/conv2d_gradfix.py:94-227 for the
 * training steps train_avatar.py:149 / train_avatarHD.py:229,276).  For  y = out_scale * conv(in_scale * x, wscale * w):
 *   data gradient    = hav_conv2d_forward on g with the transposed (and, for stride 1, flipped) weight image, in_scale :=
 *                      out_scale, and the roles of up / down exchanged (the gradient of a stride-2 convolution is a
 *                      transposed one and vice versa);
 *   weight gradient  = hav_conv2d_wgrad:  dw[co,ci,kh,kw] (+)= wscale * sum_{b,p} (out_scale[b,co] g[b,co,p]) *
 *                      (in_scale[b,ci] x[b,ci,p + (kh,kw) - pad]),  tcgen05 bf16 operands, fp32 accumulate, split over the
 *                      positions with fp32 reductions into dw ([Cout,Cin,k,k] float32; zeroed first unless accumulate != 0);
 *   scale gradients  = hav_rowscale_dot (rows = (sample, channel) images): out = a * scale[row], dot[row] = sum a * x.
 * x [B,Cin,H,W] is the forward INPUT, g [B,Cout,Ho,Wo] the gradient of the forward output (sizes as in hav_conv2d_forward).
 */
typedef struct hav_conv_wgrad_args {
  uint32_t struct_bytes; /* = sizeof(hav_conv_wgrad_args) */
  int32_t batch, cin, cout, in_h, in_w;
  int32_t ksize, up, down;
  int32_t accumulate;
  float wscale;
  const float *g;
  const float *x;
  const float *in_scale;  /* [B,Cin] or NULL */
  const float *out_scale; /* [B,Cout] or NULL */
  float *dw;
} hav_conv_wgrad_args;

int hav_conv2d_wgrad(const hav_conv_wgrad_args *args, void *stream);
int hav_rowscale_dot(float *out, float *dot, const float *a, const float *x, const float *scale, int64_t rows, int64_t n,
                     void *stream);

/* ------------------------------------------------------------------------------------------------
 * Adam over flat fp32 buffers (replaces the torch.optim.Adam steps of train_avatar.py:151 and train_avatarHD.py:231,279-280;
 * same update rule as torch/optim/adam.py with amsgrad=False, weight_decay=0, maximize=False).
 *   state[0] = number of steps taken so far (incremented by the call), state[1] = learning rate -- device memory, so the call
 *   can be captured in a CUDA graph.  grad is multiplied by grad_scale before use; zero_grad != 0 clears it in the same pass.
 *   n must be a multiple of 4 and the buffers 16-byte aligned (pad the flat layout).
 */
int hav_adam_flat(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float *state, float beta1, float beta2,
                  float eps, float grad_scale, int zero_grad, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HAVATAR_B200_H_ */
