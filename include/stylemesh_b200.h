/* stylemesh_b200 — C-ABI of the B200-native StyleMesh texture-optimisation hot path.
 *
 * The reference (lukasHoel/stylemesh) has NO native/FFI boundary: every FLOP of its hot path runs inside
 * third-party torch ops called from Python.  This header is therefore the boundary a maintainer would bind
 * (ctypes / cffi / a torch custom-op shim) underneath the reference's Python modules; each entry point names the
 * reference call site(s) it replaces.  Conventions:
 *   - plain C types only; every tensor is a raw DEVICE pointer plus explicit sizes (host pointers are marked);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is asynchronous on it;
 *   - return value: 0 = ok, negative = error (message via smb_last_error(), thread-local);
 *   - the library never touches the reference's CPU path and has no CPU fallback.
 *
 * Storage formats
 *   image / texture / gradients : fp32 planar (C,H,W), exactly the reference tensors.
 *   UV grid                     : fp32 (H,W,2) in [-1,1]  (model/texture/utils.py:56-60 to_grid()).
 *   VGG activations (internal)  : channels-last bf16 hi/lo planes, x = hi + lo (fp32-grade);
 *                                 exported to fp32 NCHW on request.
 */
#ifndef STYLEMESH_B200_H
#define STYLEMESH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMB_ABI_VERSION 1
#define SMB_NUM_VGG_CONVS 13 /* conv1_1 .. conv5_1: the layers the style/content loss can read */

/* implementation selectors for smb_ctx_set_impl */
#define SMB_IMPL_SIMT 0 /* fp32 CUDA-core cross-check kernels            */
#define SMB_IMPL_TC 1   /* tcgen05 tensor-core kernels (Gram default; convs: persistent stream-K kernel) */
#define SMB_IMPL_TC_PH 5 /* convs only (DEFAULT): CTA pair + activation halo + TMA-store epilogue for 3x3 convs, stream-K kernel otherwise */
/* (2, 3, 4 were earlier tcgen05 conv generations; retired from the library, see git history) */

typedef struct smb_ctx smb_ctx;

int smb_abi_version(void);
const char* smb_last_error(void);

/* ---------------------------------------------------------------------------------------------------------
 * Texture side
 * ------------------------------------------------------------------------------------------------------- */

/* Replaces NeuralTexture.forward / HierarchicalNeuralTexture.forward
 * (model/texture/texture.py:41-54, 96-100): out[c][y][x] = sum_l bilinear(clamp(layer_l), grid), border padding,
 * align_corners=True.  The clamp is applied on read; the stored texture is clamped by smb_adam_step.
 * layers: HOST array of num_layers device pointers, layer l is (channels, layer_h[l], layer_w[l]) fp32. */
int smb_uv_sample_fwd(const float* const* layers, const int* layer_w, const int* layer_h, int num_layers,
                      int channels, const float* grid, int H, int W, float clamp_lo, float clamp_hi, float* out,
                      void* stream);

/* Debug/parity export of the exact UV index arithmetic (ATen GridSampler.h:27-36,58-60,143-171):
 * xy0[p] = (x0,y0) north-west texel, w4[p] = (nw,ne,sw,se) weights, for a (tex_h, tex_w) layer. */
int smb_uv_texel_index(const float* grid, int num_pixels, int tex_w, int tex_h, int* xy0, float* w4, void* stream);

/* Replaces grid_sampler_2d_backward (input gradient only) for all layers plus the two gradient hooks of
 * model/model.py:198-202 (angle) and :247-251 (depth interpolation weight):
 *   grad_layers[l][c][y][x] += w_corner * grad_out[c][p] * hook0[p] * hook1[p]      (hook pointers may be NULL)
 * Accumulates (the caller / smb_adam_step keeps the buffers zeroed). */
int smb_uv_scatter_bwd(float* const* grad_layers, const int* layer_w, const int* layer_h, int num_layers,
                       int channels, const float* grid, int H, int W, const float* grad_out, const float* hook0,
                       const float* hook1, void* stream);

/* Replaces torch.optim.Adam.step (model/model.py:395) fused with NeuralTexture.normalize() (texture.py:41-44) and
 * the gradient of HierarchicalNeuralTexture.regularizer (texture.py:102-108):
 *   x = clamp(p); g' = grad*grad_scale + reg_coef*x; Adam(m, v, g'); p = x - lr/bc1 * m/(sqrt(v)/sqrt(bc2)+eps);
 *   grad = 0.   step is 1-based.  reg_coef = lambda_reg * w_l * 2 / numel(layer). */
int smb_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float clamp_lo, float clamp_hi, float reg_coef,
                  float grad_scale, void* stream);

/* Value of one regulariser term: *out_accum += coef * sum(clamp(param)^2)   (coef = lambda_reg*w_l/numel). */
int smb_texreg_value(const float* param, int64_t n, float coef, float clamp_lo, float clamp_hi, float* out_accum,
                     void* stream);

/* smb_adam_step / smb_texreg_value over ONE flat buffer holding all texture layers back to back
 * (HierarchicalNeuralTexture.layers, texture.py:80-81), in a single launch each.  Segment l covers the elements
 * [seg_begin[l], seg_begin[l+1]) (the last one ends at n) and has its own coefficient; seg_begin / seg_*coef are HOST
 * arrays of num_segments (<= 8) entries, seg_begin[0] == 0, every offset a multiple of 4.  Padding elements between
 * layers must be zero in all four buffers (they stay zero). */
int smb_adam_step_segments(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                           const int64_t* seg_begin, const float* seg_reg_coef, int num_segments, float lr, float beta1,
                           float beta2, float eps, int step, float clamp_lo, float clamp_hi, float grad_scale,
                           void* stream);
int smb_texreg_value_segments(const float* param, int64_t n, const int64_t* seg_begin, const float* seg_coef,
                              int num_segments, float clamp_lo, float clamp_hi, float* out_accum, void* stream);

/* Multi-GPU form of smb_adam_step_segments for view-sharded training (one process per GPU; replaces
 * DistributedDataParallel's gradient all-reduce + torch.optim.Adam.step of the reference's `--gpus N` runs,
 * model/optimize.py:30, model/model.py:395): reduce-scatter of the gradient, Adam on this rank's 1/world slice (the
 * moments are sharded: only that slice of exp_avg / exp_avg_sq is used) and all-gather of the new texels, in one
 * kernel over NVLink peer memory, followed by a kernel that zeroes the local gradient once every peer is done.
 * grad_ptrs / param_ptrs / flag_ptrs: HOST arrays of `world` DEVICE pointers to every rank's flat gradient, flat
 * parameter and flag buffer (own rank included; e.g. torch.distributed._symmetric_memory buffer_ptrs).  Flag buffers:
 * at least 64 uint32, zero before the first call.  epoch: strictly increasing per call (the step counter); all ranks
 * must issue the same sequence of calls.  The gradient is averaged over ranks (scale 1/world).  n: multiple of 4. */
int smb_dist_adam_step(int rank, int world, float* const* grad_ptrs, float* const* param_ptrs,
                       unsigned int* const* flag_ptrs, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const int64_t* seg_begin, const float* seg_reg_coef, int num_segments, float lr, float beta1,
                       float beta2, float eps, int step, float clamp_lo, float clamp_hi, unsigned int epoch,
                       void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * View preparation: the per-pixel work of Abstract_Dataset.__getitem__ (data/abstract_dataset.py:270-344) on the
 * device, so that a scene's views can be prepared once and stay resident in HBM.  All
 * pointers are device pointers unless marked HOST.  Resampling follows the libraries the reference calls (cv2.resize
 * INTER_LINEAR / INTER_NEAREST, PIL NEAREST): the caller builds their index / weight tables on the host
 * (stylemesh_b200/data/resample.py) and the kernels apply them, so indices are bit-exact by construction.
 * ------------------------------------------------------------------------------------------------------- */

/* get_uv_transform (model/texture/utils.py:87-91): grid = fl(fl(2 uv) - 1) of channels 0,1 of the renderer's
 * (H,W,3) float32 map; optional calculate_mask (data/scannet_dataset.py:308-326): valid = (u != 0) | (v != 0),
 * and (depth_at_uv > 0) when a depth map resampled to (H,W) is given (ScanNet; Matterport passes NULL). */
int smb_view_uv_to_grid(const float* uv_hw3, int H, int W, float* grid_hw2, unsigned char* valid,
                        const double* depth_at_uv, void* stream);

/* dst[y][x] = src[ytab[y]][xtab[x]] for 1- or 4-byte elements (nearest resampling of the mask / angle map,
 * data/abstract_dataset.py:306-311). */
int smb_view_gather2d(const void* src, int elem_bytes, int Hs, int Ws, const int* ytab, const int* xtab, int Hd, int Wd,
                      void* dst, void* stream);

/* cv2.resize(depth, INTER_LINEAR) (data/abstract_dataset.py:301-304) into float64: src_type 0 = float64,
 * 1 = float32 (rendered depth; float32 arithmetic), 2 = uint16 sensor depth divided by `divisor` (1000.0 ScanNet
 * data/scannet_dataset.py:301, 4000.0 Matterport) in float64 first.  Same size in and out: conversion only. */
int smb_view_resize_linear(const void* src, int src_type, double divisor, int Hs, int Ws, const int* yofs,
                           const double* yalpha, const int* xofs, const double* xalpha, int Hd, int Wd, double* dst,
                           void* stream);

/* calculate_depth_level (data/scannet_dataset.py:328-366) in numpy's float64 arithmetic: continuous level (float32),
 * nearest and second-nearest level (int64) and the interpolation weight of the nearest (float32) per pixel; also the
 * float32 copy of the depth that transform_label produces (may be NULL).  levels: HOST array, ascending UV heights. */
int smb_view_depth_levels(const double* depth, int64_t n, const double* levels, int num_levels, double min_depth,
                          int depth_is_f32, float* depth_level, float* depth_f32, int64_t* rounded, int64_t* other,
                          float* weight, void* stream);

/* ToTensor + pre() (model/losses/rgb_transform.py:5-11): uint8 (H,W,3) RGB -> float32 (3,H,W) BGR, (x/255 - mean)*255. */
int smb_view_rgb_pre(const unsigned char* rgb_hwc, int H, int W, float* out_chw, void* stream);

/* angle_degrees = rad2deg(acos(cos_angle)) (data/abstract_dataset.py:338). */
int smb_view_angle_degrees(const float* cos_angle, int64_t n, float* degrees, void* stream);

/* erode of model/model.py:204-208: out = x where the zero-padded 3x3 box mean of x is exactly 1, else 0. */
int smb_view_erode3x3(const float* x, int H, int W, float* out, void* stream);

/* Mask pyramid of a view (model/model.py:204-254 + content_and_style_losses.py:161,172-185) in 1 + L launches.
 * smb_view_level_masks: all L pyramid levels at the rgb resolution (H x W): level_mask[l] = erode(((rounded == l) |
 *   (other == l)) & mask), level_weight[l] = erode((rounded == l) & mask) * w + erode((other == l) & mask) * (1 - w);
 *   mask: H*W bytes (0 / non-zero), rounded / other: int64, outputs [L][H*W] fp32. */
int smb_view_level_masks(const unsigned char* mask, const int64_t* rounded, const int64_t* other,
                         const float* interp_w, int H, int W, int num_levels, float* level_mask, float* level_weight,
                         void* stream);
/* smb_view_level_plan: one pyramid level of size H x W from maps at the rgb resolution Hr x Wr.
 *   src_mask (values > 0 select; nearest-resampled), src_weight -> hook1 [H*W] (nearest; both may be NULL),
 *   angle_guidance -> hook0 [H*W] (bilinear; both may be NULL); for each of the num_layers (<= 8) VGG layers of size
 *   lh[k] x lw[k]: the nearest-downsampled {0,1} row mask and, with split != 0, the masks of the pixels whose
 *   bilinear angle_degrees is below / not below `threshold`, written back to back into layer_masks
 *   ([all | pass | fail] per layer, or [all] without the split).
 *   counts: 1 + 3*num_layers uint32 (device): selected level pixels, then (n, n_pass, n_fail) per layer. */
int smb_view_level_plan(const float* src_mask, const float* src_weight, const float* angle_guidance,
                        const float* angle_degrees, float threshold, int Hr, int Wr, int H, int W, float* hook0,
                        float* hook1, int num_layers, const int* lh, const int* lw, float* layer_masks, int split,
                        unsigned int* counts, void* stream);

/* Texture export / headless preview.
 * smb_texture_post_rgb8: post() (model/losses/rgb_transform.py:14-21) + ToPILImage quantisation
 *   (model/texture/texture.py:9-19): pre()-space BGR fp32 (3,H,W) -> RGB uint8 (H,W,3), device to device.
 * smb_mip_downsample2x: next mip level ((3, max(H/2,1), max(W/2,1)) fp32) by a 2x2 box filter.
 * smb_mip_preview: the styled view of model/optimize.py:181-208 without OpenGL: trilinear lookup of the (u, v[, lod])
 *   map (H,W,uv_channels; u = v = 0 marks pixels without geometry) into the mip chain, post(), RGB uint8 (H,W,3). */
int smb_texture_post_rgb8(const float* bgr_chw, int H, int W, unsigned char* rgb_hwc, void* stream);
int smb_mip_downsample2x(const float* src_chw, int H, int W, float* dst_chw, void* stream);
int smb_mip_preview(const float* const* mips, const int* mip_w, const int* mip_h, int num_mips, const float* uv,
                    int uv_channels, int H, int W, float lod_bias, unsigned char* rgb_hwc, void* stream);

/* UV / angle / depth rasteriser: one camera pose of the reference's OpenGL renderer
 * (scripts/scannet/render_uv: src/renderer/renderer.cpp:165-224, scannet_renderer.cpp:19-62, include/util.h:11-35,
 * shader/{uvmap,angle,depth}.{vs,frag}) without a GL context.  All pointers are device pointers except view3x4 / proj6.
 *   verts [nv][3] fp32, faces [nf][3] int32, corner_uv [nf][3][2], corner_normal [nf][3][3] (per face corner);
 *   view3x4: HOST, the first three rows of the view matrix (row-major); proj6: HOST, (P00, P02, P11, P12, P22, P23) of
 *   the projection; w x h: output size; tex_size: size of the texture the LOD is queried against (1024 in the reference);
 *   eye_scratch [nv][4] fp32 and zbuf [h*w] uint64: work buffers;
 *   uv_out = (u, v, lod), angle_out = cos(view angle) x 3, depth_out = eye depth x 3, each [h][w][3] fp32; pixels
 *   without geometry are 0; flip != 0 writes GL row j to image row h-1-j (Renderer::saveUV). */
int smb_raster_view(const float* verts, int num_verts, const int* faces, int num_faces, const float* corner_uv,
                    const float* corner_normal, const float* view3x4, const float* proj6, int w, int h, float near_plane,
                    float far_plane, float tex_size, int flip, float* eye_scratch, unsigned long long* zbuf,
                    float* uv_out, float* angle_out, float* depth_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * VGG / loss engine (replaces model/losses/content_and_style_losses.py: VGG.forward :47-70,
 * GramMatrix :74-80, masked_features :136-143, the loss loop of ContentAndStyleLoss.forward :298-348, and the
 * autograd backward of all of them).
 * ------------------------------------------------------------------------------------------------------- */
smb_ctx* smb_ctx_create(void);
void smb_ctx_destroy(smb_ctx* ctx);

/* conv_impl: SMB_IMPL_SIMT, SMB_IMPL_TC or SMB_IMPL_TC_PH (default); gram_impl: SMB_IMPL_SIMT or SMB_IMPL_TC. */
int smb_ctx_set_impl(smb_ctx* ctx, int conv_impl, int gram_impl);

/* weights_oihw[i]: HOST fp32 (Cout,Cin,3,3) of conv i in state_dict order conv1_1..conv5_1; bias[i]: HOST (Cout). */
int smb_ctx_load_vgg(smb_ctx* ctx, const float* const* weights_oihw, const float* const* bias, int num_convs);

/* Select (and lazily allocate) the working set for an H x W input; returns a slot id >= 0.  Several slots can be
 * alive at once (one per pyramid level / view resolution); a slot keeps its forward activations until reused. */
int smb_level_begin(smb_ctx* ctx, int H, int W);

/* VGG forward of image (3,H,W) fp32 up to and including conv `last_conv` (0 = conv1_1 .. 12 = conv5_1). */
int smb_level_forward(smb_ctx* ctx, int slot, const float* image, int last_conv, void* stream);

/* Inference-only forward (the content-target pass VGG(rgb), cs:294, and any other feature extraction that is never
 * back-propagated): same arithmetic, but only the layers whose bit is set in keep_mask (bit i = conv index i) are
 * guaranteed to be readable afterwards (smb_level_get_feature*, smb_level_gram).  A layer that only feeds a max-pool
 * is pooled inside the conv epilogue and never written at full resolution.  Loss terms and smb_level_backward are
 * rejected on the slot until the next smb_level_forward. */
int smb_level_forward_features(smb_ctx* ctx, int slot, const float* image, int last_conv, unsigned int keep_mask,
                               void* stream);

/* Export relu(conv_i) as fp32 (C,h,w); query its shape. */
int smb_level_feature_shape(smb_ctx* ctx, int slot, int conv, int* C, int* h, int* w);
int smb_level_get_feature(smb_ctx* ctx, int slot, int conv, float* out_nchw, void* stream);

/* Export relu(conv_i) as fp32 channels-last (h*w, C) — the layout smb_level_content_term takes its target in. */
int smb_level_get_feature_nhwc(smb_ctx* ctx, int slot, int conv, float* out_nhwc, void* stream);

/* Free every slot's device memory (e.g. after the one-off style-target pass over large style images). */
int smb_ctx_release_slots(smb_ctx* ctx);

/* Masked Gram of relu(conv_i):  G = inv_n * sum_p m_p F_p F_p^T  -> gram_out (C x C fp32). rowmask: one float per
 * pixel, exactly 0 or 1 (the reference compacts features by `mask > 0`, cs:136-143), or NULL for all ones. */
int smb_level_gram(smb_ctx* ctx, int slot, int conv, const float* rowmask, float inv_n, float* gram_out,
                   void* stream);

/* One style term on relu(conv_i): masked Gram, (optional running average), weighted MSE against up to two
 * C x C targets, and the gradient w.r.t. the features accumulated into the slot's pending gradient of that layer.
 *   loss_accum[0] += coef0 * mean((Y0 - Ghat)^2) + coef1 * mean((Y1 - Ghat)^2)
 * prev_sum/avg_len implement gram_mode 'average' (cs:319-323): Ghat = (G + prev_sum)/avg_len; pass NULL/1. */
int smb_level_style_term(smb_ctx* ctx, int slot, int conv, const float* rowmask, float inv_n, const float* target0,
                         float coef0, const float* target1, float coef1, const float* prev_sum, float avg_len,
                         float* gram_out, float* loss_accum, void* stream);

/* Content term on relu(conv_i): loss_accum[0] += coef_loss * sum_p m_p |F_p - T_p|^2 and the gradient
 * coef_grad * m_p * (F_p - T_p) accumulated into the pending gradient.  target: fp32 channels-last (h*w, C). */
int smb_level_content_term(smb_ctx* ctx, int slot, int conv, const float* target_nhwc, const float* rowmask,
                           float coef_loss, float coef_grad, float* loss_accum, void* stream);

/* Back-propagate every pending term of the slot to the input image: d_image (3,H,W) fp32 is overwritten.
 * Clears the pending gradients. */
int smb_level_backward(smb_ctx* ctx, int slot, float* d_image, void* stream);

/* Number of kernel launches this library has issued in the process so far (bench.py: "gpu_launches"). */
int64_t smb_launch_count(void);

/* Per-kernel-class device timing with CUDA events on the launch stream (bench.py's roofline pass; adds two event
 * records per launch, so it is off by default and never on during the headline timed region).
 * Classes: 0 conv1_1 fwd, 1 igemm conv fwd, 2 igemm data-grad, 3 igemm Gram-backward, 4 Gram, 5 Gram-MSE,
 *          6 pool fwd/bwd, 7 conv1_1 data-grad, 8 content MSE, 9 misc (mask / ReLU-split).
 * smb_ctx_read_timing synchronises, returns accumulated milliseconds, algorithmic FLOPs (2*P*N*K*taps) and launch
 * counts per class since the last read, and resets them.  Returns the number of classes (10). */
#define SMB_NUM_TIMING_CLASSES 10
int smb_ctx_set_timing(smb_ctx* ctx, int enabled);
int smb_ctx_read_timing(smb_ctx* ctx, float* ms, double* flops, int* launches, int n);

/* Bytes of device memory currently owned by the context. */
int64_t smb_ctx_device_bytes(smb_ctx* ctx);

/* Profiling aid: when `buf` (device memory, >= 148 * 16 uint64) is non-NULL every following stream-K conv launch
 * writes a per-CTA timeline (clock64 / globaltimer stamps and barrier-wait totals, layout in tc_igemm_v2.cu) into
 * it; NULL switches the tracing off again.  Used by tools/gpu_trace_probe.py, never on in tests or the bench. */
int smb_debug_set_igemm_trace(void* buf);

/* ---------------------------------------------------------------------------------------------------------
 * Unit-level entry points (used by the parity tests to localise a failure to one kernel).
 * tensors are fp32 NCHW on the device; conversion to the internal planes happens inside.
 * ------------------------------------------------------------------------------------------------------- */
/* y = [relu](conv3x3(x, w, b)), x (Cin,H,W), w HOST (Cout,Cin,3,3), b HOST (Cout) or NULL, y (Cout,H,W).
 * transpose_flip != 0 computes the data gradient instead: x is dY (Cout,H,W), y is dX (Cin,H,W), no bias. */
int smb_unit_conv3x3(int impl, const float* x, int Cin, int H, int W, const float* w_host, const float* b_host,
                     int Cout, int relu, int transpose_flip, float* y, void* stream);
/* y (Cout,H,W) = conv3x3(x (Cin,H,W), w (Cout,Cin,3,3)) + m * (g f): the pair + halo kernel with a fused 1x1 term
 * (f (Cout,H,W) device, g (Cout,Cout) host, rowmask (H*W) device {0,1} or NULL) - the Gram backward folded into a
 * data-gradient conv (model/losses/content_and_style_losses.py:74-80 backward). */
int smb_unit_conv3x3_fused(const float* x, int Cin, int H, int W, const float* w_host, int Cout, const float* f,
                           const float* g_host, const float* rowmask, float* y, void* stream);
/* y (C,H/2,W/2) = maxpool2x2(x (C,H,W)) */
int smb_unit_maxpool(const float* x, int C, int H, int W, float* y, void* stream);
/* dx (C,H,W) = maxpool/relu backward of g (C,H/2,W/2) through y (C,H,W) */
int smb_unit_maxpool_bwd(const float* g, const float* y, int C, int H, int W, float* dx, void* stream);
/* G (C,C) = inv_n * sum_p m_p F_p F_p^T for F (C,H,W) */
int smb_unit_gram(int impl, const float* f, int C, int H, int W, const float* rowmask, float inv_n, float* G,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STYLEMESH_B200_H */
