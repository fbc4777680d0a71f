// Internal (C++) launcher declarations shared by the .cu translation units and the engine.
// Nothing here crosses the C-ABI; see include/stylemesh_b200.h for the exported surface.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#if defined(__CUDACC__)
#define SMB_HD __host__ __device__
#else
#define SMB_HD
#endif

#define SMB_MAX_TEX_LAYERS 8
#define SMB_MAX_TEX_CHANNELS 4

namespace smb {

// ---- texture side -------------------------------------------------------------------------------------------
struct TexLayerSet {
  float* ptr[SMB_MAX_TEX_LAYERS];  // each (C, H_l, W_l) fp32 planar, as the reference Parameter (texture.py:30-32)
  int W[SMB_MAX_TEX_LAYERS];
  int H[SMB_MAX_TEX_LAYERS];
  int L;
  int C;
};

int launch_uv_sample_fwd(const TexLayerSet& tex, const float* grid, int H, int W, float clamp_lo, float clamp_hi,
                         float* out, cudaStream_t st);
int launch_uv_texel_index(const float* grid, int npix, int W, int H, int* xy0, float* w4, cudaStream_t st);
int launch_uv_scatter_bwd(const TexLayerSet& gtex, const float* grid, int H, int W, const float* gout,
                          const float* hook0, const float* hook1, cudaStream_t st);
int launch_adam(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                int step, float clamp_lo, float clamp_hi, float reg_coef, float gscale, cudaStream_t st);
int launch_sumsq_clamped(const float* x, int64_t n, float coef, float clamp_lo, float clamp_hi, float* out,
                         cudaStream_t st);
// the two kernels above over a flat buffer of consecutive segments (host arrays seg_begin[num], coef[num]), one launch
int launch_adam_segments(float* p, float* g, float* m, float* v, int64_t n, const int64_t* seg_begin,
                         const float* seg_reg_coef, int num_segments, float lr, float beta1, float beta2, float eps,
                         int step, float clamp_lo, float clamp_hi, float gscale, cudaStream_t st);
int launch_sumsq_segments(const float* x, int64_t n, const int64_t* seg_begin, const float* seg_coef, int num_segments,
                          float clamp_lo, float clamp_hi, float* out, cudaStream_t st);
// N > 1: gradient reduce-scatter + Adam on this rank's slice + parameter all-gather over peer memory (HOST arrays of
// `world` device pointers: every rank's gradient / parameter / flag buffers, own rank included)
int launch_dist_adam(int rank, int world, float* const* grad_ptrs, float* const* param_ptrs,
                     unsigned int* const* flag_ptrs, float* m, float* v, int64_t n, const int64_t* seg_begin,
                     const float* seg_reg_coef, int num_segments, float lr, float beta1, float beta2, float eps,
                     int step, float clamp_lo, float clamp_hi, unsigned int epoch, cudaStream_t st);

// ---- view preparation (view_prep_kernels.cu) ----------------------------------------------------------------------
int launch_view_uv_grid(const float* uv3, int H, int W, float* grid2, unsigned char* valid, const double* depth,
                        cudaStream_t st);
int launch_view_gather2d(const void* src, int elem_bytes, int Ws, const int* ytab, const int* xtab, int Hd, int Wd,
                         void* dst, cudaStream_t st);
int launch_view_resize_linear(const void* src, int src_type, double divisor, int Hs, int Ws, const int* yofs,
                              const double* ya, const int* xofs, const double* xa, int Hd, int Wd, double* dst,
                              cudaStream_t st);
int launch_view_depth_levels(const double* depth, int64_t n, const double* levels_host, int num_levels, double min_depth,
                             int depth_is_f32, float* cont, float* depth_f32, long long* rounded, long long* other,
                             float* weight, cudaStream_t st);
int launch_view_rgb_pre(const unsigned char* hwc, int H, int W, float* chw, cudaStream_t st);
int launch_view_angle_degrees(const float* c, int64_t n, float* deg, cudaStream_t st);
int launch_view_erode3x3(const float* x, int H, int W, float* out, cudaStream_t st);
#define SMB_MAX_MIP_LEVELS 16
int launch_texture_post_rgb8(const float* bgr_chw, int H, int W, unsigned char* rgb_hwc, cudaStream_t st);
int launch_mip_downsample2x(const float* src, int Hs, int Ws, float* dst, cudaStream_t st);
int launch_mip_preview(const float* const* mips, const int* mW, const int* mH, int num_mips, const float* uv,
                       int uv_channels, int H, int W, float lod_bias, unsigned char* rgb_hwc, cudaStream_t st);
int launch_raster_view(const float* verts, int nv, const int* faces, int nf, const float* corner_uv,
                       const float* corner_normal, const float* view3x4, const float* proj6, int w, int h, float near,
                       float far, float tex_size, int flip, float* eye_scratch, unsigned long long* zbuf, float* uv_out,
                       float* angle_out, float* depth_out, cudaStream_t st);
#define SMB_MAX_PLAN_LAYERS 8
int launch_view_level_masks(const unsigned char* mask, const long long* rounded, const long long* other,
                            const float* interp_w, int H, int W, int L, float* level_mask, float* level_weight,
                            cudaStream_t st);
int launch_view_level_plan(const float* src_mask, const float* src_weight, const float* angle_guidance,
                           const float* angle_degrees, float threshold, int Hr, int Wr, int H, int W, float* hook0,
                           float* hook1, int num_layers, const int* lh, const int* lw, float* layer_masks, int split,
                           unsigned int* counts, cudaStream_t st);

// ---- VGG side -----------------------------------------------------------------------------------------------
// Activation planes (see smb_common.cuh): channels-last bf16 hi/lo pair.
struct Act {
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int H = 0, W = 0, C = 0;
  SMB_HD int64_t pixels() const { return (int64_t)H * W; }
  SMB_HD int64_t elems() const { return (int64_t)H * W * C; }
};

// Packed GEMM-B operand: [taps][N][K] bf16 hi/lo (K contiguous).  Forward conv: N=Cout, K=Cin,
// value W[n][k][r][s] at tap r*3+s.  Data-gradient conv: N=Cin, K=Cout, value W[k][n][2-r][2-s].
struct PackedB {
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int taps = 0, N = 0, K = 0;
};

// Epilogue applied to an accumulator tile acc[p][n] of an implicit-GEMM:
//   v = acc; v *= rowscale[p]; v += bias[n]; v += addend[p][n]; v = sign_hi[p][n] > 0 ? v : 0; v = relu ? max(v,0) : v
// then any subset of: out (hi/lo), out_f32 [p][n], outm (hi/lo) = v * rowmask[p].
struct Epilogue {
  const float* rowscale = nullptr;        // [P]
  const float* bias = nullptr;            // [N]
  const float* addend = nullptr;          // [P][N] fp32
  const __nv_bfloat16* sign_hi = nullptr; // [P][N] (hi plane of the forward activation: ReLU backward mask)
  int relu = 0;
  int pool2x2 = 0;                        // igemm_ph only: out_hi/out_lo are the (H/2, W/2) planes of maxpool2x2(v)
  __nv_bfloat16* pool_hi = nullptr;       // igemm_ph only, next to out_hi/out_lo: side output maxpool2x2(v) as (H/2, W/2)
  __nv_bfloat16* pool_lo = nullptr;       //   planes (training forward of a layer that feeds a pool: no pool launch)
  __nv_bfloat16* out_hi = nullptr;
  __nv_bfloat16* out_lo = nullptr;
  float* out_f32 = nullptr;
  float* out_planar3 = nullptr;           // (3,H,W) fp32: channels 0..2 of the result, planar (first-layer data gradient)
  const float* rowmask = nullptr;         // [P]  (only with outm_*)
  __nv_bfloat16* outm_hi = nullptr;
  __nv_bfloat16* outm_lo = nullptr;
};

enum ConvImpl : int { IMPL_SIMT = 0, IMPL_TC = 1, IMPL_TC_PH = 5 };   // 2..4: retired tcgen05 generations (git history)

// generic 3x3 (taps==9, pad 1) or 1x1 (taps==1) implicit GEMM: out[p][n] = sum_tap sum_k A[p+off(tap)][k] * B[tap][n][k]
int launch_igemm_simt(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st);
int launch_igemm_tc2(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st);   // v2: persistent stream-K
void set_igemm_trace(unsigned long long* buf);   // tc_igemm_v2.cu: per-CTA timeline of the following launches
// Optional 1x1 term accumulated by igemm_ph into the same tile before the epilogue:
//   acc[p][n] += rowmask[p] * sum_k f[p][k] * g[n][k]       (Gram backward of the layer that receives the gradient:
//   f = its forward features, g = the symmetric seed matrix of gram_mse, rowmask = the term's {0,1} pixel mask)
struct FusedTerm {
  Act f;                            // same H x W as the conv input, C a multiple of 64
  PackedB g;                        // taps == 1, N == conv N, K == f.C
  const float* rowmask = nullptr;   // [H*W] or nullptr
};
int launch_igemm_ph(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st,
                    const FusedTerm* ft = nullptr);    // v5: CTA pair + A halo + TMA-store epilogue (3x3 only)

// first layer (3 -> 64) from the fp32 planar image and its data gradient (64 -> 3)
int launch_conv_first_fwd(const float* img, int H, int W, const float* w_oihw, const float* bias, int Cout,
                          const Epilogue& ep, cudaStream_t st);
int launch_conv_first_dgrad(const Act& dz, const float* w_oihw, int Cout, float* dimg, cudaStream_t st);
// tcgen05 version of the first layer: w_hi/w_lo = [64][32] bf16 (k = ci*9 + r*3 + s, zero padded to 32)
int launch_conv_first_tc(const float* img, int H, int W, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo,
                         const float* bias, const Epilogue& ep, cudaStream_t st);

int launch_maxpool_fwd(const Act& in, const Act& out, cudaStream_t st);
// dz[p][c] = y[p][c] > 0 ? (p is first arg-max of its 2x2 window ? g[window][c] : 0) + addend[p][c] : 0
int launch_maxpool_bwd_relu(const float* g_pooled, const float* addend /*nullable*/, const Act& y, const Act& dz,
                            cudaStream_t st);

int launch_act_from_nchw(const float* src, const Act& dst, cudaStream_t st);
int launch_act_to_nchw(const Act& src, float* dst, cudaStream_t st);
int launch_mask_rows(const Act& src, const float* rowmask, const Act& dst, cudaStream_t st);

// Gram: partial[s][i][j] = sum_{p in split s} Fm[p][i] * Fm[p][j]   (unnormalised), s in [0, nsplit)
// The tcgen05 kernel takes the features and the {0,1} pixel mask (nullable) and zeroes masked pixels in shared
// memory; the SIMT kernel wants the masked copy (launch_mask_rows).
int gram_num_splits(int64_t P, int C, int impl);
int launch_gram_simt(const Act& fm, float* partial, int nsplit, cudaStream_t st);
int launch_gram_tc(const Act& f, const float* rowmask, float* partial, int nsplit, cudaStream_t st);

// Gram loss + gradient seed.  G = inv_n * sum_s partial[s];  if prev_sum: Ghat = (G + prev_sum) / avg_len else Ghat = G
//   loss_out[0] += sum_t coef[t] * mean((Y_t - Ghat)^2)
//   dGhat = sum_t coef[t] * 2 (Ghat - Y_t) / C^2 ;  Bmat = (2 * inv_n / avg_len) * dGhat  -> hi/lo [C][C] (symmetric)
//   g_out (optional) = G (this step's Gram, for the 'average' cache and for inspection)
int launch_gram_mse(const float* partial, int nsplit, int C, float inv_n, const float* y0, float coef0,
                    const float* y1, float coef1, const float* prev_sum, float avg_len, float* g_out,
                    __nv_bfloat16* b_hi, __nv_bfloat16* b_lo, float* loss_out, cudaStream_t st);

// content: loss_out[0] += coef_loss * sum_p m_p sum_c (T - F)^2 ; addend[p][c] += coef_grad * m_p * (F - T)
int launch_content_mse(const Act& f, const float* target_nhwc, const float* rowmask, float coef_loss,
                       float coef_grad, float* addend, float* loss_out, cudaStream_t st);

// dz = (g ⊙ (y > 0)) for the top of the backward chain when g is fp32 [P][C]
int launch_relu_mask_split(const float* g, const Act& y, const Act& dz, cudaStream_t st);

}  // namespace smb
