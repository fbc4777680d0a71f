// K5: the per-view mask pyramid of a step in 1 + L launches and ONE host read-back.
//
// The reference builds these on every step with ~20 eager torch ops per pyramid level and several host syncs
// (model/model.py:204-257 erode / mask_depth / mask_interpolation_weight / the nearest + bilinear resamplings of the
// hooks :198-202,247-254; model/losses/content_and_style_losses.py:161,172-185 per-layer nearest masks, the angle
// pass / fail split and the per-layer means):
//
//   view_level_masks   rgb resolution, all L levels at once:
//                        level_mask[l]   = erode(((rounded == l) + (other == l)) * mask)                (model.py:211-218)
//                        level_weight[l] = erode((rounded == l) * mask) * w + erode((other == l) * mask) * (1 - w)  (:225-237)
//   view_level_plan    one pyramid level (H x W) and its VGG layers (h_k x w_k), one flat index space:
//                        hook0 = bilinear(angle_guidance -> H x W)                                      (model.py:199)
//                        hook1 = nearest(level_weight -> H x W)                                         (model.py:238)
//                        alive = #{nearest(level_mask -> H x W) > 0}                                    (model.py:219,256)
//                        per layer: m = nearest(M -> h_k x w_k), m_pass = nearest(M * passed), m_fail = nearest(M * ~passed)
//                        with passed = bilinear(angle_degrees -> H x W) < threshold, and their pixel counts (cs:161-185)
//
// torch semantics restated exactly (ATen/native/UpSample.h): nearest src = min(floor(dst * (float)in / out),
// in - 1) with the in == out and out == 2 in shortcuts; bilinear align_corners = False: src = max(0, (in / out) *
// (dst + 0.5) - 0.5), i1 = i0 + (i0 < in - 1), lambda1 = src - i0.  erode(x): x where the zero-padded 3x3 box sum / 9,
// clamped to [0, 1], equals 1 - the mask values here are 0 / 1 (the sum of two bool tensors is their OR in torch), so
// that is "all nine zero-padded neighbours set".
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
  if (in == out) return dst;
  if (out == 2 * in) return dst >> 1;
  const float scale = __fdiv_rn((float)in, (float)out);
  const int s = (int)floorf(__fmul_rn((float)dst, scale));
  return s < in - 1 ? s : in - 1;
}

// bilinear, align_corners = False (UpSample.h area_pixel_compute_source_index + upsample_bilinear2d)
__device__ __forceinline__ float bilinear_at(const float* __restrict__ src, int Hs, int Ws, int Hd, int Wd, int y, int x) {
  const float sy = __fdiv_rn((float)Hs, (float)Hd), sx = __fdiv_rn((float)Ws, (float)Wd);
  float fy = __fsub_rn(__fmul_rn(sy, __fadd_rn((float)y, 0.5f)), 0.5f);
  float fx = __fsub_rn(__fmul_rn(sx, __fadd_rn((float)x, 0.5f)), 0.5f);
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
  const float ly1 = __fsub_rn(fy, (float)y0), lx1 = __fsub_rn(fx, (float)x0);
  const float ly0 = __fsub_rn(1.f, ly1), lx0 = __fsub_rn(1.f, lx1);
  const float v00 = src[(int64_t)y0 * Ws + x0], v01 = src[(int64_t)y0 * Ws + x1];
  const float v10 = src[(int64_t)y1 * Ws + x0], v11 = src[(int64_t)y1 * Ws + x1];
  const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
  const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
  return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

__global__ void __launch_bounds__(256) view_level_masks_kernel(const unsigned char* __restrict__ mask,
                                                               const long long* __restrict__ rounded,
                                                               const long long* __restrict__ other,
                                                               const float* __restrict__ interp_w, int H, int W, int L,
                                                               float* __restrict__ level_mask,
                                                               float* __restrict__ level_weight) {
  pdl_sync();
  const int64_t n = (int64_t)H * W, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const int y = (int)(p / W), x = (int)(p % W);
    // neighbourhood: level of the nearest / second-nearest pyramid entry per pixel, -1 where the pixel is masked out
    int r9[9], o9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const int64_t q = (int64_t)yy * W + xx;
      const bool m = in && mask[q] != 0;
      r9[k] = m ? (int)rounded[q] : -1;
      o9[k] = m ? (int)other[q] : -1;
    }
    const float w = interp_w[p];
    for (int l = 0; l < L; ++l) {
      int s_all = 0, s_r = 0, s_o = 0;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int a = (r9[k] == l) ? 1 : 0, b = (o9[k] == l) ? 1 : 0;
        s_all += a | b;          // (m1 + m2) of two bool tensors is their OR in torch
        s_r += a;
        s_o += b;
      }
      const int cr = (r9[4] == l) ? 1 : 0, co = (o9[4] == l) ? 1 : 0;
      const float e_all = (s_all >= 9) ? (float)(cr | co) : 0.f;
      const float e_r = (s_r >= 9) ? (float)cr : 0.f, e_o = (s_o >= 9) ? (float)co : 0.f;
      level_mask[(int64_t)l * n + p] = e_all;
      level_weight[(int64_t)l * n + p] = __fadd_rn(__fmul_rn(e_r, w), __fmul_rn(e_o, __fsub_rn(1.f, w)));
    }
  }
}

struct LevelPlanArgs {
  const float* src_mask;        // [Hr*Wr] the level's mask at rgb resolution (values > 0 select)
  const float* src_weight;      // [Hr*Wr] or nullptr
  const float* angle_guidance;  // [Hr*Wr] or nullptr
  const float* angle_degrees;   // [Hr*Wr] or nullptr (no pass / fail split)
  float threshold;
  int Hr, Wr, H, W;
  float* hook0;                 // [H*W] or nullptr
  float* hook1;                 // [H*W] or nullptr
  int num_layers;
  int lh[SMB_MAX_PLAN_LAYERS], lw[SMB_MAX_PLAN_LAYERS];
  long long seg_begin[SMB_MAX_PLAN_LAYERS + 1];   // flat index of layer k's first pixel (segment 0 = the level itself)
  float* m_all[SMB_MAX_PLAN_LAYERS];
  float* m_pass[SMB_MAX_PLAN_LAYERS];             // nullptr without the split
  float* m_fail[SMB_MAX_PLAN_LAYERS];
  unsigned int* counts;                           // [1 + 3 * num_layers]: alive, then (n, n_pass, n_fail) per layer
};

__global__ void __launch_bounds__(256) view_level_plan_kernel(const LevelPlanArgs a, long long total) {
  pdl_sync();
  __shared__ unsigned int s_cnt[1 + 3 * SMB_MAX_PLAN_LAYERS];
  for (int i = threadIdx.x; i < 1 + 3 * SMB_MAX_PLAN_LAYERS; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long level_px = (long long)a.H * a.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    if (i < level_px) {
      const int y = (int)(i / a.W), x = (int)(i % a.W);
      const int sy = nearest_src(y, a.Hr, a.H), sx = nearest_src(x, a.Wr, a.W);
      const int64_t q = (int64_t)sy * a.Wr + sx;
      if (a.src_mask[q] > 0.f) atomicAdd(&s_cnt[0], 1u);
      if (a.hook0) a.hook0[i] = bilinear_at(a.angle_guidance, a.Hr, a.Wr, a.H, a.W, y, x);
      if (a.hook1) a.hook1[i] = a.src_weight[q];
      continue;
    }
    int k = 0;
    while (k + 1 < a.num_layers && i >= a.seg_begin[k + 1]) ++k;
    const long long j = i - a.seg_begin[k];
    const int yy = (int)(j / a.lw[k]), xx = (int)(j % a.lw[k]);
    const int ny = nearest_src(yy, a.H, a.lh[k]), nx = nearest_src(xx, a.W, a.lw[k]);        // layer pixel -> level pixel
    const int sy = nearest_src(ny, a.Hr, a.H), sx = nearest_src(nx, a.Wr, a.W);            // level pixel -> rgb pixel
    const bool m = a.src_mask[(int64_t)sy * a.Wr + sx] > 0.f;
    a.m_all[k][j] = m ? 1.f : 0.f;
    if (m) atomicAdd(&s_cnt[1 + 3 * k], 1u);
    if (a.m_pass[k]) {
      const bool passed = bilinear_at(a.angle_degrees, a.Hr, a.Wr, a.H, a.W, ny, nx) < a.threshold;
      a.m_pass[k][j] = (m && passed) ? 1.f : 0.f;
      a.m_fail[k][j] = (m && !passed) ? 1.f : 0.f;
      if (m) atomicAdd(&s_cnt[1 + 3 * k + (passed ? 1 : 2)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1 + 3 * a.num_layers; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(a.counts + i, s_cnt[i]);
}

int launch_view_level_masks(const unsigned char* mask, const long long* rounded, const long long* other,
                            const float* interp_w, int H, int W, int L, float* level_mask, float* level_weight,
                            cudaStream_t st) {
  const int64_t n = (int64_t)H * W;
  if (n == 0 || L == 0) return SMB_OK;
  SMB_LAUNCH(view_level_masks_kernel, (unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 8), 256, 0, st, mask,
             rounded, other, interp_w, H, W, L, level_mask, level_weight);
  return SMB_OK;
}

int launch_view_level_plan(const float* src_mask, const float* src_weight, const float* angle_guidance,
                           const float* angle_degrees, float threshold, int Hr, int Wr, int H, int W, float* hook0,
                           float* hook1, int num_layers, const int* lh, const int* lw, float* layer_masks, int split,
                           unsigned int* counts, cudaStream_t st) {
  SMB_REQUIRE(src_mask && counts && num_layers >= 0 && num_layers <= SMB_MAX_PLAN_LAYERS,
              "view_level_plan: need a source mask, a counter block and at most %d layers", SMB_MAX_PLAN_LAYERS);
  SMB_REQUIRE(!hook0 || angle_guidance, "view_level_plan: hook0 needs the angle guidance map");
  SMB_REQUIRE(!hook1 || src_weight, "view_level_plan: hook1 needs the level weight map");
  SMB_REQUIRE(!split || angle_degrees, "view_level_plan: the pass / fail split needs the angle map in degrees");
  SMB_REQUIRE(num_layers == 0 || layer_masks, "view_level_plan: null layer mask buffer");
  LevelPlanArgs a;
  a.src_mask = src_mask;
  a.src_weight = src_weight;
  a.angle_guidance = angle_guidance;
  a.angle_degrees = angle_degrees;
  a.threshold = threshold;
  a.Hr = Hr; a.Wr = Wr; a.H = H; a.W = W;
  a.hook0 = hook0;
  a.hook1 = hook1;
  a.num_layers = num_layers;
  a.counts = counts;
  long long pos = (long long)H * W;
  float* out = layer_masks;
  for (int k = 0; k < SMB_MAX_PLAN_LAYERS; ++k) {
    a.lh[k] = k < num_layers ? lh[k] : 1;
    a.lw[k] = k < num_layers ? lw[k] : 1;
    a.seg_begin[k] = pos;
    a.m_all[k] = a.m_pass[k] = a.m_fail[k] = nullptr;
    if (k < num_layers) {
      const long long px = (long long)lh[k] * lw[k];
      a.m_all[k] = out;
      out += px;
      if (split) {
        a.m_pass[k] = out;
        out += px;
        a.m_fail[k] = out;
        out += px;
      }
      pos += px;
    }
  }
  a.seg_begin[SMB_MAX_PLAN_LAYERS] = pos;
  if (pos == 0) return SMB_OK;
  SMB_CUDA_CHECK(cudaMemsetAsync(counts, 0, (1 + 3 * (size_t)num_layers) * sizeof(unsigned int), st));
  SMB_LAUNCH(view_level_plan_kernel, (unsigned)std::min<long long>((pos + 255) / 256, 148 * 8), 256, 0, st, a, pos);
  return SMB_OK;
}

}  // namespace smb
