// View preparation kernels: what Abstract_Dataset.__getitem__ (data/abstract_dataset.py:270-344) does to the raw
// arrays of one view on DataLoader workers, as device kernels feeding a GPU-resident view store.
//
//   uv_grid        uv (H,W,3) -> grid (H,W,2) = fl(fl(2 uv) - 1)  + valid = (u != 0) | (v != 0) [& depth > 0]
//                  (model/texture/utils.py:6-8,56-60; data/scannet_dataset.py:308-326)
//   gather2d       nearest resampling with host-built index tables (cv2 INTER_NEAREST for the angle map,
//                  PIL NEAREST for the mask: data/abstract_dataset.py:306-311)
//   resize_linear  cv2 INTER_LINEAR with host-built offset / weight tables, double arithmetic for sensor depth
//                  (data/abstract_dataset.py:301-304, data/scannet_dataset.py:319-320)
//   depth_levels   calculate_depth_level (data/scannet_dataset.py:328-366), the numpy float64 arithmetic operation by
//                  operation (no FMA contraction: every product / sum is rounded like numpy rounds it)
//   rgb_pre        ToTensor + pre() (model/losses/rgb_transform.py:5-11): uint8 HWC -> fp32 CHW, BGR, -mean, x255
//   angle_degrees  rad2deg(acos(cos)) (data/abstract_dataset.py:338)
//   erode3x3       the 3x3 erosion of model/model.py:204-208 (keep x where the zero-padded 3x3 box mean is exactly 1)
//
// These run once per view (a few hundred KB each), so they are written for exactness, not for the roofline: plain
// grid-stride kernels, all arithmetic through explicitly rounded intrinsics.
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

static inline unsigned blocks_for(int64_t n, int threads = 256) {
  return (unsigned)std::min<int64_t>(std::max<int64_t>(ceil_div64(n, threads), 1), 148 * 8);
}

__global__ void __launch_bounds__(256) view_uv_grid_kernel(const float* __restrict__ uv3, int64_t n,
                                                           float* __restrict__ grid2, unsigned char* __restrict__ valid,
                                                           const double* __restrict__ depth) {
  pdl_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const float u = uv3[3 * p], v = uv3[3 * p + 1];
    grid2[2 * p] = __fsub_rn(__fmul_rn(u, 2.0f), 1.0f);
    grid2[2 * p + 1] = __fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
    if (valid) {
      bool m = (u != 0.f) || (v != 0.f);
      if (depth) m = m && (depth[p] > 0.0);
      valid[p] = m ? 1 : 0;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) view_gather2d_kernel(const T* __restrict__ src, int Ws,
                                                            const int* __restrict__ ytab, const int* __restrict__ xtab,
                                                            int Hd, int Wd, T* __restrict__ dst) {
  pdl_sync();
  const int64_t n = (int64_t)Hd * Wd, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const int y = (int)(p / Wd), x = (int)(p % Wd);
    dst[p] = src[(int64_t)ytab[y] * Ws + xtab[x]];
  }
}

struct LinearTables {
  const int* yofs;
  const double* ya;
  const int* xofs;
  const double* xa;
};

// src value in the work type: uint16 / divisor (the reference divides the PNG by 1000.0 or 4000.0 in float64)
template <typename TIn>
__device__ __forceinline__ double load_depth(const TIn* src, int64_t i, double divisor) {
  return (double)src[i];
}
template <>
__device__ __forceinline__ double load_depth<unsigned short>(const unsigned short* src, int64_t i, double divisor) {
  return __ddiv_rn((double)src[i], divisor);
}

// WORK = double: rows = s[x0] * (1 - xa) + s[x1] * xa ; out = rows[y0] * (1 - ya) + rows[y1] * ya, each op rounded
template <typename TIn, bool WORK_F32>
__global__ void __launch_bounds__(256) view_resize_linear_kernel(const TIn* __restrict__ src, double divisor, int Hs,
                                                                 int Ws, LinearTables t, int Hd, int Wd,
                                                                 double* __restrict__ dst) {
  pdl_sync();
  const int64_t n = (int64_t)Hd * Wd, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const int y = (int)(p / Wd), x = (int)(p % Wd);
    const int y0 = t.yofs[y], x0 = t.xofs[x];
    const int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
    if (WORK_F32) {
      const float ax = (float)t.xa[x], ay = (float)t.ya[y];
      const float bx = __fsub_rn(1.f, ax), by = __fsub_rn(1.f, ay);
      const int64_t o0 = (int64_t)y0 * Ws, o1 = (int64_t)y1 * Ws;
      const float r0 = __fadd_rn(__fmul_rn((float)src[o0 + x0], bx), __fmul_rn((float)src[o0 + x1], ax));
      const float r1 = __fadd_rn(__fmul_rn((float)src[o1 + x0], bx), __fmul_rn((float)src[o1 + x1], ax));
      dst[p] = (double)__fadd_rn(__fmul_rn(r0, by), __fmul_rn(r1, ay));
    } else {
      const double ax = t.xa[x], ay = t.ya[y];
      const double bx = __dsub_rn(1.0, ax), by = __dsub_rn(1.0, ay);
      const int64_t o0 = (int64_t)y0 * Ws, o1 = (int64_t)y1 * Ws;
      const double s00 = load_depth<TIn>(src, o0 + x0, divisor), s01 = load_depth<TIn>(src, o0 + x1, divisor);
      const double s10 = load_depth<TIn>(src, o1 + x0, divisor), s11 = load_depth<TIn>(src, o1 + x1, divisor);
      const double r0 = __dadd_rn(__dmul_rn(s00, bx), __dmul_rn(s01, ax));
      const double r1 = __dadd_rn(__dmul_rn(s10, bx), __dmul_rn(s11, ax));
      dst[p] = __dadd_rn(__dmul_rn(r0, by), __dmul_rn(r1, ay));
    }
  }
}

// plain conversion to double (no resampling needed: the source already has the target size)
template <typename TIn>
__global__ void __launch_bounds__(256) view_to_f64_kernel(const TIn* __restrict__ src, double divisor, int64_t n,
                                                          double* __restrict__ dst) {
  pdl_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    dst[p] = load_depth<TIn>(src, p, divisor);
}

constexpr int VP_MAX_LEVELS = 16;
struct LevelTable {
  double v[VP_MAX_LEVELS];
  int n;
};

__global__ void __launch_bounds__(256) view_depth_levels_kernel(const double* __restrict__ depth, int64_t n,
                                                                const LevelTable lv, double min_depth, int depth_is_f32,
                                                                float* __restrict__ cont, float* __restrict__ depth_f32,
                                                                long long* __restrict__ rounded,
                                                                long long* __restrict__ other,
                                                                float* __restrict__ weight) {
  pdl_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const double d = depth[p];
    if (depth_f32) depth_f32[p] = (float)d;                            // transform_label: to_numpy -> float32
    // uv_height = 32 * (depth / min_depth); rendered (float32) depth keeps numpy's float32 arithmetic here
    double uvh;
    if (depth_is_f32) uvh = (double)__fmul_rn(32.f, __fdiv_rn((float)d, (float)min_depth));
    else uvh = __dmul_rn(32.0, __ddiv_rn(d, min_depth));
    int r = 0;
    double best = fabs(__dsub_rn(uvh, lv.v[0]));
    for (int k = 1; k < lv.n; ++k) {                                   // np.argmin: first minimum
      const double a = fabs(__dsub_rn(uvh, lv.v[k]));
      if (a < best) {
        best = a;
        r = k;
      }
    }
    const double res = __dsub_rn(lv.v[r], uvh);
    int o = r + (res > 0.0 ? -1 : (res == 0.0 ? 0 : 1));
    o = max(0, min(o, lv.n - 1));
    const double hd = fabs(__dsub_rn(lv.v[r], lv.v[o]));
    double w = fabs(__ddiv_rn(res, __dadd_rn(hd, 1e-6)));
    if (hd == 0.0) w = 0.0;
    w = __dsub_rn(1.0, w);
    double c = (res > 0.0) ? __dadd_rn((double)o, w) : __dsub_rn((double)o, w);
    if (w == 1.0) c = (double)r;
    cont[p] = (float)c;
    rounded[p] = r;
    other[p] = o;
    weight[p] = (float)w;
  }
}

__global__ void __launch_bounds__(256) view_rgb_pre_kernel(const unsigned char* __restrict__ hwc, int64_t P,
                                                           float* __restrict__ chw) {
  pdl_sync();
  const float mean_bgr[3] = {0.40760392f, 0.45795686f, 0.48501961f};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                      // output channel c = input channel 2 - c (BGR)
      const float x = __fdiv_rn((float)hwc[3 * p + (2 - c)], 255.f);
      chw[(int64_t)c * P + p] = __fmul_rn(__fdiv_rn(__fsub_rn(x, mean_bgr[c]), 1.f), 255.f);
    }
  }
}

__global__ void __launch_bounds__(256) view_angle_degrees_kernel(const float* __restrict__ c, int64_t n,
                                                                 float* __restrict__ deg) {
  pdl_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    deg[p] = __fmul_rn(acosf(c[p]), 57.29577951308232f);               // torch.rad2deg: x * (180 / pi)
}

// erode (model/model.py:204-208): keep x where conv2d(x, ones(3,3), padding=1) / 9 clamped to [0,1] equals 1
__global__ void __launch_bounds__(256) view_erode3x3_kernel(const float* __restrict__ x, int H, int W,
                                                            float* __restrict__ out) {
  pdl_sync();
  const int64_t n = (int64_t)H * W, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const int y = (int)(p / W), xx = (int)(p % W);
    float s = 0.f;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, x2 = xx + dx;
        if (yy >= 0 && yy < H && x2 >= 0 && x2 < W) s = __fadd_rn(s, x[(int64_t)yy * W + x2]);
      }
    const float m = fminf(fmaxf(__fdiv_rn(s, 9.f), 0.f), 1.f);
    out[p] = (m == 1.f) ? x[p] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------
int launch_view_uv_grid(const float* uv3, int H, int W, float* grid2, unsigned char* valid, const double* depth,
                        cudaStream_t st) {
  const int64_t n = (int64_t)H * W;
  if (n == 0) return SMB_OK;
  SMB_LAUNCH(view_uv_grid_kernel, blocks_for(n), 256, 0, st, uv3, n, grid2, valid, depth);
  return SMB_OK;
}

int launch_view_gather2d(const void* src, int elem_bytes, int Ws, const int* ytab, const int* xtab, int Hd, int Wd,
                         void* dst, cudaStream_t st) {
  const int64_t n = (int64_t)Hd * Wd;
  if (n == 0) return SMB_OK;
  if (elem_bytes == 1) {
    SMB_LAUNCH(view_gather2d_kernel<unsigned char>, blocks_for(n), 256, 0, st, (const unsigned char*)src, Ws, ytab, xtab,
               Hd, Wd, (unsigned char*)dst);
  } else if (elem_bytes == 4) {
    SMB_LAUNCH(view_gather2d_kernel<unsigned int>, blocks_for(n), 256, 0, st, (const unsigned int*)src, Ws, ytab, xtab, Hd,
               Wd, (unsigned int*)dst);
  } else {
    set_error("view_gather2d: element size %d not supported (1 or 4 bytes)", elem_bytes);
    return SMB_ERR_ARG;
  }
  return SMB_OK;
}

int launch_view_resize_linear(const void* src, int src_type, double divisor, int Hs, int Ws, const int* yofs,
                              const double* ya, const int* xofs, const double* xa, int Hd, int Wd, double* dst,
                              cudaStream_t st) {
  const int64_t n = (int64_t)Hd * Wd;
  if (n == 0) return SMB_OK;
  const bool same = (Hs == Hd && Ws == Wd);
  LinearTables t{yofs, ya, xofs, xa};
  switch (src_type) {
    case 0:
      if (same) SMB_LAUNCH(view_to_f64_kernel<double>, blocks_for(n), 256, 0, st, (const double*)src, divisor, n, dst);
      else SMB_LAUNCH((view_resize_linear_kernel<double, false>), blocks_for(n), 256, 0, st, (const double*)src,
          divisor, Hs, Ws, t, Hd, Wd, dst);
      break;
    case 1:
      if (same) SMB_LAUNCH(view_to_f64_kernel<float>, blocks_for(n), 256, 0, st, (const float*)src, divisor, n, dst);
      else SMB_LAUNCH((view_resize_linear_kernel<float, true>), blocks_for(n), 256, 0, st, (const float*)src, divisor,
          Hs, Ws, t, Hd, Wd, dst);
      break;
    case 2:
      if (same) SMB_LAUNCH(view_to_f64_kernel<unsigned short>, blocks_for(n), 256, 0, st, (const unsigned short*)src,
          divisor, n, dst);
      else SMB_LAUNCH((view_resize_linear_kernel<unsigned short, false>), blocks_for(n), 256, 0, st,
          (const unsigned short*)src, divisor, Hs, Ws, t, Hd, Wd, dst);
      break;
    default:
      set_error("view_resize_linear: src_type %d (0 = float64, 1 = float32, 2 = uint16 / divisor)", src_type);
      return SMB_ERR_ARG;
  }
  return SMB_OK;
}

int launch_view_depth_levels(const double* depth, int64_t n, const double* levels_host, int num_levels, double min_depth,
                             int depth_is_f32, float* cont, float* depth_f32, long long* rounded, long long* other,
                             float* weight, cudaStream_t st) {
  SMB_REQUIRE(num_levels >= 1 && num_levels <= VP_MAX_LEVELS, "view_depth_levels: 1..%d pyramid levels", VP_MAX_LEVELS);
  SMB_REQUIRE(min_depth > 0.0, "view_depth_levels: min_depth must be positive");
  if (n == 0) return SMB_OK;
  LevelTable lv;
  lv.n = num_levels;
  for (int i = 0; i < VP_MAX_LEVELS; ++i) lv.v[i] = i < num_levels ? levels_host[i] : 0.0;
  SMB_LAUNCH(view_depth_levels_kernel, blocks_for(n), 256, 0, st, depth, n, lv, min_depth, depth_is_f32, cont, depth_f32,
             rounded, other, weight);
  return SMB_OK;
}

int launch_view_rgb_pre(const unsigned char* hwc, int H, int W, float* chw, cudaStream_t st) {
  const int64_t P = (int64_t)H * W;
  if (P == 0) return SMB_OK;
  SMB_LAUNCH(view_rgb_pre_kernel, blocks_for(P), 256, 0, st, hwc, P, chw);
  return SMB_OK;
}

int launch_view_angle_degrees(const float* c, int64_t n, float* deg, cudaStream_t st) {
  if (n == 0) return SMB_OK;
  SMB_LAUNCH(view_angle_degrees_kernel, blocks_for(n), 256, 0, st, c, n, deg);
  return SMB_OK;
}

int launch_view_erode3x3(const float* x, int H, int W, float* out, cudaStream_t st) {
  const int64_t n = (int64_t)H * W;
  if (n == 0) return SMB_OK;
  SMB_LAUNCH(view_erode3x3_kernel, blocks_for(n), 256, 0, st, x, H, W, out);
  return SMB_OK;
}

}  // namespace smb
