// Texture export and headless preview.
//
//   texture_post_rgb8   post() (model/losses/rgb_transform.py:14-21) + ToPILImage's quantisation on the device:
//                       pre()-space BGR fp32 (3,H,W) -> RGB uint8 (H,W,3).  Same fp32 operations in the same order as
//                       the reference chain  x.mul_(1/255) -> Normalize(mean = -m, std = 1) -> [2,1,0] -> clamp(0,1)
//                       -> pic.mul(255).byte()  (model/texture/texture.py:9-19, :123-127), so the bytes are the ones
//                       the reference writes into its epoch_*_texture.jpg (before JPEG coding).
//   mip_downsample2x    one level of the mip chain (2x2 box filter = glGenerateMipmap on a power-of-two texture,
//                       scripts/scannet/render_uv/src/renderer/renderer.cpp:134,139; odd sizes: floor, clamped taps)
//   mip_preview         the reference's post-run "render_mipmap" step (model/optimize.py:181-208 runs the OpenGL
//                       renderer with GL_LINEAR_MIPMAP_LINEAR, renderer.cpp:116-117) without a GL context: the UV maps
//                       of a view already hold (u, v, mip LOD) per pixel (shader/uvmap.frag:8-13), so the styled view
//                       is a trilinear lookup into the mip chain: two bilinear taps (GL texel centres, clamp to edge)
//                       blended by frac(lod + bias); pixels without geometry (u = v = 0) stay black.
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

__device__ __forceinline__ unsigned char post_quantise(float x, float mean) {
  float v = __fmul_rn(x, 1.0f / 255.0f);                 // x.mul_(1. / 255)
  v = __fdiv_rn(__fsub_rn(v, -mean), 1.0f);              // Normalize(mean = -m, std = 1): (x - mean) / std
  v = fminf(fmaxf(v, 0.f), 1.f);                         // clamp(0, 1)
  return (unsigned char)__fmul_rn(v, 255.0f);            // ToPILImage: pic.mul(255).byte()  (truncates)
}

__constant__ float kMeanBGR[3] = {0.40760392f, 0.45795686f, 0.48501961f};

__global__ void __launch_bounds__(256) texture_post_rgb8_kernel(const float* __restrict__ bgr, int64_t npix,
                                                                unsigned char* __restrict__ rgb8) {
  pdl_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
#pragma unroll
    for (int c = 0; c < 3; ++c)                          // output channel c (RGB) = input channel 2 - c (BGR)
      rgb8[3 * p + c] = post_quantise(bgr[(int64_t)(2 - c) * npix + p], kMeanBGR[2 - c]);
  }
}

__global__ void __launch_bounds__(256) mip_downsample2x_kernel(const float* __restrict__ src, int Hs, int Ws,
                                                               float* __restrict__ dst, int Hd, int Wd) {
  pdl_sync();
  const int64_t n = (int64_t)3 * Hd * Wd, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int x = (int)(i % Wd), y = (int)((i / Wd) % Hd), c = (int)(i / ((int64_t)Wd * Hd));
    const int x0 = min(2 * x, Ws - 1), x1 = min(2 * x + 1, Ws - 1), y0 = min(2 * y, Hs - 1), y1 = min(2 * y + 1, Hs - 1);
    const float* s = src + (int64_t)c * Hs * Ws;
    dst[i] = 0.25f * (s[(int64_t)y0 * Ws + x0] + s[(int64_t)y0 * Ws + x1] + s[(int64_t)y1 * Ws + x0] +
                      s[(int64_t)y1 * Ws + x1]);
  }
}

struct MipChain {
  const float* ptr[SMB_MAX_MIP_LEVELS];
  int W[SMB_MAX_MIP_LEVELS], H[SMB_MAX_MIP_LEVELS];
  int n;
};

__device__ __forceinline__ float gl_bilinear(const float* __restrict__ img, int H, int W, float u, float v) {
  // GL_LINEAR, GL_CLAMP_TO_EDGE: texel centres at (i + 0.5) / W
  const float fx = fminf(fmaxf(u * W - 0.5f, 0.f), (float)(W - 1)), fy = fminf(fmaxf(v * H - 0.5f, 0.f), (float)(H - 1));
  const int x0 = (int)fx, y0 = (int)fy, x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  const float ax = fx - x0, ay = fy - y0;
  const float top = img[(int64_t)y0 * W + x0] * (1.f - ax) + img[(int64_t)y0 * W + x1] * ax;
  const float bot = img[(int64_t)y1 * W + x0] * (1.f - ax) + img[(int64_t)y1 * W + x1] * ax;
  return top * (1.f - ay) + bot * ay;
}

__global__ void __launch_bounds__(256) mip_preview_kernel(const MipChain mips, const float* __restrict__ uv,
                                                          int uv_channels, int64_t npix, float lod_bias,
                                                          unsigned char* __restrict__ rgb8) {
  pdl_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
    const float u = uv[uv_channels * p], v = uv[uv_channels * p + 1];
    if (u == 0.f && v == 0.f) {                          // no geometry behind this pixel (the renderer's clear colour)
      rgb8[3 * p] = rgb8[3 * p + 1] = rgb8[3 * p + 2] = 0;
      continue;
    }
    float lod = (uv_channels > 2 ? uv[uv_channels * p + 2] : 0.f) + lod_bias;
    lod = fminf(fmaxf(lod, 0.f), (float)(mips.n - 1));
    const int l0 = (int)lod, l1 = min(l0 + 1, mips.n - 1);
    const float t = lod - l0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                        // output channel c (RGB) = texture channel 2 - c (BGR)
      const int ch = 2 - c;
      const float a = gl_bilinear(mips.ptr[l0] + (int64_t)ch * mips.H[l0] * mips.W[l0], mips.H[l0], mips.W[l0], u, v);
      const float b = gl_bilinear(mips.ptr[l1] + (int64_t)ch * mips.H[l1] * mips.W[l1], mips.H[l1], mips.W[l1], u, v);
      rgb8[3 * p + c] = post_quantise(a * (1.f - t) + b * t, kMeanBGR[ch]);
    }
  }
}

static inline unsigned grid_for(int64_t n) {
  return (unsigned)std::min<int64_t>(std::max<int64_t>(ceil_div64(n, 256), 1), 148 * 8);
}

int launch_texture_post_rgb8(const float* bgr_chw, int H, int W, unsigned char* rgb_hwc, cudaStream_t st) {
  const int64_t n = (int64_t)H * W;
  if (n == 0) return SMB_OK;
  SMB_LAUNCH(texture_post_rgb8_kernel, grid_for(n), 256, 0, st, bgr_chw, n, rgb_hwc);
  return SMB_OK;
}

int launch_mip_downsample2x(const float* src, int Hs, int Ws, float* dst, cudaStream_t st) {
  const int Hd = std::max(Hs / 2, 1), Wd = std::max(Ws / 2, 1);
  SMB_LAUNCH(mip_downsample2x_kernel, grid_for((int64_t)3 * Hd * Wd), 256, 0, st, src, Hs, Ws, dst, Hd, Wd);
  return SMB_OK;
}

int launch_mip_preview(const float* const* mips, const int* mW, const int* mH, int num_mips, const float* uv,
                       int uv_channels, int H, int W, float lod_bias, unsigned char* rgb_hwc, cudaStream_t st) {
  SMB_REQUIRE(num_mips >= 1 && num_mips <= SMB_MAX_MIP_LEVELS, "mip_preview: 1..%d mip levels", SMB_MAX_MIP_LEVELS);
  SMB_REQUIRE(uv_channels == 2 || uv_channels == 3, "mip_preview: uv maps have 2 (u, v) or 3 (u, v, lod) channels");
  MipChain c;
  for (int i = 0; i < SMB_MAX_MIP_LEVELS; ++i) {
    c.ptr[i] = i < num_mips ? mips[i] : nullptr;
    c.W[i] = i < num_mips ? mW[i] : 1;
    c.H[i] = i < num_mips ? mH[i] : 1;
    SMB_REQUIRE(i >= num_mips || (c.ptr[i] && c.W[i] > 0 && c.H[i] > 0), "mip_preview: bad mip level %d", i);
  }
  c.n = num_mips;
  const int64_t n = (int64_t)H * W;
  if (n == 0) return SMB_OK;
  SMB_LAUNCH(mip_preview_kernel, grid_for(n), 256, 0, st, c, uv, uv_channels, n, lod_bias, rgb_hwc);
  return SMB_OK;
}

}  // namespace smb
