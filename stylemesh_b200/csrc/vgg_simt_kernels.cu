// CUDA-core (SIMT) kernels of the VGG/loss side: the 3-channel first layer and its data gradient, max-pool
// forward/backward, layout converters, masked Gram (validation path), Gram-MSE, content-MSE, and a register-tiled
// fp32 implicit-GEMM that serves as the on-device cross-check of the tcgen05 implicit-GEMM (tc_igemm.cu).
//
// Reference behaviour restated (never copied):
//   model/losses/content_and_style_losses.py:47-70   VGG.forward: relu(conv3x3 pad 1), MaxPool2d(2,2) floor mode
//   model/losses/content_and_style_losses.py:74-80   GramMatrix: bmm(F, F^T) / (h*w)   (h*w == Nvalid after :136-143)
//   model/losses/content_and_style_losses.py:325-348 w * f * MSE(target, current)
#include "smb_common.cuh"
#include "smb_epilogue.cuh"
#include "smb_kernels.h"

namespace smb {

// ============================================================================================================
// layout converters
// ============================================================================================================
__global__ void __launch_bounds__(256) act_from_nchw_kernel(const float* __restrict__ src, Act dst) {
  pdl_sync();
  // one thread per (pixel, 8-channel group); reads are coalesced across pixels for each channel
  const int64_t P = dst.pixels();
  const int groups = dst.C >> 3;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * groups) return;
  const int64_t p = idx % P;
  const int g = (int)(idx / P);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (int64_t)(g * 8 + j) * P + p);
  uint4 h, l;
  split2_pack(v[0], v[1], h.x, l.x);
  split2_pack(v[2], v[3], h.y, l.y);
  split2_pack(v[4], v[5], h.z, l.z);
  split2_pack(v[6], v[7], h.w, l.w);
  *reinterpret_cast<uint4*>(dst.hi + p * dst.C + g * 8) = h;
  *reinterpret_cast<uint4*>(dst.lo + p * dst.C + g * 8) = l;
}

__global__ void __launch_bounds__(256) act_to_nchw_kernel(Act src, float* __restrict__ dst) {
  pdl_sync();
  const int64_t P = src.pixels();
  const int groups = src.C >> 3;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * groups) return;
  const int64_t p = idx % P;
  const int g = (int)(idx / P);
  const uint4 h = *reinterpret_cast<const uint4*>(src.hi + p * src.C + g * 8);
  const uint4 l = *reinterpret_cast<const uint4*>(src.lo + p * src.C + g * 8);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    dst[(int64_t)(g * 8 + 2 * j) * P + p] = bf16lo_to_f(hw[j]) + bf16lo_to_f(lw[j]);
    dst[(int64_t)(g * 8 + 2 * j + 1) * P + p] = bf16hi_to_f(hw[j]) + bf16hi_to_f(lw[j]);
  }
}

__global__ void __launch_bounds__(256) mask_rows_kernel(Act src, const float* __restrict__ rowmask, Act dst) {
  pdl_sync();
  const int64_t n8 = src.elems() >> 3;
  const int g_per_row = src.C >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float m = __ldg(rowmask + i / g_per_row);
    uint4 h = reinterpret_cast<const uint4*>(src.hi)[i];
    uint4 l = reinterpret_cast<const uint4*>(src.lo)[i];
    if (m == 0.f) {
      h = make_uint4(0, 0, 0, 0);
      l = make_uint4(0, 0, 0, 0);
    } else if (m != 1.f) {
      uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = (bf16lo_to_f(hw[j]) + bf16lo_to_f(lw[j])) * m;
        const float b = (bf16hi_to_f(hw[j]) + bf16hi_to_f(lw[j])) * m;
        split2_pack(a, b, hw[j], lw[j]);
      }
      h = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      l = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
    reinterpret_cast<uint4*>(dst.hi)[i] = h;
    reinterpret_cast<uint4*>(dst.lo)[i] = l;
  }
}

__global__ void __launch_bounds__(256) relu_mask_split_kernel(const float* __restrict__ g, Act y, Act dz) {
  pdl_sync();
  const int64_t n4 = y.elems() >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    const uint2 s = __ldg(reinterpret_cast<const uint2*>(y.hi) + i);
    float v[4] = {gv.x, gv.y, gv.z, gv.w};
    const uint32_t e[4] = {s.x & 0xffffu, s.x >> 16, s.y & 0xffffu, s.y >> 16};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (!((e[j] & 0x8000u) == 0 && (e[j] & 0x7fffu) != 0)) v[j] = 0.f;
    uint2 h, l;
    split2_pack(v[0], v[1], h.x, l.x);
    split2_pack(v[2], v[3], h.y, l.y);
    reinterpret_cast<uint2*>(dz.hi)[i] = h;
    reinterpret_cast<uint2*>(dz.lo)[i] = l;
  }
}

// ============================================================================================================
// first layer: conv1_1 (3 -> Cout=64) straight from the fp32 planar image; one thread per pixel
// ============================================================================================================
template <int COUT>
__global__ void __launch_bounds__(128) conv_first_fwd_kernel(const float* __restrict__ img, int H, int W,
                                                             const float* __restrict__ w_oihw, Epilogue ep) {
  pdl_sync();
  __shared__ float sw[27][COUT];   // [ci*9 + r*3 + s][co]
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
    const int co = i / 27, k = i % 27;   // w_oihw[co][ci][r][s] is contiguous in k = ci*9+r*3+s
    sw[k][co] = w_oihw[i];
  }
  __syncthreads();
  const int64_t P = (int64_t)H * W;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int y = (int)(p / W), x = (int)(p % W);
  float in[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yy = y + r - 1, xx = x + s - 1;
        in[ci * 9 + r * 3 + s] =
            (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (int64_t)ci * P + (int64_t)yy * W + xx) : 0.f;
      }
#pragma unroll 1
  for (int c0 = 0; c0 < COUT; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(&sw[k][c0 + j]);
        acc[j] = fmaf(in[k], wv.x, acc[j]);
        acc[j + 1] = fmaf(in[k], wv.y, acc[j + 1]);
        acc[j + 2] = fmaf(in[k], wv.z, acc[j + 2]);
        acc[j + 3] = fmaf(in[k], wv.w, acc[j + 3]);
      }
    }
    epilogue_store<16>(ep, p, c0, COUT, acc);
  }
}

// data gradient of the first layer: dimg[ci][p] = sum_{r,s,co} dz(y-(r-1), x-(s-1))[co] * w[co][ci][r][s]
template <int COUT>
__global__ void __launch_bounds__(128) conv_first_dgrad_kernel(Act dz, const float* __restrict__ w_oihw,
                                                               float* __restrict__ dimg) {
  pdl_sync();
  __shared__ float4 sw[9][COUT];    // [r*3+s][co] = (w[co][0][r][s], w[co][1][r][s], w[co][2][r][s], 0)
  for (int i = threadIdx.x; i < 9 * COUT; i += blockDim.x) {
    const int rs = i / COUT, co = i % COUT;
    sw[rs][co] = make_float4(w_oihw[co * 27 + rs], w_oihw[co * 27 + 9 + rs], w_oihw[co * 27 + 18 + rs], 0.f);
  }
  __syncthreads();
  const int H = dz.H, W = dz.W;
  const int64_t P = (int64_t)H * W;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int y = (int)(p / W), x = (int)(p % W);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 1
  for (int rs = 0; rs < 9; ++rs) {
    const int r = rs / 3, s = rs % 3;
    const int yy = y - (r - 1), xx = x - (s - 1);
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const int64_t q = ((int64_t)yy * W + xx) * COUT;
#pragma unroll 2
    for (int c8 = 0; c8 < COUT; c8 += 8) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(dz.hi + q + c8));
      const uint4 l = __ldg(reinterpret_cast<const uint4*>(dz.lo + q + c8));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v0 = bf16lo_to_f(hw[j]) + bf16lo_to_f(lw[j]);
        const float v1 = bf16hi_to_f(hw[j]) + bf16hi_to_f(lw[j]);
        const float4 w0 = sw[rs][c8 + 2 * j];
        const float4 w1 = sw[rs][c8 + 2 * j + 1];
        a0 = fmaf(v0, w0.x, a0); a1 = fmaf(v0, w0.y, a1); a2 = fmaf(v0, w0.z, a2);
        a0 = fmaf(v1, w1.x, a0); a1 = fmaf(v1, w1.y, a1); a2 = fmaf(v1, w1.z, a2);
      }
    }
  }
  dimg[p] = a0;
  dimg[P + p] = a1;
  dimg[2 * P + p] = a2;
}

// ============================================================================================================
// max-pool 2x2 stride 2 (floor mode)
// ============================================================================================================
__device__ __forceinline__ void load8(const Act& a, int64_t off, float (&v)[8], uint4& h, uint4& l) {
  h = __ldg(reinterpret_cast<const uint4*>(a.hi + off));
  l = __ldg(reinterpret_cast<const uint4*>(a.lo + off));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = bf16lo_to_f(hw[j]) + bf16lo_to_f(lw[j]);
    v[2 * j + 1] = bf16hi_to_f(hw[j]) + bf16hi_to_f(lw[j]);
  }
}

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(Act in, Act out) {
  pdl_sync();
  const int groups = in.C >> 3;
  const int64_t total = out.pixels() * groups;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const int64_t po = idx / groups;
  const int yo = (int)(po / out.W), xo = (int)(po % out.W);
  float best[8];
  uint16_t bh[8], bl[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = 2 * yo + (k >> 1), xx = 2 * xo + (k & 1);
    float v[8];
    uint4 h, l;
    load8(in, ((int64_t)yy * in.W + xx) * in.C + g * 8, v, h, l);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint16_t eh = (j & 1) ? (uint16_t)(hw[j >> 1] >> 16) : (uint16_t)(hw[j >> 1] & 0xffffu);
      const uint16_t el = (j & 1) ? (uint16_t)(lw[j >> 1] >> 16) : (uint16_t)(lw[j >> 1] & 0xffffu);
      if (k == 0 || v[j] > best[j]) {
        best[j] = v[j];
        bh[j] = eh;
        bl[j] = el;
      }
    }
  }
  uint4 h, l;
  h.x = bh[0] | ((uint32_t)bh[1] << 16); h.y = bh[2] | ((uint32_t)bh[3] << 16);
  h.z = bh[4] | ((uint32_t)bh[5] << 16); h.w = bh[6] | ((uint32_t)bh[7] << 16);
  l.x = bl[0] | ((uint32_t)bl[1] << 16); l.y = bl[2] | ((uint32_t)bl[3] << 16);
  l.z = bl[4] | ((uint32_t)bl[5] << 16); l.w = bl[6] | ((uint32_t)bl[7] << 16);
  *reinterpret_cast<uint4*>(out.hi + po * out.C + g * 8) = h;
  *reinterpret_cast<uint4*>(out.lo + po * out.C + g * 8) = l;
}

// one thread per (2x2 window, 8-channel group): reads the four y pixels once, writes the four dz pixels.
// Windows hanging over an odd trailing row/column carry no pooled gradient (floor mode) but their pixels still get
// dz = addend masked by ReLU (or zero).
__global__ void __launch_bounds__(256, 3) maxpool_bwd_relu_kernel(const float* __restrict__ gp,
                                                               const float* __restrict__ addend, Act y, Act dz) {
  pdl_sync();
  const int groups = y.C >> 3;
  const int Ho = y.H >> 1, Wo = y.W >> 1;
  const int Hc = (y.H + 1) >> 1, Wc = (y.W + 1) >> 1;
  const int64_t total = (int64_t)Hc * Wc * groups;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const int64_t wq = idx / groups;
  const int wy = (int)(wq / Wc), wx = (int)(wq % Wc);
  const bool pooled = (wy < Ho) && (wx < Wo);
  float v[4][8];
  bool exists[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = 2 * wy + (k >> 1), xx = 2 * wx + (k & 1);
    exists[k] = (yy < y.H) && (xx < y.W);
    if (exists[k]) {
      uint4 h, l;
      load8(y, ((int64_t)yy * y.W + xx) * y.C + g * 8, v[k], h, l);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = -1.f;
    }
  }
  float gv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gv[j] = 0.f;
  if (pooled) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp + ((int64_t)wy * Wo + wx) * y.C + g * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gp + ((int64_t)wy * Wo + wx) * y.C + g * 8 + 4));
    gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w;
    gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
  }
  int arg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    arg[j] = 0;
    float best = v[0][j];
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (v[k][j] > best) { best = v[k][j]; arg[j] = k; }      // first maximum in window scan order wins
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!exists[k]) continue;
    const int yy = 2 * wy + (k >> 1), xx = 2 * wx + (k & 1);
    const int64_t off = ((int64_t)yy * y.W + xx) * y.C + g * 8;
    float out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = 0.f;
    if (addend) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(addend + off));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(addend + off + 4));
      out[0] = a0.x; out[1] = a0.y; out[2] = a0.z; out[3] = a0.w;
      out[4] = a1.x; out[5] = a1.y; out[6] = a1.z; out[7] = a1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float val = out[j] + ((pooled && arg[j] == k) ? gv[j] : 0.f);
      out[j] = (v[k][j] > 0.f) ? val : 0.f;
    }
    uint4 h, l;
    split2_pack(out[0], out[1], h.x, l.x);
    split2_pack(out[2], out[3], h.y, l.y);
    split2_pack(out[4], out[5], h.z, l.z);
    split2_pack(out[6], out[7], h.w, l.w);
    *reinterpret_cast<uint4*>(dz.hi + off) = h;
    *reinterpret_cast<uint4*>(dz.lo + off) = l;
  }
}

// ============================================================================================================
// fp32 SIMT implicit GEMM (cross-check path): tile = 8x8 pixels x 64 outputs, K chunks of 16
// ============================================================================================================
constexpr int SG_TP = 8;          // patch edge
constexpr int SG_BN = 64;
constexpr int SG_KC = 16;
constexpr int SG_HALO = SG_TP + 2;

__global__ void __launch_bounds__(128) igemm_simt_kernel(Act a, PackedB b, Epilogue ep, int tiles_x) {
  pdl_sync();
  __shared__ float As[SG_HALO * SG_HALO][SG_KC + 1];
  __shared__ float Ws[9][SG_BN][SG_KC + 1];
  const int tid = threadIdx.x;
  const int tn = tid & 15;        // 4 outputs: n0 + tn*4 ..
  const int tm = tid >> 4;        // patch row 0..7
  const int ty0 = (blockIdx.x / tiles_x) * SG_TP, tx0 = (blockIdx.x % tiles_x) * SG_TP;
  const int n0 = blockIdx.y * SG_BN;
  const int H = a.H, W = a.W, K = b.K, N = b.N, taps = b.taps;

  float acc[SG_TP][4];
#pragma unroll
  for (int i = 0; i < SG_TP; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_KC) {
    __syncthreads();
    // ---- stage A halo patch (zero outside the image) ----
    for (int i = tid; i < SG_HALO * SG_HALO * 2; i += blockDim.x) {
      const int pix = i >> 1, half = i & 1;
      const int yy = ty0 + pix / SG_HALO - 1, xx = tx0 + pix % SG_HALO - 1;
      float v[8];
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        uint4 h, l;
        load8(a, ((int64_t)yy * W + xx) * a.C + k0 + half * 8, v, h, l);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) As[pix][half * 8 + j] = v[j];
    }
    // ---- stage B chunk ----
    for (int i = tid; i < taps * SG_BN * 2; i += blockDim.x) {
      const int half = i & 1, n = (i >> 1) % SG_BN, tap = (i >> 1) / SG_BN;
      const int64_t off = ((int64_t)tap * N + n0 + n) * K + k0 + half * 8;
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(b.hi + off));
      const uint4 l = __ldg(reinterpret_cast<const uint4*>(b.lo + off));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Ws[tap][n][half * 8 + 2 * j] = bf16lo_to_f(hw[j]) + bf16lo_to_f(lw[j]);
        Ws[tap][n][half * 8 + 2 * j + 1] = bf16hi_to_f(hw[j]) + bf16hi_to_f(lw[j]);
      }
    }
    __syncthreads();
    // ---- multiply ----
    if (taps == 9) {
#pragma unroll 1
      for (int r = 0; r < 3; ++r) {
#pragma unroll 4
        for (int k = 0; k < SG_KC; ++k) {
          float av[SG_HALO];
#pragma unroll
          for (int x = 0; x < SG_HALO; ++x) av[x] = As[(tm + r) * SG_HALO + x][k];
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = Ws[r * 3 + s][tn * 4 + j][k];
#pragma unroll
            for (int x = 0; x < SG_TP; ++x)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[x][j] = fmaf(av[x + s], wv[j], acc[x][j]);
          }
        }
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < SG_KC; ++k) {
        float wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = Ws[0][tn * 4 + j][k];
#pragma unroll
        for (int x = 0; x < SG_TP; ++x) {
          const float av = As[(tm + 1) * SG_HALO + x + 1][k];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[x][j] = fmaf(av, wv[j], acc[x][j]);
        }
      }
    }
  }
  // ---- epilogue ----
  const int yy = ty0 + tm;
  if (yy < H) {
#pragma unroll
    for (int x = 0; x < SG_TP; ++x) {
      const int xx = tx0 + x;
      if (xx < W) epilogue_store<4>(ep, (int64_t)yy * W + xx, n0 + tn * 4, N, acc[x]);
    }
  }
}

// ============================================================================================================
// masked Gram (SIMT cross-check): partial[s][i][j] = sum_{p in split s} Fm[p][i] Fm[p][j]
// ============================================================================================================
constexpr int GS_T = 32;   // output tile edge and pixel chunk
__global__ void __launch_bounds__(256) gram_simt_kernel(Act fm, float* __restrict__ partial, int nsplit,
                                                        int64_t pix_per_split) {
  pdl_sync();
  __shared__ float Fi[GS_T][GS_T + 1], Fj[GS_T][GS_T + 1];
  const int C = fm.C;
  const int i0 = blockIdx.x * GS_T, j0 = blockIdx.y * GS_T, s = blockIdx.z;
  const int64_t P = fm.pixels();
  const int64_t pbeg = (int64_t)s * pix_per_split, pend = min(P, pbeg + pix_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16x16 threads, 2x2 outputs each
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int64_t p0 = pbeg; p0 < pend; p0 += GS_T) {
    __syncthreads();
    for (int i = threadIdx.x; i < GS_T * GS_T; i += blockDim.x) {
      const int pp = i / GS_T, c = i % GS_T;
      const int64_t p = p0 + pp;
      float vi = 0.f, vj = 0.f;
      if (p < pend) {
        vi = merge2(fm.hi[p * C + i0 + c], fm.lo[p * C + i0 + c]);
        vj = merge2(fm.hi[p * C + j0 + c], fm.lo[p * C + j0 + c]);
      }
      Fi[pp][c] = vi;
      Fj[pp][c] = vj;
    }
    __syncthreads();
#pragma unroll 8
    for (int pp = 0; pp < GS_T; ++pp) {
      const float a0 = Fi[pp][ty * 2], a1 = Fi[pp][ty * 2 + 1];
      const float b0 = Fj[pp][tx * 2], b1 = Fj[pp][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
  }
  float* out = partial + (int64_t)s * C * C;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) out[(int64_t)(i0 + ty * 2 + a) * C + j0 + tx * 2 + bb] = acc[a][bb];
}

// ============================================================================================================
// Gram-MSE: reduce split partials, normalise, (optional running average), loss, gradient seed matrix
// ============================================================================================================
struct GramMseArgs {
  const float* partial;
  int nsplit, C;
  float inv_n;
  const float* y0;
  float coef0;
  const float* y1;
  float coef1;
  const float* prev_sum;
  float avg_len;
  float* g_out;
  __nv_bfloat16 *b_hi, *b_lo;
  float* loss_out;
};

// element e of the Gram matrix once its split partials are summed: outputs + this element's share of the loss
__device__ __forceinline__ float gram_mse_finish(const GramMseArgs& a, int64_t e, float sum, float inv_cc) {
  const float inv_len = 1.f / a.avg_len;
  const float g = sum * a.inv_n;
  if (a.g_out) a.g_out[e] = g;
  float ghat = g;
  if (a.prev_sum) ghat = (g + __ldg(a.prev_sum + e)) * inv_len;
  float d = 0.f, lacc = 0.f;
  {
    const float diff = ghat - __ldg(a.y0 + e);
    lacc = fmaf(a.coef0 * diff, diff, lacc);
    d = fmaf(a.coef0, diff, d);
  }
  if (a.y1) {
    const float diff = ghat - __ldg(a.y1 + e);
    lacc = fmaf(a.coef1 * diff, diff, lacc);
    d = fmaf(a.coef1, diff, d);
  }
  // dL/dGhat = 2 d / C^2 ; gradient seed for dF = Fm * Bmat:  Bmat = (2 inv_n / len) * dL/dGhat
  const float bm = (2.f * a.inv_n * inv_len) * (2.f * d * inv_cc);
  __nv_bfloat16 h, l;
  split2(bm, h, l);
  a.b_hi[e] = h;
  a.b_lo[e] = l;
  return lacc;
}

// Many splits, few elements (C <= 128: up to 148 splits of 4096 / 16384 elements).
__global__ void __launch_bounds__(256) gram_mse_kernel(const GramMseArgs a) {
  pdl_sync();
  // block = 32 consecutive Gram elements x 8 split groups: group sg sums partial[sg], partial[sg+8], ... (coalesced
  // 128-byte rows), the groups are combined in a fixed order => deterministic, and no thread walks all splits alone
  __shared__ float s_acc[8][33];
  const int64_t CC = (int64_t)a.C * a.C;
  const float inv_cc = 1.f / (float)CC;
  const int lane = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (e < CC) {
    // eight loads in flight per thread, added in the original order (a plain loop waits one L2 round trip per split)
    const float* src = a.partial + e;
    int s = sg;
    for (; s + 56 < a.nsplit; s += 64) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(src + (int64_t)(s + 8 * k) * CC);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += v[k];
    }
    for (; s < a.nsplit; s += 8) acc += __ldg(src + (int64_t)s * CC);
  }
  s_acc[sg][lane] = acc;
  __syncthreads();
  if (sg != 0) return;
  float lacc = 0.f;
  if (e < CC) {
    float g = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) g += s_acc[k][lane];
    lacc = gram_mse_finish(a, e, g, inv_cc);
  }
  lacc = warp_sum(lacc);
  if (lane == 0) atomicAdd(a.loss_out, lacc * inv_cc);
}

// Few splits, many elements (C >= 256: at most 18 splits of 65536 / 262144 elements): one thread per element walks
// the splits in order, grid-stride over the matrix, ONE loss atomic per block (the grouped kernel above would issue
// 8192 same-address atomics at C = 512, which serialise in L2).
__global__ void __launch_bounds__(256) gram_mse_wide_kernel(const GramMseArgs a) {
  pdl_sync();
  const int64_t CC = (int64_t)a.C * a.C;
  const float inv_cc = 1.f / (float)CC;
  float lacc = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < CC; e += (int64_t)gridDim.x * 256) {
    const float* src = a.partial + e;
    float acc = 0.f;
    int s = 0;
    for (; s + 8 <= a.nsplit; s += 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(src + (int64_t)(s + k) * CC);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += v[k];
    }
    for (; s < a.nsplit; ++s) acc += __ldg(src + (int64_t)s * CC);
    lacc += gram_mse_finish(a, e, acc, inv_cc);
  }
  lacc = block_sum(lacc);
  if (threadIdx.x == 0) atomicAdd(a.loss_out, lacc * inv_cc);
}

// ============================================================================================================
// content MSE on a masked feature map
// ============================================================================================================
__global__ void __launch_bounds__(256) content_mse_kernel(Act f, const float* __restrict__ target,
                                                          const float* __restrict__ rowmask, float coef_loss,
                                                          float coef_grad, float* __restrict__ addend,
                                                          float* __restrict__ loss_out) {
  pdl_sync();
  const int64_t n8 = f.elems() >> 3;
  const int g_per_row = f.C >> 3;
  float lacc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float m = rowmask ? __ldg(rowmask + i / g_per_row) : 1.f;
    float v[8];
    uint4 h, l;
    load8(f, i * 8, v, h, l);
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(target + i * 8));
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(target + i * 8 + 4));
    const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    float d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d[j] = (v[j] - t[j]) * m;
      lacc = fmaf(d[j], d[j], lacc);
      d[j] *= coef_grad;
    }
    float4* ad = reinterpret_cast<float4*>(addend + i * 8);
    const float4 o0 = ad[0], o1 = ad[1];   // accumulate: several terms may target the same layer
    ad[0] = make_float4(o0.x + d[0], o0.y + d[1], o0.z + d[2], o0.w + d[3]);
    ad[1] = make_float4(o1.x + d[4], o1.y + d[5], o1.z + d[6], o1.w + d[7]);
  }
  lacc = block_sum(lacc);
  if (threadIdx.x == 0) atomicAdd(loss_out, lacc * coef_loss);
}

// ============================================================================================================
// host launchers
// ============================================================================================================
static int grid_for(int64_t work, int threads, int cap = 148 * 32) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(work, threads), cap));
}

int launch_act_from_nchw(const float* src, const Act& dst, cudaStream_t st) {
  SMB_REQUIRE(dst.C % 8 == 0, "act_from_nchw: C=%d must be a multiple of 8", dst.C);
  const int64_t work = dst.pixels() * (dst.C >> 3);
  if (work == 0) return SMB_OK;
  SMB_LAUNCH(act_from_nchw_kernel, (unsigned)ceil_div64(work, 256), 256, 0, st, src, dst);
  return SMB_OK;
}
int launch_act_to_nchw(const Act& src, float* dst, cudaStream_t st) {
  SMB_REQUIRE(src.C % 8 == 0, "act_to_nchw: C=%d must be a multiple of 8", src.C);
  const int64_t work = src.pixels() * (src.C >> 3);
  if (work == 0) return SMB_OK;
  SMB_LAUNCH(act_to_nchw_kernel, (unsigned)ceil_div64(work, 256), 256, 0, st, src, dst);
  return SMB_OK;
}
int launch_mask_rows(const Act& src, const float* rowmask, const Act& dst, cudaStream_t st) {
  if (src.elems() == 0) return SMB_OK;
  SMB_LAUNCH(mask_rows_kernel, grid_for(src.elems() >> 3, 256), 256, 0, st, src, rowmask, dst);
  return SMB_OK;
}
int launch_relu_mask_split(const float* g, const Act& y, const Act& dz, cudaStream_t st) {
  if (y.elems() == 0) return SMB_OK;
  SMB_LAUNCH(relu_mask_split_kernel, grid_for(y.elems() >> 2, 256), 256, 0, st, g, y, dz);
  return SMB_OK;
}

int launch_conv_first_fwd(const float* img, int H, int W, const float* w_oihw, const float* bias, int Cout,
                          const Epilogue& ep_in, cudaStream_t st) {
  SMB_REQUIRE(Cout == 64, "conv_first_fwd: only Cout=64 (VGG conv1_1) is built, got %d", Cout);
  Epilogue ep = ep_in;
  ep.bias = bias;
  const int64_t P = (int64_t)H * W;
  if (P == 0) return SMB_OK;
  SMB_LAUNCH(conv_first_fwd_kernel<64>, (unsigned)ceil_div64(P, 128), 128, 0, st, img, H, W, w_oihw, ep);
  return SMB_OK;
}
int launch_conv_first_dgrad(const Act& dz, const float* w_oihw, int Cout, float* dimg, cudaStream_t st) {
  SMB_REQUIRE(Cout == 64 && dz.C == 64, "conv_first_dgrad: only Cout=64 is built");
  const int64_t P = dz.pixels();
  if (P == 0) return SMB_OK;
  SMB_LAUNCH(conv_first_dgrad_kernel<64>, (unsigned)ceil_div64(P, 128), 128, 0, st, dz, w_oihw, dimg);
  return SMB_OK;
}

int launch_maxpool_fwd(const Act& in, const Act& out, cudaStream_t st) {
  SMB_REQUIRE(out.H == in.H / 2 && out.W == in.W / 2 && out.C == in.C && in.C % 8 == 0, "maxpool_fwd: bad shapes");
  const int64_t work = out.pixels() * (in.C >> 3);
  if (work == 0) return SMB_OK;
  SMB_LAUNCH(maxpool_fwd_kernel, (unsigned)ceil_div64(work, 256), 256, 0, st, in, out);
  return SMB_OK;
}
int launch_maxpool_bwd_relu(const float* g_pooled, const float* addend, const Act& y, const Act& dz,
                            cudaStream_t st) {
  SMB_REQUIRE(y.C % 8 == 0 && dz.C == y.C && dz.H == y.H && dz.W == y.W, "maxpool_bwd: bad shapes");
  const int64_t work = (int64_t)((y.H + 1) / 2) * ((y.W + 1) / 2) * (y.C >> 3);
  if (work == 0) return SMB_OK;
  SMB_LAUNCH(maxpool_bwd_relu_kernel, (unsigned)ceil_div64(work, 256), 256, 0, st, g_pooled, addend, y, dz);
  return SMB_OK;
}

int launch_igemm_simt(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st) {
  SMB_REQUIRE(b.taps == 9 || b.taps == 1, "igemm: taps must be 1 or 9");
  SMB_REQUIRE(a.C == b.K && b.K % SG_KC == 0 && b.N % SG_BN == 0, "igemm_simt: K=%d (mult of 16), N=%d (mult of 64)",
              b.K, b.N);
  if (a.pixels() == 0) return SMB_OK;
  const int tiles_x = ceil_div(a.W, SG_TP), tiles_y = ceil_div(a.H, SG_TP);
  dim3 grid(tiles_x * tiles_y, b.N / SG_BN);
  SMB_LAUNCH(igemm_simt_kernel, grid, 128, 0, st, a, b, ep, tiles_x);
  return SMB_OK;
}

int gram_tc_bn(int C);   // tc_kernels.cu

int gram_num_splits(int64_t P, int C, int impl) {
  if (P <= 0) return 1;
  // tcgen05 kernel: one CTA per SM (192 KB ring), so ONE wave of <= 148 CTAs - a second wave pays prologue and
  // pipeline fill again; a split is a whole number of 64-pixel stages and none is empty
  const int tiles = (impl == IMPL_TC) ? ((C + 127) / 128) * (C / gram_tc_bn(C)) : (C / GS_T) * (C / GS_T);
  const int64_t want = std::max(1, (impl == IMPL_TC ? 148 : 148 * 2) / std::max(1, tiles));
  const int64_t stages = ceil_div64(P, 64);
  const int64_t per = ceil_div64(stages, std::min(want, stages));
  return (int)ceil_div64(stages, per);
}

int launch_gram_simt(const Act& fm, float* partial, int nsplit, cudaStream_t st) {
  SMB_REQUIRE(fm.C % GS_T == 0, "gram_simt: C=%d must be a multiple of 32", fm.C);
  const int64_t P = fm.pixels();
  const int64_t pps = ceil_div64(ceil_div64(std::max<int64_t>(P, 1), 64), nsplit) * 64;
  dim3 grid(fm.C / GS_T, fm.C / GS_T, nsplit);
  SMB_LAUNCH(gram_simt_kernel, grid, 256, 0, st, fm, partial, nsplit, pps);
  return SMB_OK;
}

int launch_gram_mse(const float* partial, int nsplit, int C, float inv_n, const float* y0, float coef0,
                    const float* y1, float coef1, const float* prev_sum, float avg_len, float* g_out,
                    __nv_bfloat16* b_hi, __nv_bfloat16* b_lo, float* loss_out, cudaStream_t st) {
  SMB_REQUIRE(y0 != nullptr && avg_len >= 1.f, "gram_mse: need a target and avg_len >= 1");
  const int64_t CC = (int64_t)C * C;
  GramMseArgs a{partial, nsplit, C, inv_n, y0, coef0, y1, coef1, prev_sum, avg_len, g_out, b_hi, b_lo, loss_out};
  if (nsplit <= 32 && CC >= 65536) {
    SMB_LAUNCH(gram_mse_wide_kernel, (unsigned)std::min<int64_t>(ceil_div64(CC, 256), 148 * 4), 256, 0, st, a);
  } else {
    SMB_LAUNCH(gram_mse_kernel, (unsigned)ceil_div64(CC, 32), 256, 0, st, a);
  }
  return SMB_OK;
}

int launch_content_mse(const Act& f, const float* target_nhwc, const float* rowmask, float coef_loss,
                       float coef_grad, float* addend, float* loss_out, cudaStream_t st) {
  if (f.elems() == 0) return SMB_OK;
  SMB_LAUNCH(content_mse_kernel,
             grid_for(f.elems() >> 3, 256, 148 * 8), 256, 0, st, f, target_nhwc, rowmask, coef_loss, coef_grad, addend, loss_out);
  return SMB_OK;
}

}  // namespace smb
