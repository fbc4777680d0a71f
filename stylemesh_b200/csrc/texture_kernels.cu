// Texture-side kernels: hierarchical bilinear UV sample (forward), UV scatter-add (backward, with the
// angle/depth gradient-hook weights folded in), fused clamp + L2-regulariser + Adam update, regulariser value.
//
// Reference behaviour restated here (never copied):
//   model/texture/texture.py:41-54   NeuralTexture.normalize()/forward(): clamp to [-123.68,151.061], then
//                                    F.grid_sample(bilinear, padding_mode='border', align_corners=True)
//   model/texture/texture.py:96-108  HierarchicalNeuralTexture.forward() = sum over layers; regularizer()
//   model/model.py:198-202,247-251   gradient hooks  grad *= bilinear(angle_guidance), grad *= depth weight
//   model/model.py:387-401           torch.optim.Adam(lr, betas (0.9,0.999), eps 1e-8, wd 0)
// The coordinate arithmetic follows ATen/native/GridSampler.h (grid_sampler_unnormalize, clip_coordinates,
// within_bounds_2d) operation by operation in fp32 with explicit *_rn intrinsics so that no FMA contraction can
// change an index:  ix = ((gx + 1) / 2) * (W - 1);  ix = min(W-1, max(ix, 0));  x0 = floor(ix).
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

struct TexCoord {
  int x0, y0;          // north-west texel
  float w_nw, w_ne, w_sw, w_se;
};

__device__ __forceinline__ TexCoord uv_to_texel(float gx, float gy, int W, int H) {
  // grid_sampler_unnormalize(align_corners=True): ((coord + 1) / 2) * (size - 1)
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(H - 1));
  // clip_coordinates (padding_mode=border): min(size-1, max(in, 0))
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  const float fx = floorf(ix), fy = floorf(iy);
  TexCoord t;
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  const float x1 = __fadd_rn(fx, 1.f), y1 = __fadd_rn(fy, 1.f);
  const float ax = __fsub_rn(x1, ix), bx = __fsub_rn(ix, fx);
  const float ay = __fsub_rn(y1, iy), by = __fsub_rn(iy, fy);
  t.w_nw = __fmul_rn(ax, ay);
  t.w_ne = __fmul_rn(bx, ay);
  t.w_sw = __fmul_rn(ax, by);
  t.w_se = __fmul_rn(bx, by);
  return t;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ------------------------------------------------------------------------------------------
// forward: out[c][p] = sum_l sum_corner w * clamp(tex_l[c][y][x])      one thread per pixel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) uv_sample_fwd_kernel(TexLayerSet tex, const float2* __restrict__ grid,
                                                            int npix, float clamp_lo, float clamp_hi,
                                                            float* __restrict__ out) {
  pdl_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const float2 g = __ldg(grid + p);
  float acc[SMB_MAX_TEX_CHANNELS];
#pragma unroll
  for (int c = 0; c < SMB_MAX_TEX_CHANNELS; ++c) acc[c] = 0.f;
  for (int l = 0; l < tex.L; ++l) {
    const int W = tex.W[l], H = tex.H[l];
    const TexCoord t = uv_to_texel(g.x, g.y, W, H);
    const bool x1ok = (t.x0 + 1) < W, y1ok = (t.y0 + 1) < H;   // within_bounds_2d for the +1 corners
    const size_t plane = (size_t)W * H;
    const float* base = tex.ptr[l] + (size_t)t.y0 * W + t.x0;
#pragma unroll
    for (int c = 0; c < SMB_MAX_TEX_CHANNELS; ++c) {
      if (c < tex.C) {
        const float* b = base + c * plane;
        float v = 0.f;
        v = fmaf(clampf(__ldg(b), clamp_lo, clamp_hi), t.w_nw, v);
        if (x1ok) v = fmaf(clampf(__ldg(b + 1), clamp_lo, clamp_hi), t.w_ne, v);
        if (y1ok) v = fmaf(clampf(__ldg(b + W), clamp_lo, clamp_hi), t.w_sw, v);
        if (x1ok && y1ok) v = fmaf(clampf(__ldg(b + W + 1), clamp_lo, clamp_hi), t.w_se, v);
        acc[c] += v;   // torch.stack(...).sum(0): layer results are added in layer order
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SMB_MAX_TEX_CHANNELS; ++c)
    if (c < tex.C) out[(size_t)c * npix + p] = acc[c];
}

// debug/export: the exact integer texel and the four fp32 weights (for the bit-exact index test)
__global__ void uv_texel_index_kernel(const float2* __restrict__ grid, int npix, int W, int H,
                                      int* __restrict__ xy0, float* __restrict__ w4) {
  pdl_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const float2 g = grid[p];
  const TexCoord t = uv_to_texel(g.x, g.y, W, H);
  xy0[2 * p] = t.x0;
  xy0[2 * p + 1] = t.y0;
  w4[4 * p] = t.w_nw;
  w4[4 * p + 1] = t.w_ne;
  w4[4 * p + 2] = t.w_sw;
  w4[4 * p + 3] = t.w_se;
}

// ------------------------------------------------------------------------------------------
// backward: gtex_l[c][y][x] += w_corner * (gout[c][p] * hook0[p] * hook1[p])
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) uv_scatter_bwd_kernel(TexLayerSet gtex, const float2* __restrict__ grid,
                                                             int npix, const float* __restrict__ gout,
                                                             const float* __restrict__ hook0,
                                                             const float* __restrict__ hook1) {
  pdl_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = p < npix;
  float g[SMB_MAX_TEX_CHANNELS];
  // hooks run in registration order: angle first (model.py:195-202), then depth (model.py:245-251);
  // autograd applies them most-recent-last == same order, one fp32 multiply each.
  bool any = false;
#pragma unroll
  for (int c = 0; c < SMB_MAX_TEX_CHANNELS; ++c) {
    g[c] = 0.f;
    if (in_range && c < gtex.C) {
      float v = __ldg(gout + (size_t)c * npix + p);
      if (hook0) v = __fmul_rn(v, __ldg(hook0 + p));
      if (hook1) v = __fmul_rn(v, __ldg(hook1 + p));
      g[c] = v;
      any |= (v != 0.f);
    }
  }
  float2 uv = make_float2(0.f, 0.f);
  if (in_range) uv = __ldg(grid + p);
  // Warp aggregation: every invalid pixel carries uv == (-1,-1) (model/texture/utils.py:6-8 on uv == 0) and scatters
  // into texel (0,0) of every layer — ~10 % of a view hammering 12 addresses serialises the L2 atomic unit
  // (measured 298 us of a 3.8 ms step).  When all contributing lanes of a warp share one uv, sum in registers
  // and let one lane issue the atomics.
  const unsigned full = 0xffffffffu;
  const unsigned amask = __ballot_sync(full, any);
  if (amask == 0u) return;
  const int src = __ffs(amask) - 1;
  const float ux = __shfl_sync(full, uv.x, src), uy = __shfl_sync(full, uv.y, src);
  const bool same = !any || (__float_as_uint(uv.x) == __float_as_uint(ux) && __float_as_uint(uv.y) == __float_as_uint(uy));
  const bool uniform = __all_sync(full, same) && (__popc(amask) > 1);
  if (uniform) {
#pragma unroll
    for (int c = 0; c < SMB_MAX_TEX_CHANNELS; ++c) g[c] = warp_sum(any ? g[c] : 0.f);
    if ((int)(threadIdx.x & 31) != src) return;
    uv = make_float2(ux, uy);
  } else if (!any) {
    return;
  }
  for (int l = 0; l < gtex.L; ++l) {
    const int W = gtex.W[l], H = gtex.H[l];
    const TexCoord t = uv_to_texel(uv.x, uv.y, W, H);
    const bool x1ok = (t.x0 + 1) < W, y1ok = (t.y0 + 1) < H;
    const size_t plane = (size_t)W * H;
    float* base = gtex.ptr[l] + (size_t)t.y0 * W + t.x0;
#pragma unroll
    for (int c = 0; c < SMB_MAX_TEX_CHANNELS; ++c) {
      if (c < gtex.C && g[c] != 0.f) {
        float* b = base + c * plane;
        atomicAdd(b, __fmul_rn(t.w_nw, g[c]));
        if (x1ok) atomicAdd(b + 1, __fmul_rn(t.w_ne, g[c]));
        if (y1ok) atomicAdd(b + W, __fmul_rn(t.w_sw, g[c]));
        if (x1ok && y1ok) atomicAdd(b + W + 1, __fmul_rn(t.w_se, g[c]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused Adam:  x = clamp(p);  g' = g*gscale + reg*x;  m,v update;  p = x - step;  g = 0
// (torch/optim/adam.py single-tensor path: m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g,g,1-b2);
//  denom = sqrt(v)/sqrt(bc2) + eps; p.addcdiv_(m, denom, value=-lr/bc1))
// ------------------------------------------------------------------------------------------
struct AdamScalars {
  float one_minus_b1, b2, one_minus_b2, inv_sqrt_bc2, eps, neg_step;
  float clamp_lo, clamp_hi, reg_coef, gscale;
};

__device__ __forceinline__ void adam_elem(float& p, float& g, float& m, float& v, const AdamScalars& s) {
  const float x = clampf(p, s.clamp_lo, s.clamp_hi);
  const float gg = fmaf(s.reg_coef, x, g * s.gscale);
  m = fmaf(s.one_minus_b1, gg - m, m);
  v = fmaf(s.one_minus_b2 * gg, gg, v * s.b2);
  const float denom = sqrtf(v) * s.inv_sqrt_bc2 + s.eps;
  p = fmaf(s.neg_step, m / denom, x);
  g = 0.f;
}

__global__ void __launch_bounds__(256) adam_clamp_reg_kernel(float* __restrict__ p, float* __restrict__ g,
                                                             float* __restrict__ m, float* __restrict__ v,
                                                             int64_t n, AdamScalars s) {
  pdl_sync();
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    adam_elem(P.x, G.x, M.x, V.x, s);
    adam_elem(P.y, G.y, M.y, V.y, s);
    adam_elem(P.z, G.z, M.z, V.z, s);
    adam_elem(P.w, G.w, M.w, V.w, s);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(g)[i] = G;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  // tail (n not a multiple of 4)
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_elem(p[i], g[i], m[i], v[i], s);
}

// The same update over a flat buffer of up to SMB_MAX_TEX_LAYERS consecutive segments (one texture layer each, every
// segment starting at a multiple of 4 floats) in ONE launch: segment l covers [begin[l], begin[l+1]) and differs only
// in its regulariser coefficient.  Four launches of 59 + 13 + 6 + 5 us become one at the HBM roofline.
struct SegTable {
  int64_t begin[SMB_MAX_TEX_LAYERS + 1];
  float coef[SMB_MAX_TEX_LAYERS];
  int n;
};
__device__ __forceinline__ float seg_coef(const SegTable& t, int64_t i) {
  float c = t.coef[0];
#pragma unroll
  for (int l = 1; l < SMB_MAX_TEX_LAYERS; ++l)
    if (l < t.n && i >= t.begin[l]) c = t.coef[l];
  return c;
}

__global__ void __launch_bounds__(256) adam_clamp_reg_seg_kernel(float* __restrict__ p, float* __restrict__ g,
                                                                 float* __restrict__ m, float* __restrict__ v,
                                                                 int64_t n, AdamScalars s, const SegTable seg) {
  pdl_sync();
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    s.reg_coef = seg_coef(seg, i << 2);
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    adam_elem(P.x, G.x, M.x, V.x, s);
    adam_elem(P.y, G.y, M.y, V.y, s);
    adam_elem(P.z, G.z, M.z, V.z, s);
    adam_elem(P.w, G.w, M.w, V.w, s);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(g)[i] = G;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    s.reg_coef = seg_coef(seg, i);
    adam_elem(p[i], g[i], m[i], v[i], s);
  }
}

// out += sum_l coef_l * sum_{i in segment l} clamp(x_i)^2
__global__ void __launch_bounds__(256) sumsq_clamped_seg_kernel(const float* __restrict__ x, int64_t n, float clamp_lo,
                                                                float clamp_hi, const SegTable seg,
                                                                float* __restrict__ out) {
  pdl_sync();
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 X = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float a = clampf(X.x, clamp_lo, clamp_hi), b = clampf(X.y, clamp_lo, clamp_hi);
    const float c = clampf(X.z, clamp_lo, clamp_hi), d = clampf(X.w, clamp_lo, clamp_hi);
    acc = fmaf(seg_coef(seg, i << 2), fmaf(a, a, fmaf(b, b, fmaf(c, c, d * d))), acc);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float c = clampf(x[i], clamp_lo, clamp_hi);
    acc = fmaf(seg_coef(seg, i) * c, c, acc);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// ------------------------------------------------------------------------------------------
// Multi-GPU step: gradient reduce-scatter + Adam + parameter all-gather in ONE kernel over NVLink peer memory.
//
// Every rank holds a full texture replica and a full local gradient (views are sharded over ranks); the
// NCCL version all-reduces the 66.8 MB gradient (2 (N-1)/N x 66.8 MB per rank on the wire, 0.19 ms at N = 2) and then
// every rank runs the same 470 MB Adam pass.  Here rank r owns the r-th 1/N slice of the flat buffer:
//   A  all ranks have finished their backward (release/acquire flags in each other's memory, system scope);
//   B  g = sum_p grad_p[slice] read straight from the peers (fixed order: deterministic), Adam on the slice with this
//      rank's shard of the moments, the new texels are stored into EVERY rank's parameter buffer;
//   C  the last block tells the peers that this rank has finished reading their gradients / writing their texels.
// dist_adam_finish_kernel then waits for all peers' C flags (now nobody reads this rank's gradient or writes its
// texels any more) and zeroes the local gradient for the next step.  Wire traffic per rank: (N-1)/N x 66.8 MB in
// and out, Adam traffic and moments memory divided by N, replicas bit-identical by construction (one writer per texel).
// ------------------------------------------------------------------------------------------
constexpr int DIST_MAX_WORLD = 16;
struct DistArgs {
  const float* grad[DIST_MAX_WORLD];     // every rank's flat gradient (peer pointers; [rank] is local)
  float* param[DIST_MAX_WORLD];          // every rank's flat parameters
  unsigned int* flags[DIST_MAX_WORLD];   // every rank's flag block: [0..15] ready(src), [16..31] done(src), [32] block counter
  float* m;
  float* v;
  int64_t n4;                            // float4 elements in the flat buffer
  int rank, world;
  unsigned int epoch;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// thread p < world waits until rank p has published `epoch` in this rank's flag slot.  Ranks may be seconds apart at
// the first step (one-off style-target pass, allocations), so the watchdog is generous: ~30 s, then trap
__device__ __forceinline__ void dist_wait_all(const unsigned int* my_flags, int world, unsigned int epoch, int what) {
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(my_flags + threadIdx.x) - epoch) < 0) {
      __nanosleep(200);
      if (clock64() - t0 > 60000000000LL) {
        printf("[smb] dist_adam watchdog: phase %d, block %d waiting for rank %d (epoch %u)\n", what, (int)blockIdx.x,
               (int)threadIdx.x, epoch);
        asm volatile("trap;");
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) dist_adam_kernel(const DistArgs a, AdamScalars s, const SegTable seg) {
  pdl_sync();
  // ---- A: this rank's gradient is complete (the launch follows the scatter kernel in stream order) -> tell everyone
  if (blockIdx.x == 0 && (int)threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.rank, a.epoch);
  }
  dist_wait_all(a.flags[a.rank], a.world, a.epoch, 0);
  // ---- B: this rank's slice
  const int64_t per = (a.n4 + a.world - 1) / a.world;
  const int64_t lo = (int64_t)a.rank * per, hi = (lo + per < a.n4) ? lo + per : a.n4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int p = 0; p < a.world; ++p) {            // fixed order; peer memory is read past L1
      const float4 t = __ldcg(reinterpret_cast<const float4*>(a.grad[p]) + i);
      G.x += t.x; G.y += t.y; G.z += t.z; G.w += t.w;
    }
    s.reg_coef = seg_coef(seg, i << 2);
    float4 P = reinterpret_cast<float4*>(a.param[a.rank])[i];
    float4 M = reinterpret_cast<float4*>(a.m)[i], V = reinterpret_cast<float4*>(a.v)[i];
    adam_elem(P.x, G.x, M.x, V.x, s);
    adam_elem(P.y, G.y, M.y, V.y, s);
    adam_elem(P.z, G.z, M.z, V.z, s);
    adam_elem(P.w, G.w, M.w, V.w, s);
    reinterpret_cast<float4*>(a.m)[i] = M;
    reinterpret_cast<float4*>(a.v)[i] = V;
    for (int p = 0; p < a.world; ++p) reinterpret_cast<float4*>(a.param[p])[i] = P;
  }
  // ---- C: last block of this rank -> "done" to everyone
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) {
    unsigned int* ctr = a.flags[a.rank] + 32;
    s_last = (atomicAdd(ctr, 1u) == gridDim.x - 1) ? 1u : 0u;
    if (s_last) *ctr = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + 16 + a.rank, a.epoch);
  }
}

__global__ void __launch_bounds__(256) dist_adam_finish_kernel(const DistArgs a, float* __restrict__ grad_local) {
  pdl_sync();
  dist_wait_all(a.flags[a.rank] + 16, a.world, a.epoch, 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += stride)
    reinterpret_cast<float4*>(grad_local)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// regulariser value:  out += coef * sum(clamp(x)^2)     (coef = lambda * w_l / N_l)
__global__ void __launch_bounds__(256) sumsq_clamped_kernel(const float* __restrict__ x, int64_t n, float coef,
                                                            float clamp_lo, float clamp_hi,
                                                            float* __restrict__ out) {
  pdl_sync();
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float c = clampf(x[i], clamp_lo, clamp_hi);
    acc = fmaf(c, c, acc);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc * coef);
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
int launch_uv_sample_fwd(const TexLayerSet& tex, const float* grid, int H, int W, float clamp_lo, float clamp_hi,
                         float* out, cudaStream_t st) {
  const int npix = H * W;
  if (npix == 0) return SMB_OK;
  SMB_LAUNCH(uv_sample_fwd_kernel, ceil_div(npix, 256), 256, 0, st, tex, reinterpret_cast<const float2*>(grid), npix,
             clamp_lo, clamp_hi, out);
  return SMB_OK;
}

int launch_uv_texel_index(const float* grid, int npix, int W, int H, int* xy0, float* w4, cudaStream_t st) {
  if (npix == 0) return SMB_OK;
  SMB_LAUNCH(uv_texel_index_kernel, ceil_div(npix, 256), 256, 0, st, reinterpret_cast<const float2*>(grid), npix, W, H,
             xy0, w4);
  return SMB_OK;
}

int launch_uv_scatter_bwd(const TexLayerSet& gtex, const float* grid, int H, int W, const float* gout,
                          const float* hook0, const float* hook1, cudaStream_t st) {
  const int npix = H * W;
  if (npix == 0) return SMB_OK;
  SMB_LAUNCH(uv_scatter_bwd_kernel, ceil_div(npix, 256), 256, 0, st, gtex, reinterpret_cast<const float2*>(grid), npix,
             gout, hook0, hook1);
  return SMB_OK;
}

static AdamScalars adam_scalars(float lr, float beta1, float beta2, float eps, int step, float clamp_lo,
                                float clamp_hi, float reg_coef, float gscale) {
  // scalar prep in double, as python floats are in torch/optim/adam.py
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  AdamScalars s;
  s.one_minus_b1 = (float)(1.0 - (double)beta1);
  s.b2 = beta2;
  s.one_minus_b2 = (float)(1.0 - (double)beta2);
  s.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  s.eps = eps;
  s.neg_step = (float)(-(double)lr / bc1);
  s.clamp_lo = clamp_lo;
  s.clamp_hi = clamp_hi;
  s.reg_coef = reg_coef;
  s.gscale = gscale;
  return s;
}

int launch_adam(float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                int step, float clamp_lo, float clamp_hi, float reg_coef, float gscale, cudaStream_t st) {
  if (n == 0) return SMB_OK;
  SMB_REQUIRE(step >= 1, "adam: step must be >= 1 (got %d)", step);
  const AdamScalars s = adam_scalars(lr, beta1, beta2, eps, step, clamp_lo, clamp_hi, reg_coef, gscale);
  const int64_t n4 = (n + 3) >> 2;
  int blocks = (int)std::min<int64_t>(ceil_div64(n4, 256), (int64_t)148 * 16);
  SMB_LAUNCH(adam_clamp_reg_kernel, blocks, 256, 0, st, p, g, m, v, n, s);
  return SMB_OK;
}

static int fill_segments(SegTable* t, int64_t n, const int64_t* begin, const float* coef, int num) {
  SMB_REQUIRE(begin && coef && num >= 1 && num <= SMB_MAX_TEX_LAYERS, "segments: need 1..%d segments", SMB_MAX_TEX_LAYERS);
  SMB_REQUIRE(begin[0] == 0, "segments: the first segment must start at 0");
  for (int l = 0; l < SMB_MAX_TEX_LAYERS; ++l) {
    t->begin[l] = l < num ? begin[l] : n;
    t->coef[l] = l < num ? coef[l] : 0.f;
    if (l > 0 && l < num)
      SMB_REQUIRE(begin[l] >= begin[l - 1] && begin[l] <= n && begin[l] % 4 == 0,
                  "segments: offsets must ascend, stay inside the buffer and be multiples of 4 (segment %d)", l);
  }
  t->begin[SMB_MAX_TEX_LAYERS] = n;
  t->n = num;
  return SMB_OK;
}

int launch_adam_segments(float* p, float* g, float* m, float* v, int64_t n, const int64_t* seg_begin,
                         const float* seg_reg_coef, int num_segments, float lr, float beta1, float beta2, float eps,
                         int step, float clamp_lo, float clamp_hi, float gscale, cudaStream_t st) {
  if (n == 0) return SMB_OK;
  SMB_REQUIRE(step >= 1, "adam: step must be >= 1 (got %d)", step);
  SegTable seg;
  int rc = fill_segments(&seg, n, seg_begin, seg_reg_coef, num_segments);
  if (rc) return rc;
  const AdamScalars s = adam_scalars(lr, beta1, beta2, eps, step, clamp_lo, clamp_hi, 0.f, gscale);
  int blocks = (int)std::min<int64_t>(ceil_div64((n + 3) >> 2, 256), (int64_t)148 * 16);
  SMB_LAUNCH(adam_clamp_reg_seg_kernel, blocks, 256, 0, st, p, g, m, v, n, s, seg);
  return SMB_OK;
}

int launch_dist_adam(int rank, int world, float* const* grad_ptrs, float* const* param_ptrs,
                     unsigned int* const* flag_ptrs, float* m, float* v, int64_t n, const int64_t* seg_begin,
                     const float* seg_reg_coef, int num_segments, float lr, float beta1, float beta2, float eps,
                     int step, float clamp_lo, float clamp_hi, unsigned int epoch, cudaStream_t st) {
  SMB_REQUIRE(world >= 2 && world <= DIST_MAX_WORLD && rank >= 0 && rank < world, "dist_adam: rank %d of %d", rank, world);
  SMB_REQUIRE(step >= 1 && n > 0 && n % 4 == 0, "dist_adam: step >= 1 and a flat buffer of a multiple of 4 floats");
  SegTable seg;
  int rc = fill_segments(&seg, n, seg_begin, seg_reg_coef, num_segments);
  if (rc) return rc;
  DistArgs a;
  for (int p = 0; p < DIST_MAX_WORLD; ++p) {
    a.grad[p] = p < world ? grad_ptrs[p] : nullptr;
    a.param[p] = p < world ? param_ptrs[p] : nullptr;
    a.flags[p] = p < world ? flag_ptrs[p] : nullptr;
    SMB_REQUIRE(p >= world || (a.grad[p] && a.param[p] && a.flags[p]), "dist_adam: null peer pointer for rank %d", p);
  }
  a.m = m;
  a.v = v;
  a.n4 = n >> 2;
  a.rank = rank;
  a.world = world;
  a.epoch = epoch;
  const AdamScalars s = adam_scalars(lr, beta1, beta2, eps, step, clamp_lo, clamp_hi, 0.f, 1.f / (float)world);
  const int64_t per = ceil_div64(a.n4, world);
  const int blocks = (int)std::min<int64_t>(std::max<int64_t>(ceil_div64(per, 256), 1), (int64_t)148 * 8);
  SMB_LAUNCH(dist_adam_kernel, blocks, 256, 0, st, a, s, seg);
  const int zblocks = (int)std::min<int64_t>(ceil_div64(a.n4, 256), (int64_t)148 * 8);
  SMB_LAUNCH(dist_adam_finish_kernel, zblocks, 256, 0, st, a, grad_ptrs[rank]);
  return SMB_OK;
}

int launch_sumsq_segments(const float* x, int64_t n, const int64_t* seg_begin, const float* seg_coef, int num_segments,
                          float clamp_lo, float clamp_hi, float* out, cudaStream_t st) {
  if (n == 0) return SMB_OK;
  SegTable seg;
  int rc = fill_segments(&seg, n, seg_begin, seg_coef, num_segments);
  if (rc) return rc;
  int blocks = (int)std::min<int64_t>(ceil_div64((n + 3) >> 2, 256 * 4), (int64_t)148 * 8);
  SMB_LAUNCH(sumsq_clamped_seg_kernel, blocks, 256, 0, st, x, n, clamp_lo, clamp_hi, seg, out);
  return SMB_OK;
}

int launch_sumsq_clamped(const float* x, int64_t n, float coef, float clamp_lo, float clamp_hi, float* out,
                         cudaStream_t st) {
  if (n == 0) return SMB_OK;
  int blocks = (int)std::min<int64_t>(ceil_div64(n, 256 * 8), (int64_t)148 * 8);
  SMB_LAUNCH(sumsq_clamped_kernel, blocks, 256, 0, st, x, n, coef, clamp_lo, clamp_hi, out);
  return SMB_OK;
}

}  // namespace smb
