// Host-side engine behind the C-ABI (include/stylemesh_b200.h): VGG weights in kernel layout, per-resolution
// working sets ("slots"), forward / loss-term / backward orchestration.  Every arithmetic step is one of the
// kernels in texture_kernels.cu, vgg_simt_kernels.cu or tc_kernels.cu; nothing here computes on the CPU except
// the one-time weight repacking.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/stylemesh_b200.h"
#include "smb_common.cuh"
#include "smb_kernels.h"
#include "tc_common.cuh"

namespace smb {

// ---- error string ------------------------------------------------------------------------------------------
static thread_local char g_err[2048] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

namespace tc {
TmapCache& tmap_cache() {
  static TmapCache c;
  return c;
}
}  // namespace tc

// programmatic dependent launch (smb_common.cuh) is on unless SMB_PDL=0
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SMB_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// ---- VGG-19 topology up to conv5_1 (model/losses/content_and_style_losses.py:11-32,47-66) ---------------------
static const int kCin[SMB_NUM_VGG_CONVS] = {3, 64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512};
static const int kCout[SMB_NUM_VGG_CONVS] = {64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512};
static const bool kPoolBefore[SMB_NUM_VGG_CONVS] = {false, false, true, false, true, false, false,
                                                    false, true,  false, false, false, true};

struct DeviceArena {
  std::vector<void*> ptrs;
  int64_t bytes = 0;
  template <typename T>
  int alloc(T** out, int64_t count) {
    void* p = nullptr;
    const size_t n = (size_t)std::max<int64_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
      return SMB_ERR_CUDA;
    }
    ptrs.push_back(p);
    bytes += (int64_t)n;
    *out = reinterpret_cast<T*>(p);
    return SMB_OK;
  }
  int alloc_act(Act* a, int H, int W, int C) {
    a->H = H;
    a->W = W;
    a->C = C;
    int rc = alloc(&a->hi, a->elems());
    if (rc) return rc;
    return alloc(&a->lo, a->elems());
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    bytes = 0;
  }
  ~DeviceArena() { release(); }
};

struct ConvLayer {
  PackedB fwd, dgrad;     // conv 0: fwd = [1][64][32] im2col-ordered operand of conv_first_tc, dgrad = N padded to 16
  float* bias = nullptr;
  float* w_oihw = nullptr;   // device copy of the original layout (first layer only)
};

struct Slot {
  int H = 0, W = 0;
  int h[SMB_NUM_VGG_CONVS], w[SMB_NUM_VGG_CONVS];
  Act y[SMB_NUM_VGG_CONVS];        // relu(conv_i)
  Act pooled[SMB_NUM_VGG_CONVS];   // input of conv_i when a pool precedes it
  Act dz[SMB_NUM_VGG_CONVS];       // gradient w.r.t. the pre-activation of conv_i (lazy)
  float* pend[SMB_NUM_VGG_CONVS];  // pending loss gradient w.r.t. relu(conv_i), fp32 [P][C] (lazy)
  bool has_pend[SMB_NUM_VGG_CONVS];
  float* gpool[SMB_NUM_VGG_CONVS]; // fp32 gradient w.r.t. pooled input of conv_i (lazy)
  Act fm;                          // masked copy scratch (sized for the largest layer, lazy)
  float* gram_partial = nullptr;
  int64_t gram_partial_elems = 0;
  __nv_bfloat16 *bmat_hi = nullptr, *bmat_lo = nullptr;
  // Style term whose Gram backward is deferred into the data-gradient conv of the layer above (igemm_ph fused 1x1
  // term): its seed matrix and a private copy of its pixel mask live until smb_level_backward.
  __nv_bfloat16* fb_hi[SMB_NUM_VGG_CONVS];
  __nv_bfloat16* fb_lo[SMB_NUM_VGG_CONVS];
  float* fmask[SMB_NUM_VGG_CONVS];
  bool fused[SMB_NUM_VGG_CONVS], fused_masked[SMB_NUM_VGG_CONVS];
  int last_done = -1;
  unsigned valid = 0;            // bit i: y[i] holds the features of the last forward (inference-only passes skip some)
  bool inference_only = false;   // last forward was smb_level_forward_features: no loss terms, no backward
  DeviceArena arena;
  Slot() {
    for (int i = 0; i < SMB_NUM_VGG_CONVS; ++i) {
      pend[i] = nullptr;
      has_pend[i] = false;
      gpool[i] = nullptr;
      fb_hi[i] = fb_lo[i] = nullptr;
      fmask[i] = nullptr;
      fused[i] = fused_masked[i] = false;
    }
  }
};

// ---- optional per-kernel-class timing (CUDA events on the launch stream; used by bench.py's roofline pass) -------
enum TimingClass : int {
  CLS_CONV_FIRST = 0, CLS_IGEMM_FWD, CLS_IGEMM_DGRAD, CLS_IGEMM_GRAMBWD, CLS_GRAM, CLS_GRAM_MSE, CLS_POOL,
  CLS_FIRST_DGRAD, CLS_CONTENT, CLS_MISC, CLS_COUNT
};

struct Timing {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int cls; size_t e0, e1; };
  std::vector<Rec> recs;
  double flops[CLS_COUNT] = {0};
  cudaEvent_t next() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  ~Timing() {
    for (auto e : pool) cudaEventDestroy(e);
  }
};

struct ScopedTimer {
  Timing* t;
  cudaStream_t st;
  size_t e1 = 0;
  ScopedTimer(Timing& tm, int cls, cudaStream_t s, double flops = 0.0) : t(tm.on ? &tm : nullptr), st(s) {
    if (!t) return;
    const size_t i0 = t->used;
    cudaEventRecord(t->next(), st);
    e1 = t->used;
    t->next();
    t->recs.push_back({cls, i0, e1});
    t->flops[cls] += flops;
  }
  ~ScopedTimer() {
    if (t) cudaEventRecord(t->pool[e1], st);
  }
};

}  // namespace smb

using namespace smb;

struct smb_ctx {
  Timing timing;
  int conv_impl = IMPL_TC_PH, gram_impl = IMPL_TC;
  bool vgg_loaded = false;
  ConvLayer conv[SMB_NUM_VGG_CONVS];
  DeviceArena weights;
  std::vector<std::unique_ptr<Slot>> slots;
};

namespace smb {

static int upload_packed(DeviceArena& ar, PackedB* out, const std::vector<float>& vals, int taps, int N, int K) {
  const int64_t n = (int64_t)taps * N * K;
  std::vector<uint16_t> hi(n), lo(n);
  for (int64_t i = 0; i < n; ++i) {
    const uint16_t h = host_f2bf(vals[i]);
    hi[i] = h;
    lo[i] = host_f2bf(vals[i] - host_bf2f(h));
  }
  out->taps = taps;
  out->N = N;
  out->K = K;
  int rc = ar.alloc(&out->hi, n);
  if (rc) return rc;
  rc = ar.alloc(&out->lo, n);
  if (rc) return rc;
  SMB_CUDA_CHECK(cudaMemcpy(out->hi, hi.data(), n * 2, cudaMemcpyHostToDevice));
  SMB_CUDA_CHECK(cudaMemcpy(out->lo, lo.data(), n * 2, cudaMemcpyHostToDevice));
  return SMB_OK;
}

// forward operand: B[tap=r*3+s][n=co][k=ci] = w[co][ci][r][s]
static void pack_fwd(const float* w, int Cout, int Cin, std::vector<float>& out) {
  out.resize((size_t)9 * Cout * Cin);
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s)
      for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
          out[((size_t)(r * 3 + s) * Cout + co) * Cin + ci] = w[(((size_t)co * Cin + ci) * 3 + r) * 3 + s];
}
// data-gradient operand: B[tap=r*3+s][n=ci][k=co] = w[co][ci][2-r][2-s]
static void pack_dgrad(const float* w, int Cout, int Cin, std::vector<float>& out) {
  out.resize((size_t)9 * Cout * Cin);
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s)
      for (int ci = 0; ci < Cin; ++ci)
        for (int co = 0; co < Cout; ++co)
          out[((size_t)(r * 3 + s) * Cin + ci) * Cout + co] = w[(((size_t)co * Cin + ci) * 3 + (2 - r)) * 3 + (2 - s)];
}

static int igemm(int impl, const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st,
                 const FusedTerm* ft = nullptr) {
  if (ft && !(impl == IMPL_TC_PH && b.taps == 9 && b.N % 64 == 0)) {
    set_error("fused 1x1 term needs the igemm_ph 3x3 kernel");
    return SMB_ERR_STATE;
  }
  if (impl == IMPL_TC_PH)
    return (b.taps == 9 && (b.N % 64 == 0 || b.N == 16)) ? launch_igemm_ph(a, b, ep, st, ft)
                                                         : launch_igemm_tc2(a, b, ep, st);
  if (impl == IMPL_TC) return launch_igemm_tc2(a, b, ep, st);
  return launch_igemm_simt(a, b, ep, st);
}
// masked Gram partials: the tcgen05 kernel masks pixels in shared memory, the SIMT kernel reads a masked copy
static int gram(int impl, const Act& f, const float* rowmask, const Act& scratch, float* partial, int nsplit,
                cudaStream_t st) {
  if (impl == IMPL_TC) return launch_gram_tc(f, rowmask, partial, nsplit, st);
  Act src = f;
  if (rowmask) {
    Act fm = scratch;
    fm.H = f.H;
    fm.W = f.W;
    fm.C = f.C;
    int rc = launch_mask_rows(f, rowmask, fm, st);
    if (rc) return rc;
    src = fm;
  }
  return launch_gram_simt(src, partial, nsplit, st);
}
static int igemm_timed(smb_ctx* ctx, int cls, const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st,
                       const FusedTerm* ft = nullptr) {
  const double fl = 2.0 * (double)a.pixels() * b.N * b.K * b.taps + (ft ? 2.0 * (double)a.pixels() * b.N * ft->g.K : 0.0);
  ScopedTimer tm(ctx->timing, cls, st, fl);
  return igemm(ctx->conv_impl, a, b, ep, st, ft);
}

static bool pool_side_enabled() {
  static const bool on = [] {
    const char* e = getenv("SMB_PH_POOL_SIDE");
    return !(e && atoi(e) == 0);
  }();
  return on;
}

// SMB_PH_FUSE=0 keeps every Gram backward as its own launch (fp32 pending gradient + addend read in the conv epilogue)
static bool fuse_enabled() {
  static const bool on = [] {
    const char* e = getenv("SMB_PH_FUSE");
    return !(e && e[0] == '0');
  }();
  return on;
}

static Slot* get_slot(smb_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) {
    set_error("invalid context or slot id %d", slot);
    return nullptr;
  }
  return ctx->slots[slot].get();
}

static int ensure_scratch(Slot& s, int gram_impl) {
  if (!s.fm.hi && gram_impl != IMPL_TC) {      // masked feature copy: only the SIMT Gram kernel needs one
    int64_t max_elems = 0;
    for (int i = 0; i < SMB_NUM_VGG_CONVS; ++i) max_elems = std::max(max_elems, s.y[i].elems());
    s.fm.H = 1;
    s.fm.W = 1;
    s.fm.C = 1;
    int rc = s.arena.alloc(&s.fm.hi, max_elems);
    if (rc) return rc;
    rc = s.arena.alloc(&s.fm.lo, max_elems);
    if (rc) return rc;
  }
  if (!s.bmat_hi) {
    int rc = s.arena.alloc(&s.bmat_hi, 512 * 512);
    if (rc) return rc;
    rc = s.arena.alloc(&s.bmat_lo, 512 * 512);
    if (rc) return rc;
  }
  return SMB_OK;
}

static int ensure_gram_partial(Slot& s, int64_t elems) {
  if (s.gram_partial_elems >= elems) return SMB_OK;
  s.gram_partial_elems = elems;
  return s.arena.alloc(&s.gram_partial, elems);   // the previous (smaller) block stays in the arena until destroy
}

static int ensure_pend(Slot& s, int conv) {
  if (s.pend[conv]) return SMB_OK;
  return s.arena.alloc(&s.pend[conv], s.y[conv].elems());
}

// pend[conv][p][c] (+)= m_p * sum_k F[p][k] * Bmat[c][k]   (m in {0,1}: the row scale equals multiplying masked features)
static int gram_backward_to_pend(smb_ctx* ctx, Slot& s, int conv, __nv_bfloat16* b_hi, __nv_bfloat16* b_lo,
                                 const float* rowmask, cudaStream_t st) {
  int rc = ensure_pend(s, conv);
  if (rc) return rc;
  const Act& feat = s.y[conv];
  PackedB b;
  b.hi = b_hi;
  b.lo = b_lo;
  b.taps = 1;
  b.N = feat.C;
  b.K = feat.C;
  Epilogue ep;
  ep.rowscale = rowmask;
  ep.out_f32 = s.pend[conv];
  if (s.has_pend[conv]) ep.addend = s.pend[conv];
  rc = igemm_timed(ctx, CLS_IGEMM_GRAMBWD, feat, b, ep, st);
  if (rc) return rc;
  s.has_pend[conv] = true;
  return SMB_OK;
}

// masked Gram partials of layer `conv`
static int gram_partials(smb_ctx* ctx, Slot& s, int conv, const float* rowmask, int* nsplit, cudaStream_t st) {
  int rc = ensure_scratch(s, ctx->gram_impl);
  if (rc) return rc;
  const Act& src = s.y[conv];
  const int ns = gram_num_splits(src.pixels(), src.C, ctx->gram_impl);
  rc = ensure_gram_partial(s, (int64_t)ns * src.C * src.C);
  if (rc) return rc;
  {
    ScopedTimer tm(ctx->timing, CLS_GRAM, st, 2.0 * (double)src.pixels() * src.C * src.C);
    rc = gram(ctx->gram_impl, src, rowmask, s.fm, s.gram_partial, ns, st);
  }
  if (rc) return rc;
  *nsplit = ns;
  return SMB_OK;
}

__global__ void gram_reduce_kernel(const float* __restrict__ partial, int nsplit, int64_t CC, float inv_n,
                                   float* __restrict__ out) {
  pdl_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < CC; e += stride) {
    float g = 0.f;
    for (int s = 0; s < nsplit; ++s) g += partial[(int64_t)s * CC + e];
    out[e] = g * inv_n;
  }
}

}  // namespace smb

// =============================================================================================================
// C-ABI
// =============================================================================================================
extern "C" {

int smb_abi_version(void) { return SMB_ABI_VERSION; }
const char* smb_last_error(void) { return smb::get_error(); }

// ---- texture ------------------------------------------------------------------------------------------------
static int fill_layers(TexLayerSet* t, float* const* layers, const int* lw, const int* lh, int L, int C) {
  SMB_REQUIRE(L >= 1 && L <= SMB_MAX_TEX_LAYERS, "num_layers=%d out of range [1,%d]", L, SMB_MAX_TEX_LAYERS);
  SMB_REQUIRE(C >= 1 && C <= SMB_MAX_TEX_CHANNELS, "channels=%d out of range [1,%d]", C, SMB_MAX_TEX_CHANNELS);
  t->L = L;
  t->C = C;
  for (int l = 0; l < L; ++l) {
    SMB_REQUIRE(layers[l] != nullptr && lw[l] >= 1 && lh[l] >= 1, "texture layer %d is null or empty", l);
    t->ptr[l] = layers[l];
    t->W[l] = lw[l];
    t->H[l] = lh[l];
  }
  return SMB_OK;
}

int smb_uv_sample_fwd(const float* const* layers, const int* layer_w, const int* layer_h, int num_layers,
                      int channels, const float* grid, int H, int W, float clamp_lo, float clamp_hi, float* out,
                      void* stream) {
  SMB_REQUIRE(layers && layer_w && layer_h && grid && out, "uv_sample_fwd: null argument");
  SMB_REQUIRE(H >= 0 && W >= 0, "uv_sample_fwd: negative size");
  TexLayerSet t;
  int rc = fill_layers(&t, const_cast<float* const*>(layers), layer_w, layer_h, num_layers, channels);
  if (rc) return rc;
  return launch_uv_sample_fwd(t, grid, H, W, clamp_lo, clamp_hi, out, (cudaStream_t)stream);
}

int smb_uv_texel_index(const float* grid, int num_pixels, int tex_w, int tex_h, int* xy0, float* w4, void* stream) {
  SMB_REQUIRE(grid && xy0 && w4 && tex_w >= 1 && tex_h >= 1 && num_pixels >= 0, "uv_texel_index: bad argument");
  return launch_uv_texel_index(grid, num_pixels, tex_w, tex_h, xy0, w4, (cudaStream_t)stream);
}

int smb_uv_scatter_bwd(float* const* grad_layers, const int* layer_w, const int* layer_h, int num_layers,
                       int channels, const float* grid, int H, int W, const float* grad_out, const float* hook0,
                       const float* hook1, void* stream) {
  SMB_REQUIRE(grad_layers && layer_w && layer_h && grid && grad_out, "uv_scatter_bwd: null argument");
  TexLayerSet t;
  int rc = fill_layers(&t, grad_layers, layer_w, layer_h, num_layers, channels);
  if (rc) return rc;
  return launch_uv_scatter_bwd(t, grid, H, W, grad_out, hook0, hook1, (cudaStream_t)stream);
}

int smb_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float clamp_lo, float clamp_hi, float reg_coef,
                  float grad_scale, void* stream) {
  SMB_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0, "adam_step: null argument");
  return launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, clamp_lo, clamp_hi, reg_coef,
                     grad_scale, (cudaStream_t)stream);
}

int smb_texreg_value(const float* param, int64_t n, float coef, float clamp_lo, float clamp_hi, float* out_accum,
                     void* stream) {
  SMB_REQUIRE(param && out_accum && n >= 0, "texreg_value: null argument");
  return launch_sumsq_clamped(param, n, coef, clamp_lo, clamp_hi, out_accum, (cudaStream_t)stream);
}

int smb_adam_step_segments(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                           const int64_t* seg_begin, const float* seg_reg_coef, int num_segments, float lr, float beta1,
                           float beta2, float eps, int step, float clamp_lo, float clamp_hi, float grad_scale,
                           void* stream) {
  SMB_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0, "adam_step_segments: null argument");
  return launch_adam_segments(param, grad, exp_avg, exp_avg_sq, n, seg_begin, seg_reg_coef, num_segments, lr, beta1,
                              beta2, eps, step, clamp_lo, clamp_hi, grad_scale, (cudaStream_t)stream);
}

int smb_texreg_value_segments(const float* param, int64_t n, const int64_t* seg_begin, const float* seg_coef,
                              int num_segments, float clamp_lo, float clamp_hi, float* out_accum, void* stream) {
  SMB_REQUIRE(param && out_accum && n >= 0, "texreg_value_segments: null argument");
  return launch_sumsq_segments(param, n, seg_begin, seg_coef, num_segments, clamp_lo, clamp_hi, out_accum,
                               (cudaStream_t)stream);
}

int smb_dist_adam_step(int rank, int world, float* const* grad_ptrs, float* const* param_ptrs,
                       unsigned int* const* flag_ptrs, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const int64_t* seg_begin, const float* seg_reg_coef, int num_segments, float lr, float beta1,
                       float beta2, float eps, int step, float clamp_lo, float clamp_hi, unsigned int epoch,
                       void* stream) {
  SMB_REQUIRE(grad_ptrs && param_ptrs && flag_ptrs && exp_avg && exp_avg_sq, "dist_adam_step: null argument");
  return launch_dist_adam(rank, world, grad_ptrs, param_ptrs, flag_ptrs, exp_avg, exp_avg_sq, n, seg_begin, seg_reg_coef,
                          num_segments, lr, beta1, beta2, eps, step, clamp_lo, clamp_hi, epoch, (cudaStream_t)stream);
}

// ---- view preparation ---------------------------------------------------------------------------------------------
int smb_view_uv_to_grid(const float* uv_hw3, int H, int W, float* grid_hw2, unsigned char* valid,
                        const double* depth_at_uv, void* stream) {
  SMB_REQUIRE(uv_hw3 && grid_hw2 && H >= 0 && W >= 0, "view_uv_to_grid: bad argument");
  SMB_REQUIRE(valid || !depth_at_uv, "view_uv_to_grid: a depth map without a mask output");
  return launch_view_uv_grid(uv_hw3, H, W, grid_hw2, valid, depth_at_uv, (cudaStream_t)stream);
}

int smb_view_gather2d(const void* src, int elem_bytes, int Hs, int Ws, const int* ytab, const int* xtab, int Hd, int Wd,
                      void* dst, void* stream) {
  SMB_REQUIRE(src && ytab && xtab && dst && Hs > 0 && Ws > 0 && Hd >= 0 && Wd >= 0, "view_gather2d: bad argument");
  return launch_view_gather2d(src, elem_bytes, Ws, ytab, xtab, Hd, Wd, dst, (cudaStream_t)stream);
}

int smb_view_resize_linear(const void* src, int src_type, double divisor, int Hs, int Ws, const int* yofs,
                           const double* yalpha, const int* xofs, const double* xalpha, int Hd, int Wd, double* dst,
                           void* stream) {
  SMB_REQUIRE(src && dst && Hs > 0 && Ws > 0 && Hd >= 0 && Wd >= 0, "view_resize_linear: bad argument");
  SMB_REQUIRE((Hs == Hd && Ws == Wd) || (yofs && yalpha && xofs && xalpha), "view_resize_linear: missing tables");
  SMB_REQUIRE(src_type != 2 || divisor > 0.0, "view_resize_linear: uint16 input needs a positive divisor");
  return launch_view_resize_linear(src, src_type, divisor, Hs, Ws, yofs, yalpha, xofs, xalpha, Hd, Wd, dst,
                                   (cudaStream_t)stream);
}

int smb_view_depth_levels(const double* depth, int64_t n, const double* levels, int num_levels, double min_depth,
                          int depth_is_f32, float* depth_level, float* depth_f32, int64_t* rounded, int64_t* other,
                          float* weight, void* stream) {
  SMB_REQUIRE(depth && levels && depth_level && rounded && other && weight && n >= 0, "view_depth_levels: null argument");
  return launch_view_depth_levels(depth, n, levels, num_levels, min_depth, depth_is_f32, depth_level, depth_f32,
                                  reinterpret_cast<long long*>(rounded), reinterpret_cast<long long*>(other), weight,
                                  (cudaStream_t)stream);
}

int smb_view_rgb_pre(const unsigned char* rgb_hwc, int H, int W, float* out_chw, void* stream) {
  SMB_REQUIRE(rgb_hwc && out_chw && H >= 0 && W >= 0, "view_rgb_pre: bad argument");
  return launch_view_rgb_pre(rgb_hwc, H, W, out_chw, (cudaStream_t)stream);
}

int smb_view_angle_degrees(const float* cos_angle, int64_t n, float* degrees, void* stream) {
  SMB_REQUIRE(cos_angle && degrees && n >= 0, "view_angle_degrees: bad argument");
  return launch_view_angle_degrees(cos_angle, n, degrees, (cudaStream_t)stream);
}

int smb_view_erode3x3(const float* x, int H, int W, float* out, void* stream) {
  SMB_REQUIRE(x && out && x != out && H >= 0 && W >= 0, "view_erode3x3: bad argument (in-place is not supported)");
  return launch_view_erode3x3(x, H, W, out, (cudaStream_t)stream);
}

int smb_view_level_masks(const unsigned char* mask, const int64_t* rounded, const int64_t* other, const float* interp_w,
                         int H, int W, int num_levels, float* level_mask, float* level_weight, void* stream) {
  SMB_REQUIRE(mask && rounded && other && interp_w && level_mask && level_weight && H >= 0 && W >= 0 && num_levels >= 0,
              "view_level_masks: bad argument");
  return launch_view_level_masks(mask, reinterpret_cast<const long long*>(rounded),
                                 reinterpret_cast<const long long*>(other), interp_w, H, W, num_levels, level_mask,
                                 level_weight, (cudaStream_t)stream);
}

int smb_view_level_plan(const float* src_mask, const float* src_weight, const float* angle_guidance,
                        const float* angle_degrees, float threshold, int Hr, int Wr, int H, int W, float* hook0,
                        float* hook1, int num_layers, const int* lh, const int* lw, float* layer_masks, int split,
                        unsigned int* counts, void* stream) {
  SMB_REQUIRE(Hr > 0 && Wr > 0 && H > 0 && W > 0, "view_level_plan: empty map");
  SMB_REQUIRE(num_layers == 0 || (lh && lw), "view_level_plan: null layer sizes");
  return launch_view_level_plan(src_mask, src_weight, angle_guidance, angle_degrees, threshold, Hr, Wr, H, W, hook0,
                                hook1, num_layers, lh, lw, layer_masks, split, counts, (cudaStream_t)stream);
}

int smb_texture_post_rgb8(const float* bgr_chw, int H, int W, unsigned char* rgb_hwc, void* stream) {
  SMB_REQUIRE(bgr_chw && rgb_hwc && H >= 0 && W >= 0, "texture_post_rgb8: bad argument");
  return launch_texture_post_rgb8(bgr_chw, H, W, rgb_hwc, (cudaStream_t)stream);
}

int smb_mip_downsample2x(const float* src_chw, int H, int W, float* dst_chw, void* stream) {
  SMB_REQUIRE(src_chw && dst_chw && H > 0 && W > 0, "mip_downsample2x: bad argument");
  return launch_mip_downsample2x(src_chw, H, W, dst_chw, (cudaStream_t)stream);
}

int smb_mip_preview(const float* const* mips, const int* mip_w, const int* mip_h, int num_mips, const float* uv,
                    int uv_channels, int H, int W, float lod_bias, unsigned char* rgb_hwc, void* stream) {
  SMB_REQUIRE(mips && mip_w && mip_h && uv && rgb_hwc && H >= 0 && W >= 0, "mip_preview: bad argument");
  return launch_mip_preview(mips, mip_w, mip_h, num_mips, uv, uv_channels, H, W, lod_bias, rgb_hwc,
                            (cudaStream_t)stream);
}

int smb_raster_view(const float* verts, int num_verts, const int* faces, int num_faces, const float* corner_uv,
                    const float* corner_normal, const float* view3x4, const float* proj6, int w, int h, float near_plane,
                    float far_plane, float tex_size, int flip, float* eye_scratch, unsigned long long* zbuf,
                    float* uv_out, float* angle_out, float* depth_out, void* stream) {
  return launch_raster_view(verts, num_verts, faces, num_faces, corner_uv, corner_normal, view3x4, proj6, w, h,
                            near_plane, far_plane, tex_size, flip, eye_scratch, zbuf, uv_out, angle_out, depth_out,
                            (cudaStream_t)stream);
}

// ---- context --------------------------------------------------------------------------------------------------
smb_ctx* smb_ctx_create(void) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("smb_ctx_create: no CUDA device is current (%s)", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
    set_error("smb_ctx_create: device %d is not an sm_100-class GPU (compute capability %d.%d); this library "
              "contains sm_100a code only", dev, prop.major, prop.minor);
    return nullptr;
  }
  return new smb_ctx();
}

void smb_ctx_destroy(smb_ctx* ctx) { delete ctx; }

int smb_ctx_set_impl(smb_ctx* ctx, int conv_impl, int gram_impl) {
  SMB_REQUIRE(ctx, "null context");
  SMB_REQUIRE((conv_impl == IMPL_SIMT || conv_impl == IMPL_TC || conv_impl == IMPL_TC_PH) &&
                  (gram_impl == IMPL_SIMT || gram_impl == IMPL_TC),
              "conv_impl must be SMB_IMPL_SIMT / SMB_IMPL_TC / SMB_IMPL_TC_PH, gram_impl SMB_IMPL_SIMT / SMB_IMPL_TC");
  ctx->conv_impl = conv_impl;
  ctx->gram_impl = gram_impl;
  return SMB_OK;
}

int smb_ctx_load_vgg(smb_ctx* ctx, const float* const* weights_oihw, const float* const* bias, int num_convs) {
  SMB_REQUIRE(ctx && weights_oihw && bias, "load_vgg: null argument");
  SMB_REQUIRE(num_convs == SMB_NUM_VGG_CONVS, "load_vgg: expected %d conv layers (conv1_1..conv5_1), got %d",
              SMB_NUM_VGG_CONVS, num_convs);
  ctx->weights.release();
  ctx->vgg_loaded = false;
  std::vector<float> tmp;
  for (int i = 0; i < SMB_NUM_VGG_CONVS; ++i) {
    SMB_REQUIRE(weights_oihw[i] && bias[i], "load_vgg: conv %d has a null weight or bias", i);
    ConvLayer& c = ctx->conv[i];
    int rc = ctx->weights.alloc(&c.bias, kCout[i]);
    if (rc) return rc;
    SMB_CUDA_CHECK(cudaMemcpy(c.bias, bias[i], kCout[i] * sizeof(float), cudaMemcpyHostToDevice));
    if (i == 0) {
      const int64_t n = (int64_t)kCout[0] * kCin[0] * 9;
      rc = ctx->weights.alloc(&c.w_oihw, n);
      if (rc) return rc;
      SMB_CUDA_CHECK(cudaMemcpy(c.w_oihw, weights_oihw[0], n * sizeof(float), cudaMemcpyHostToDevice));
      // forward operand of the first layer for conv_first_tc: [n=co][k = ci*9 + r*3 + s], K padded 27 -> 32
      tmp.assign((size_t)kCout[0] * 32, 0.f);
      for (int co = 0; co < kCout[0]; ++co)
        for (int k = 0; k < 27; ++k) tmp[(size_t)co * 32 + k] = weights_oihw[0][(size_t)co * 27 + k];
      rc = upload_packed(ctx->weights, &c.fwd, tmp, 1, kCout[0], 32);
      if (rc) return rc;
      // data-gradient operand of the first layer for the tensor-core path: N = 3 input channels padded to 16
      // B[tap=r*3+s][n=ci][k=co] = w[co][ci][2-r][2-s]
      tmp.assign((size_t)9 * 16 * kCout[0], 0.f);
      for (int r = 0; r < 3; ++r)
        for (int s2 = 0; s2 < 3; ++s2)
          for (int ci = 0; ci < 3; ++ci)
            for (int co = 0; co < kCout[0]; ++co)
              tmp[((size_t)(r * 3 + s2) * 16 + ci) * kCout[0] + co] =
                  weights_oihw[0][(((size_t)co * 3 + ci) * 3 + (2 - r)) * 3 + (2 - s2)];
      rc = upload_packed(ctx->weights, &c.dgrad, tmp, 9, 16, kCout[0]);
      if (rc) return rc;
    } else {
      pack_fwd(weights_oihw[i], kCout[i], kCin[i], tmp);
      rc = upload_packed(ctx->weights, &c.fwd, tmp, 9, kCout[i], kCin[i]);
      if (rc) return rc;
      pack_dgrad(weights_oihw[i], kCout[i], kCin[i], tmp);
      rc = upload_packed(ctx->weights, &c.dgrad, tmp, 9, kCin[i], kCout[i]);
      if (rc) return rc;
    }
  }
  ctx->vgg_loaded = true;
  return SMB_OK;
}

int smb_level_begin(smb_ctx* ctx, int H, int W) {
  if (!ctx) {
    set_error("null context");
    return SMB_ERR_ARG;
  }
  SMB_REQUIRE(H >= 16 && W >= 16, "level_begin: input %dx%d is too small for four 2x2 pools", H, W);
  for (size_t i = 0; i < ctx->slots.size(); ++i)
    if (ctx->slots[i]->H == H && ctx->slots[i]->W == W) return (int)i;
  std::unique_ptr<Slot> s(new Slot());
  s->H = H;
  s->W = W;
  int h = H, w = W;
  for (int i = 0; i < SMB_NUM_VGG_CONVS; ++i) {
    if (kPoolBefore[i]) {
      h /= 2;
      w /= 2;
      int rc = s->arena.alloc_act(&s->pooled[i], h, w, kCin[i]);
      if (rc) return rc;
    }
    s->h[i] = h;
    s->w[i] = w;
    int rc = s->arena.alloc_act(&s->y[i], h, w, kCout[i]);
    if (rc) return rc;
  }
  ctx->slots.push_back(std::move(s));
  return (int)ctx->slots.size() - 1;
}

// keep_mask < 0: training forward (every layer materialised).  Otherwise an inference-only pass: a layer that only
// feeds a max-pool and is not in keep_mask leaves igemm_ph through the pooled epilogue (a quarter of the bytes, no
// pool launch) and its full-resolution features are never written.
static int level_forward_impl(smb_ctx* ctx, int slot, const float* image, int last_conv, long long keep_mask,
                              void* stream) {
  Slot* sp = get_slot(ctx, slot);
  if (!sp) return SMB_ERR_ARG;
  Slot& s = *sp;
  SMB_REQUIRE(ctx->vgg_loaded, "level_forward: VGG weights not loaded (smb_ctx_load_vgg)");
  SMB_REQUIRE(image && last_conv >= 0 && last_conv < SMB_NUM_VGG_CONVS, "level_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const bool inference = keep_mask >= 0;
  s.valid = 0;
  {
    Epilogue ep;
    ep.relu = 1;
    ep.out_hi = s.y[0].hi;
    ep.out_lo = s.y[0].lo;
    ScopedTimer tm(ctx->timing, CLS_CONV_FIRST, st, 2.0 * 27 * 64 * (double)s.H * s.W);
    int rc;
    if (ctx->conv_impl != IMPL_SIMT)
      rc = launch_conv_first_tc(image, s.H, s.W, ctx->conv[0].fwd.hi, ctx->conv[0].fwd.lo, ctx->conv[0].bias, ep, st);
    else
      rc = launch_conv_first_fwd(image, s.H, s.W, ctx->conv[0].w_oihw, ctx->conv[0].bias, kCout[0], ep, st);
    if (rc) return rc;
    s.valid |= 1u;
  }
  bool pooled_ready = false;     // pooled[i] was written by the epilogue of conv i-1
  for (int i = 1; i <= last_conv; ++i) {
    Act x = s.y[i - 1];
    if (kPoolBefore[i]) {
      if (!pooled_ready) {
        ScopedTimer tm(ctx->timing, CLS_POOL, st);
        int rc = launch_maxpool_fwd(s.y[i - 1], s.pooled[i], st);
        if (rc) return rc;
      }
      x = s.pooled[i];
    }
    const bool feeds_pool = ctx->conv_impl == IMPL_TC_PH && i < last_conv && kPoolBefore[i + 1] && kCout[i] % 64 == 0;
    const bool pool_out = inference && feeds_pool && !((keep_mask >> i) & 1);
    // a materialised layer that feeds a pool writes the pooled planes as a side output of its epilogue (no pool launch,
    // no second read of the full-resolution features); SMB_PH_POOL_SIDE=0 keeps the separate pool kernel
    const bool pool_side = feeds_pool && !pool_out && pool_side_enabled();
    Epilogue ep;
    ep.bias = ctx->conv[i].bias;
    ep.relu = 1;
    ep.out_hi = pool_out ? s.pooled[i + 1].hi : s.y[i].hi;
    ep.out_lo = pool_out ? s.pooled[i + 1].lo : s.y[i].lo;
    ep.pool2x2 = pool_out ? 1 : 0;
    if (pool_side) {
      ep.pool_hi = s.pooled[i + 1].hi;
      ep.pool_lo = s.pooled[i + 1].lo;
    }
    int rc = igemm_timed(ctx, CLS_IGEMM_FWD, x, ctx->conv[i].fwd, ep, st);
    if (rc) return rc;
    pooled_ready = pool_out || pool_side;
    if (!pool_out) s.valid |= 1u << i;
  }
  s.last_done = last_conv;
  s.inference_only = inference;
  for (int i = 0; i < SMB_NUM_VGG_CONVS; ++i) s.has_pend[i] = s.fused[i] = false;
  return SMB_OK;
}

int smb_level_forward(smb_ctx* ctx, int slot, const float* image, int last_conv, void* stream) {
  return level_forward_impl(ctx, slot, image, last_conv, -1, stream);
}

int smb_level_forward_features(smb_ctx* ctx, int slot, const float* image, int last_conv, unsigned int keep_mask,
                               void* stream) {
  return level_forward_impl(ctx, slot, image, last_conv, (long long)keep_mask, stream);
}

int smb_level_feature_shape(smb_ctx* ctx, int slot, int conv, int* C, int* h, int* w) {
  Slot* sp = get_slot(ctx, slot);
  if (!sp) return SMB_ERR_ARG;
  SMB_REQUIRE(conv >= 0 && conv < SMB_NUM_VGG_CONVS && C && h && w, "feature_shape: bad argument");
  *C = kCout[conv];
  *h = sp->h[conv];
  *w = sp->w[conv];
  return SMB_OK;
}

int smb_level_get_feature(smb_ctx* ctx, int slot, int conv, float* out_nchw, void* stream) {
  Slot* sp = get_slot(ctx, slot);
  if (!sp) return SMB_ERR_ARG;
  SMB_REQUIRE(conv >= 0 && conv <= sp->last_done && out_nchw, "get_feature: conv %d not computed (last=%d)", conv,
              sp->last_done);
  SMB_REQUIRE((sp->valid >> conv) & 1u, "get_feature: conv %d was not kept by smb_level_forward_features", conv);
  return launch_act_to_nchw(sp->y[conv], out_nchw, (cudaStream_t)stream);
}

namespace smb {
__global__ void act_to_f32_kernel(Act src, float* __restrict__ dst) {
  pdl_sync();
  const int64_t n2 = src.elems() >> 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const uint32_t h = reinterpret_cast<const uint32_t*>(src.hi)[i];
    const uint32_t l = reinterpret_cast<const uint32_t*>(src.lo)[i];
    reinterpret_cast<float2*>(dst)[i] = make_float2(bf16lo_to_f(h) + bf16lo_to_f(l), bf16hi_to_f(h) + bf16hi_to_f(l));
  }
}
}  // namespace smb

int smb_level_get_feature_nhwc(smb_ctx* ctx, int slot, int conv, float* out_nhwc, void* stream) {
  Slot* sp = get_slot(ctx, slot);
  if (!sp) return SMB_ERR_ARG;
  SMB_REQUIRE(conv >= 0 && conv <= sp->last_done && out_nhwc, "get_feature_nhwc: conv %d not computed (last=%d)",
              conv, sp->last_done);
  SMB_REQUIRE((sp->valid >> conv) & 1u, "get_feature_nhwc: conv %d was not kept by smb_level_forward_features", conv);
  const Act& a = sp->y[conv];
  if (a.elems() == 0) return SMB_OK;
  SMB_LAUNCH(smb::act_to_f32_kernel, (unsigned)std::min<int64_t>(ceil_div64(a.elems() >> 1,
             256), 148 * 16), 256, 0, (cudaStream_t)stream, a, out_nhwc);
  return SMB_OK;
}

int smb_ctx_release_slots(smb_ctx* ctx) {
  SMB_REQUIRE(ctx, "null context");
  SMB_CUDA_CHECK(cudaDeviceSynchronize());
  ctx->slots.clear();
  return SMB_OK;
}

int smb_level_gram(smb_ctx* ctx, int slot, int conv, const float* rowmask, float inv_n, float* gram_out,
                   void* stream) {
  Slot* sp = get_slot(ctx, slot);
  if (!sp) return SMB_ERR_ARG;
  SMB_REQUIRE(conv >= 0 && conv <= sp->last_done && gram_out, "level_gram: conv %d not computed (last=%d)", conv,
              sp->last_done);
  SMB_REQUIRE((sp->valid >> conv) & 1u, "level_gram: conv %d was not kept by smb_level_forward_features", conv);
  cudaStream_t st = (cudaStream_t)stream;
  int ns = 0;
  int rc = gram_partials(ctx, *sp, conv, rowmask, &ns, st);
  if (rc) return rc;
  const int64_t CC = (int64_t)sp->y[conv].C * sp->y[conv].C;
  SMB_LAUNCH(gram_reduce_kernel, (unsigned)std::min<int64_t>(ceil_div64(CC, 256), 592), 256, 0, st,
             sp->gram_partial, ns, CC, inv_n, gram_out);
  return SMB_OK;
}

int smb_level_style_term(smb_ctx* ctx, int slot, int conv, const float* rowmask, float inv_n, const float* target0,
                         float coef0, const float* target1, float coef1, const float* prev_sum, float avg_len,
                         float* gram_out, float* loss_accum, void* stream) {
  Slot* sp = get_slot(ctx, slot);
  if (!sp) return SMB_ERR_ARG;
  Slot& s = *sp;
  SMB_REQUIRE(conv >= 0 && conv <= s.last_done, "style_term: conv %d not computed (last=%d)", conv, s.last_done);
  if (s.inference_only) {
    set_error("style_term: the slot holds an inference-only forward (smb_level_forward_features)");
    return SMB_ERR_STATE;
  }
  SMB_REQUIRE(target0 && loss_accum, "style_term: null target or loss accumulator");
  cudaStream_t st = (cudaStream_t)stream;
  const Act& feat = s.y[conv];
  int ns = 0;
  int rc = gram_partials(ctx, s, conv, rowmask, &ns, st);
  if (rc) return rc;
  // Gram backward dF = m * (F . Bmat): folded into the data-gradient conv of the layer above when that conv runs on
  // igemm_ph at the same resolution (conv1_2, conv2_2, conv3_2, conv4_2 for r11..r41); resolved in smb_level_backward
  const bool defer = fuse_enabled() && ctx->conv_impl == IMPL_TC_PH && inv_n != 0.f && !s.fused[conv] &&
                     conv + 1 <= s.last_done && !kPoolBefore[conv + 1];
  __nv_bfloat16 *b_hi = s.bmat_hi, *b_lo = s.bmat_lo;
  if (defer) {
    if (!s.fb_hi[conv]) {
      rc = s.arena.alloc(&s.fb_hi[conv], (int64_t)feat.C * feat.C);
      if (rc) return rc;
      rc = s.arena.alloc(&s.fb_lo[conv], (int64_t)feat.C * feat.C);
      if (rc) return rc;
      rc = s.arena.alloc(&s.fmask[conv], feat.pixels());
      if (rc) return rc;
    }
    b_hi = s.fb_hi[conv];
    b_lo = s.fb_lo[conv];
  }
  {
    ScopedTimer tm(ctx->timing, CLS_GRAM_MSE, st);
    rc = launch_gram_mse(s.gram_partial, ns, feat.C, inv_n, target0, coef0, target1, coef1, prev_sum,
                         prev_sum ? avg_len : 1.f, gram_out, b_hi, b_lo, loss_accum, st);
  }
  if (rc) return rc;
  if (inv_n == 0.f) return SMB_OK;   // empty mask: constant loss, zero gradient (cs:140-141)
  if (defer) {
    if (rowmask)      // the caller's mask tensor need not outlive this call
      SMB_CUDA_CHECK(cudaMemcpyAsync(s.fmask[conv], rowmask, feat.pixels() * sizeof(float), cudaMemcpyDeviceToDevice, st));
    s.fused[conv] = true;
    s.fused_masked[conv] = rowmask != nullptr;
    return SMB_OK;
  }
  return gram_backward_to_pend(ctx, s, conv, s.bmat_hi, s.bmat_lo, rowmask, st);
}

int smb_level_content_term(smb_ctx* ctx, int slot, int conv, const float* target_nhwc, const float* rowmask,
                           float coef_loss, float coef_grad, float* loss_accum, void* stream) {
  Slot* sp = get_slot(ctx, slot);
  if (!sp) return SMB_ERR_ARG;
  Slot& s = *sp;
  SMB_REQUIRE(conv >= 0 && conv <= s.last_done, "content_term: conv %d not computed (last=%d)", conv, s.last_done);
  if (s.inference_only) {
    set_error("content_term: the slot holds an inference-only forward (smb_level_forward_features)");
    return SMB_ERR_STATE;
  }
  SMB_REQUIRE(target_nhwc && rowmask && loss_accum, "content_term: null argument");
  int rc = ensure_pend(s, conv);
  if (rc) return rc;
  if (!s.has_pend[conv])
    SMB_CUDA_CHECK(cudaMemsetAsync(s.pend[conv], 0, s.y[conv].elems() * sizeof(float), (cudaStream_t)stream));
  {
    ScopedTimer tm(ctx->timing, CLS_CONTENT, (cudaStream_t)stream);
    rc = launch_content_mse(s.y[conv], target_nhwc, rowmask, coef_loss, coef_grad, s.pend[conv], loss_accum,
                            (cudaStream_t)stream);
  }
  if (rc) return rc;
  s.has_pend[conv] = true;
  return SMB_OK;
}

int smb_level_backward(smb_ctx* ctx, int slot, float* d_image, void* stream) {
  Slot* sp = get_slot(ctx, slot);
  if (!sp) return SMB_ERR_ARG;
  Slot& s = *sp;
  SMB_REQUIRE(d_image, "level_backward: null output");
  cudaStream_t st = (cudaStream_t)stream;
  int top = -1;
  for (int i = s.last_done; i >= 0; --i)
    if (s.has_pend[i] || s.fused[i]) {
      top = i;
      break;
    }
  // deferred Gram backwards that no data-gradient conv can absorb (the top of the chain, or another conv kernel
  // selected since) become ordinary pending gradients now
  for (int i = 0; i <= top; ++i)
    if (s.fused[i] && (i == top || ctx->conv_impl != IMPL_TC_PH)) {
      int rc = gram_backward_to_pend(ctx, s, i, s.fb_hi[i], s.fb_lo[i], s.fused_masked[i] ? s.fmask[i] : nullptr, st);
      if (rc) return rc;
      s.fused[i] = false;
    }
  if (top < 0) {   // no loss term touched this level: zero gradient
    SMB_CUDA_CHECK(cudaMemsetAsync(d_image, 0, (size_t)3 * s.H * s.W * sizeof(float), st));
    return SMB_OK;
  }
  for (int i = 0; i <= top; ++i)
    if (!s.dz[i].hi) {
      int rc = s.arena.alloc_act(&s.dz[i], s.h[i], s.w[i], kCout[i]);
      if (rc) return rc;
    }
  // top of the chain: dz = pend ⊙ (y > 0)
  int rc;
  {
    ScopedTimer tm(ctx->timing, CLS_MISC, st);
    rc = launch_relu_mask_split(s.pend[top], s.y[top], s.dz[top], st);
  }
  if (rc) return rc;
  for (int i = top; i >= 1; --i) {
    const int j = i - 1;   // layer receiving the gradient
    if (kPoolBefore[i]) {
      if (!s.gpool[i]) {
        rc = s.arena.alloc(&s.gpool[i], s.pooled[i].elems());
        if (rc) return rc;
      }
      Epilogue ep;
      ep.out_f32 = s.gpool[i];
      rc = igemm_timed(ctx, CLS_IGEMM_DGRAD, s.dz[i], ctx->conv[i].dgrad, ep, st);
      if (rc) return rc;
      ScopedTimer tm(ctx->timing, CLS_POOL, st);
      rc = launch_maxpool_bwd_relu(s.gpool[i], s.has_pend[j] ? s.pend[j] : nullptr, s.y[j], s.dz[j], st);
      if (rc) return rc;
    } else {
      Epilogue ep;
      if (s.has_pend[j]) ep.addend = s.pend[j];
      ep.sign_hi = s.y[j].hi;
      ep.out_hi = s.dz[j].hi;
      ep.out_lo = s.dz[j].lo;
      FusedTerm ft;
      if (s.fused[j]) {
        ft.f = s.y[j];
        ft.g.hi = s.fb_hi[j];
        ft.g.lo = s.fb_lo[j];
        ft.g.taps = 1;
        ft.g.N = ft.g.K = s.y[j].C;
        ft.rowmask = s.fused_masked[j] ? s.fmask[j] : nullptr;
      }
      rc = igemm_timed(ctx, CLS_IGEMM_DGRAD, s.dz[i], ctx->conv[i].dgrad, ep, st, s.fused[j] ? &ft : nullptr);
      if (rc) return rc;
    }
  }
  {
    ScopedTimer tm(ctx->timing, CLS_FIRST_DGRAD, st, 2.0 * 27 * 64 * (double)s.H * s.W);
    if (ctx->conv_impl != IMPL_SIMT) {
      Epilogue ep;                          // tcgen05 implicit GEMM with N padded 3 -> 16, planar fp32 output
      ep.out_planar3 = d_image;
      rc = igemm(ctx->conv_impl, s.dz[0], ctx->conv[0].dgrad, ep, st);
    } else {
      rc = launch_conv_first_dgrad(s.dz[0], ctx->conv[0].w_oihw, kCout[0], d_image, st);
    }
  }
  if (rc) return rc;
  for (int i = 0; i < SMB_NUM_VGG_CONVS; ++i) s.has_pend[i] = s.fused[i] = false;
  return SMB_OK;
}

int64_t smb_launch_count(void) { return (int64_t)smb::launch_count(); }

int smb_ctx_set_timing(smb_ctx* ctx, int enabled) {
  SMB_REQUIRE(ctx, "null context");
  SMB_CUDA_CHECK(cudaDeviceSynchronize());
  ctx->timing.on = enabled != 0;
  ctx->timing.used = 0;
  ctx->timing.recs.clear();
  for (int i = 0; i < CLS_COUNT; ++i) ctx->timing.flops[i] = 0.0;
  return SMB_OK;
}

int smb_ctx_read_timing(smb_ctx* ctx, float* ms, double* flops, int* launches, int n) {
  SMB_REQUIRE(ctx && ms && flops && launches && n >= CLS_COUNT, "read_timing: need arrays of at least %d entries",
              (int)CLS_COUNT);
  SMB_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < n; ++i) {
    ms[i] = 0.f;
    flops[i] = 0.0;
    launches[i] = 0;
  }
  for (const auto& r : ctx->timing.recs) {
    float t = 0.f;
    SMB_CUDA_CHECK(cudaEventElapsedTime(&t, ctx->timing.pool[r.e0], ctx->timing.pool[r.e1]));
    ms[r.cls] += t;
    launches[r.cls] += 1;
  }
  for (int i = 0; i < CLS_COUNT; ++i) flops[i] = ctx->timing.flops[i];
  ctx->timing.used = 0;
  ctx->timing.recs.clear();
  for (int i = 0; i < CLS_COUNT; ++i) ctx->timing.flops[i] = 0.0;
  return CLS_COUNT;
}

int64_t smb_ctx_device_bytes(smb_ctx* ctx) {
  if (!ctx) return 0;
  int64_t b = ctx->weights.bytes;
  for (auto& s : ctx->slots) b += s->arena.bytes;
  return b;
}

int smb_debug_set_igemm_trace(void* buf) {
  smb::set_igemm_trace(reinterpret_cast<unsigned long long*>(buf));
  return SMB_OK;
}

// ---- unit-level entry points ------------------------------------------------------------------------------------
int smb_unit_conv3x3(int impl, const float* x, int Cin, int H, int W, const float* w_host, const float* b_host,
                     int Cout, int relu, int transpose_flip, float* y, void* stream) {
  SMB_REQUIRE(x && w_host && y, "unit_conv3x3: null argument");
  SMB_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "unit_conv3x3: Cin=%d and Cout=%d must be multiples of 64", Cin, Cout);
  cudaStream_t st = (cudaStream_t)stream;
  DeviceArena ar;
  const int K = transpose_flip ? Cout : Cin, N = transpose_flip ? Cin : Cout;
  Act a, o;
  int rc = ar.alloc_act(&a, H, W, K);
  if (rc) return rc;
  rc = ar.alloc_act(&o, H, W, N);
  if (rc) return rc;
  std::vector<float> tmp;
  PackedB b;
  if (transpose_flip) pack_dgrad(w_host, Cout, Cin, tmp); else pack_fwd(w_host, Cout, Cin, tmp);
  rc = upload_packed(ar, &b, tmp, 9, N, K);
  if (rc) return rc;
  float* bias = nullptr;
  if (b_host && !transpose_flip) {
    rc = ar.alloc(&bias, Cout);
    if (rc) return rc;
    SMB_CUDA_CHECK(cudaMemcpyAsync(bias, b_host, Cout * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  rc = launch_act_from_nchw(x, a, st);
  if (rc) return rc;
  Epilogue ep;
  ep.bias = bias;
  ep.relu = relu;
  ep.out_hi = o.hi;
  ep.out_lo = o.lo;
  rc = igemm(impl, a, b, ep, st);
  if (rc) return rc;
  rc = launch_act_to_nchw(o, y, st);
  if (rc) return rc;
  SMB_CUDA_CHECK(cudaStreamSynchronize(st));
  return SMB_OK;
}

int smb_unit_conv3x3_fused(const float* x, int Cin, int H, int W, const float* w_host, int Cout, const float* f,
                           const float* g_host, const float* rowmask, float* y, void* stream) {
  SMB_REQUIRE(x && w_host && f && g_host && y, "unit_conv3x3_fused: null argument");
  SMB_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "unit_conv3x3_fused: Cin=%d and Cout=%d must be multiples of 64", Cin, Cout);
  cudaStream_t st = (cudaStream_t)stream;
  DeviceArena ar;
  Act a, o;
  FusedTerm ft;
  int rc = ar.alloc_act(&a, H, W, Cin);
  if (rc) return rc;
  rc = ar.alloc_act(&o, H, W, Cout);
  if (rc) return rc;
  rc = ar.alloc_act(&ft.f, H, W, Cout);
  if (rc) return rc;
  std::vector<float> tmp;
  PackedB b;
  pack_fwd(w_host, Cout, Cin, tmp);
  rc = upload_packed(ar, &b, tmp, 9, Cout, Cin);
  if (rc) return rc;
  tmp.assign(g_host, g_host + (size_t)Cout * Cout);
  rc = upload_packed(ar, &ft.g, tmp, 1, Cout, Cout);
  if (rc) return rc;
  ft.rowmask = rowmask;
  rc = launch_act_from_nchw(x, a, st);
  if (rc) return rc;
  rc = launch_act_from_nchw(f, ft.f, st);
  if (rc) return rc;
  Epilogue ep;
  ep.out_hi = o.hi;
  ep.out_lo = o.lo;
  rc = igemm(IMPL_TC_PH, a, b, ep, st, &ft);
  if (rc) return rc;
  rc = launch_act_to_nchw(o, y, st);
  if (rc) return rc;
  SMB_CUDA_CHECK(cudaStreamSynchronize(st));
  return SMB_OK;
}

int smb_unit_maxpool(const float* x, int C, int H, int W, float* y, void* stream) {
  SMB_REQUIRE(x && y && C % 8 == 0, "unit_maxpool: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DeviceArena ar;
  Act a, o;
  int rc = ar.alloc_act(&a, H, W, C);
  if (rc) return rc;
  rc = ar.alloc_act(&o, H / 2, W / 2, C);
  if (rc) return rc;
  rc = launch_act_from_nchw(x, a, st);
  if (rc) return rc;
  rc = launch_maxpool_fwd(a, o, st);
  if (rc) return rc;
  rc = launch_act_to_nchw(o, y, st);
  if (rc) return rc;
  SMB_CUDA_CHECK(cudaStreamSynchronize(st));
  return SMB_OK;
}

namespace smb {
__global__ void nchw_to_nhwc_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t P) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * C) return;
  const int64_t p = i / C;
  const int c = (int)(i % C);
  dst[i] = src[(int64_t)c * P + p];
}
}  // namespace smb

int smb_unit_maxpool_bwd(const float* g, const float* y, int C, int H, int W, float* dx, void* stream) {
  SMB_REQUIRE(g && y && dx && C % 8 == 0, "unit_maxpool_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DeviceArena ar;
  Act ya, dz;
  int rc = ar.alloc_act(&ya, H, W, C);
  if (rc) return rc;
  rc = ar.alloc_act(&dz, H, W, C);
  if (rc) return rc;
  float* g_nhwc = nullptr;
  const int64_t Pp = (int64_t)(H / 2) * (W / 2);
  rc = ar.alloc(&g_nhwc, Pp * C);
  if (rc) return rc;
  rc = launch_act_from_nchw(y, ya, st);
  if (rc) return rc;
  if (Pp > 0) {
    SMB_LAUNCH(smb::nchw_to_nhwc_f32_kernel, (unsigned)ceil_div64(Pp * C, 256), 256, 0, st, g, g_nhwc, C, Pp);
  }
  rc = launch_maxpool_bwd_relu(g_nhwc, nullptr, ya, dz, st);
  if (rc) return rc;
  rc = launch_act_to_nchw(dz, dx, st);
  if (rc) return rc;
  SMB_CUDA_CHECK(cudaStreamSynchronize(st));
  return SMB_OK;
}

int smb_unit_gram(int impl, const float* f, int C, int H, int W, const float* rowmask, float inv_n, float* G,
                  void* stream) {
  SMB_REQUIRE(f && G && C % 64 == 0, "unit_gram: C=%d must be a multiple of 64", C);
  cudaStream_t st = (cudaStream_t)stream;
  DeviceArena ar;
  Act a, m;
  int rc = ar.alloc_act(&a, H, W, C);
  if (rc) return rc;
  rc = launch_act_from_nchw(f, a, st);
  if (rc) return rc;
  if (rowmask && impl != IMPL_TC) {
    rc = ar.alloc_act(&m, H, W, C);
    if (rc) return rc;
  }
  const int ns = gram_num_splits(a.pixels(), C, impl);
  float* partial = nullptr;
  rc = ar.alloc(&partial, (int64_t)ns * C * C);
  if (rc) return rc;
  rc = gram(impl, a, rowmask, m, partial, ns, st);
  if (rc) return rc;
  const int64_t CC = (int64_t)C * C;
  SMB_LAUNCH(smb::gram_reduce_kernel, (unsigned)std::min<int64_t>(ceil_div64(CC, 256), 592), 256, 0, st, partial, ns,
             CC, inv_n, G);
  SMB_CUDA_CHECK(cudaStreamSynchronize(st));
  return SMB_OK;
}

}  // extern "C"
