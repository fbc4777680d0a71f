// igemm_tc2_kernel — persistent, stream-K, warp-specialised tcgen05 implicit GEMM (second generation of the conv
// kernel in tc_kernels.cu; same math, same operands, same epilogue).
//
// What changed against igemm_tc_kernel and why (numbers from profiles/r01_*.md):
//  * one CTA per SM for the whole launch (grid = #SMs), each owning a CONTIGUOUS range of (tile, K-iteration) work
//    units ("stream-K"): conv3_x has 150 output tiles for 148 SMs and conv4_x 160 — a tile-per-CTA launch runs two
//    waves at ~51-54 % occupancy; with stream-K every SM executes total/148 units.  A tile that straddles two (or
//    more) CTAs is finished by the CTA that owns its FIRST K-iteration: the others park their fp32 partial tile in a
//    global workspace slot and publish an epoch flag; the owner adds the partials (fixed order => deterministic).
//  * TMEM accumulators are double-buffered (BN <= 128): the epilogue of tile i overlaps the MMAs of tile i+1, and
//    barrier init / TMEM allocation / descriptor prefetch happen once per launch instead of once per tile.
//  * two accumulators per tile: the large hi*hi products go to `main`, the 2^-8-smaller cross terms (lo*hi, hi*lo)
//    to `corr`, summed in fp32 registers by the epilogue.  The tensor core truncates (round-toward-zero) on every
//    accumulate, a systematic bias proportional to the number of accumulate steps times ulp(accumulator)
//    (measured -2.4e-5 relative on an all-positive K=4608 conv); keeping the small terms out of the large
//    accumulator cuts the chain on `main` to a third and makes the truncation on `corr` negligible.
#include <cstdlib>

#include "tc_common.cuh"
#include "smb_epilogue.cuh"
#include "smb_kernels.h"

namespace smb {
using namespace tc;

constexpr int I2_THREADS = 192;
constexpr int I2_BM = 128;
constexpr int I2_BK = 64;
constexpr int I2_A_BYTES = I2_BM * I2_BK * 2;
constexpr int I2_SMEM_BUDGET = 196608;
constexpr int I2_SMEM_EXTRA = 1024 + 256;
constexpr int I2_MAX_GRID = 148;

struct IGemm2Params {
  int H, W, TH, TW, tiles_x, tiles_n, kchunks, taps, N;
  int ipt;                    // K-iterations (taps x K-chunks) per output tile
  long long total_units;      // tiles * ipt
  float* ws;                  // [grid][128][BN] fp32 partial tiles
  unsigned int* flags;        // [grid]
  unsigned int epoch;
  int dbg;                    // timing experiments only (SMB_IGEMM_DEBUG): 1 = hi*hi MMA only, 2 = no MMAs at all
  unsigned long long* trace;  // optional [grid][16] per-CTA timeline (smb_debug_set_igemm_trace), nullptr = off
  Epilogue ep;
};

template <int BN>
struct I2Cfg {
  static constexpr int B_BYTES = BN * I2_BK * 2;
  static constexpr int STAGE_BYTES = 2 * I2_A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = I2_SMEM_BUDGET / STAGE_BYTES;
  static constexpr int NBUF = (512 / (2 * BN)) >= 2 ? 2 : 1;      // tile buffers in TMEM (each = main + corr)
  static constexpr int TMEM_COLS = (NBUF * 2 * BN) <= 256 ? 256 : 512;
  static_assert(STAGES >= 2, "need at least a double buffer");
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// per-CTA timeline slots (tools/gpu_trace_probe.py prints them)
enum : int { TR_GT_IN = 0, TR_GT_OUT, TR_CLK_IN, TR_CLK_PROLOGUE, TR_CLK_TMA_END, TR_CLK_MMA_FIRST, TR_CLK_MMA_END,
             TR_CLK_EPI_FIRST, TR_CLK_EPI_END, TR_W_FLAGS, TR_W_TMEM_FULL, TR_W_FULL, TR_W_TMEM_EMPTY, TR_W_EMPTY,
             TR_CLK_OUT, TR_SMID };
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(I2_THREADS, 1)
igemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const IGemm2Params prm) {
  using Cfg = I2Cfg<BN>;
  const long long G = gridDim.x, cta = blockIdx.x;
  const long long u0 = cta * prm.total_units / G, u1 = (cta + 1) * prm.total_units / G;
  if (u0 >= u1) {                             // more CTAs than work units (uniform exit, nothing allocated yet)
    pdl_sync();
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;      // [NBUF]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;           // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ipt = prm.ipt;
  unsigned long long* tr = prm.trace ? prm.trace + (size_t)cta * 16 : nullptr;
  if (tr && threadIdx.x == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr[TR_GT_IN] = global_ns();
    tr[TR_CLK_IN] = (unsigned long long)clock64();
    tr[TR_SMID] = smid;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);      // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();                                // everything above overlaps the tail of the preceding launch
  if (tr && threadIdx.x == 0) tr[TR_CLK_PROLOGUE] = (unsigned long long)clock64();

  if (warp == 0) {
    // ===================== TMA producer: streams every K-iteration of the CTA's unit range =====================
    if (elect_one()) {
      long long g = 0;
      long long w_empty = 0;
      for (long long u = u0; u < u1; ++u, ++g) {
        const int stage = (int)(g % Cfg::STAGES);
        const uint32_t phase = (uint32_t)(g / Cfg::STAGES) & 1u;
        const int tile = (int)(u / ipt), it = (int)(u % ipt);
        const int m_tile = tile / prm.tiles_n, n_tile = tile % prm.tiles_n;
        const int y0 = (m_tile / prm.tiles_x) * prm.TH, x0 = (m_tile % prm.tiles_x) * prm.TW;
        const int tap = it / prm.kchunks, kc = it % prm.kchunks;
        const int dy = (prm.taps == 9) ? tap / 3 - 1 : 0;
        const int dx = (prm.taps == 9) ? tap % 3 - 1 : 0;
        const long long tw0 = tr ? clock64() : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1u, 21);
        if (tr) w_empty += clock64() - tw0;
        if (prm.dbg == 4) {                 // timing experiment: no operand traffic at all
          mbar_arrive(&full_bar[stage]);
          continue;
        }
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
        tma_load_3d(st, &tmA_hi, &full_bar[stage], kc * I2_BK, x0 + dx, y0 + dy);
        tma_load_3d(st + I2_A_BYTES, &tmA_lo, &full_bar[stage], kc * I2_BK, x0 + dx, y0 + dy);
        tma_load_3d(st + 2 * I2_A_BYTES, &tmB_hi, &full_bar[stage], kc * I2_BK, n_tile * BN, tap);
        tma_load_3d(st + 2 * I2_A_BYTES + Cfg::B_BYTES, &tmB_lo, &full_bar[stage], kc * I2_BK, n_tile * BN, tap);
      }
      if (tr) {
        tr[TR_CLK_TMA_END] = (unsigned long long)clock64();
        tr[TR_W_EMPTY] = (unsigned long long)w_empty;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(I2_BM, BN, 0, 0);
      long long g = 0;
      int seg = 0;
      long long w_full = 0, w_tempty = 0;
      for (long long u = u0; u < u1; ++seg) {
        const int ks = (int)(u % ipt);
        const long long left = u1 - u;
        const int ke = (left < (long long)(ipt - ks)) ? ks + (int)left : ipt;
        const int buf = seg % Cfg::NBUF;
        const uint32_t use = (uint32_t)(seg / Cfg::NBUF);
        const long long tw1 = tr ? clock64() : 0;
        mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u, 22);      // epilogue has drained this buffer
        if (tr) w_tempty += clock64() - tw1;
        tc_fence_after();
        const uint32_t t_main = tmem_base + (uint32_t)(buf * 2 * BN);
        const uint32_t t_corr = t_main + (uint32_t)BN;
        for (int it = ks; it < ke; ++it, ++g) {
          const int stage = (int)(g % Cfg::STAGES);
          const uint32_t phase = (uint32_t)(g / Cfg::STAGES) & 1u;
          const long long tw2 = tr ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase, 23);
          if (tr) {
            const long long now = clock64();
            w_full += now - tw2;
            if (g == 0) tr[TR_CLK_MMA_FIRST] = (unsigned long long)now;
          }
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t a_lo = a_hi + I2_A_BYTES;
          const uint32_t b_hi = a_hi + 2 * I2_A_BYTES;
          const uint32_t b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
          for (int k = 0; k < I2_BK / 16; ++k) {
            const uint64_t dah = make_smem_desc_sw128(a_hi + k * 32, 16, 1024);
            const uint64_t dal = make_smem_desc_sw128(a_lo + k * 32, 16, 1024);
            const uint64_t dbh = make_smem_desc_sw128(b_hi + k * 32, 16, 1024);
            const uint64_t dbl = make_smem_desc_sw128(b_lo + k * 32, 16, 1024);
            const uint32_t acc = (uint32_t)((it > ks) || (k > 0));
            if (prm.dbg == 0 || prm.dbg >= 3) {
              umma_f16(t_corr, dal, dbh, idesc, acc);
              umma_f16(t_corr, dah, dbl, idesc, 1u);
            }
            if (prm.dbg <= 1 || prm.dbg >= 3) umma_f16(t_main, dah, dbh, idesc, acc);
          }
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(&tmem_full_bar[buf]);
        u += (ke - ks);
      }
      if (tr) {
        tr[TR_CLK_MMA_END] = (unsigned long long)clock64();
        tr[TR_W_FULL] = (unsigned long long)w_full;
        tr[TR_W_TMEM_EMPTY] = (unsigned long long)w_tempty;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* my_slot = prm.ws + (size_t)cta * I2_BM * BN;
    int seg = 0;
    long long w_flags = 0, w_tfull = 0;
    const bool tr_me = tr && warp == 2 && lane == 0;
    for (long long u = u0; u < u1; ++seg) {
      const int tile = (int)(u / ipt);
      const int ks = (int)(u % ipt);
      const long long left = u1 - u;
      const int ke = (left < (long long)(ipt - ks)) ? ks + (int)left : ipt;
      const int buf = seg % Cfg::NBUF;
      const uint32_t use = (uint32_t)(seg / Cfg::NBUF);
      // Ownership: the CTA holding the FIRST K-part of a tile (ks == 0) finishes it.  In forward unit order that
      // part is the owner's LAST segment, while the CTAs holding the later K-parts meet the tile as the FIRST
      // segment of their range and publish their partial at once — the owner never waits long and there is no
      // CTA-to-CTA serialisation (waiting on lower-indexed CTAs, i.e. on their last segment, would chain them).
      const bool owner = (ks == 0);
      const int m_tile = tile / prm.tiles_n, n_tile = tile % prm.tiles_n;
      const int y0 = (m_tile / prm.tiles_x) * prm.TH, x0 = (m_tile % prm.tiles_x) * prm.TW;
      const int yy = y0 + row / prm.TW, xx = x0 + row % prm.TW;
      const bool valid = (yy < prm.H) && (xx < prm.W);
      const int64_t p = (int64_t)yy * prm.W + xx;
      const int n0 = n_tile * BN;

      // peers holding the remaining K-range of this tile: CTAs cta+1, cta+2, ... whose range starts inside the tile
      int npeer = 0;
      const long long tw3 = tr_me ? clock64() : 0;
      if (owner && ke < ipt) {
        const long long tile_end = (long long)(tile + 1) * ipt;
        long long c = cta + 1;
        while (c < G && c * prm.total_units / G < tile_end) {
          // lane 0 spins with plain volatile loads (an acquire load per poll costs an L1 invalidate, CCTL.IVALL,
          // every iteration: 9 % of all samples in the first ncu capture); one acquire per lane once it is set
          if (lane == 0) {
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile const unsigned int*>(prm.flags + c) != prm.epoch) {
              __nanosleep(64);
              if (clock64() - t0 > 4000000000LL) {
                printf("[smb] igemm2 stream-K watchdog: CTA %d waiting for partial of CTA %d\n", (int)cta, (int)c);
                asm volatile("trap;");
              }
            }
          }
          __syncwarp();
          (void)ld_acquire_gpu(prm.flags + c);
          ++npeer;
          ++c;
        }
      }

      const long long tw4 = tr_me ? clock64() : 0;
      mbar_wait(&tmem_full_bar[buf], use & 1u, 24);
      if (tr_me) {
        w_flags += tw4 - tw3;
        w_tfull += clock64() - tw4;
      }
      tc_fence_after();
      const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN);
      const uint32_t t_corr = t_main + (uint32_t)BN;
      if constexpr (BN == 16) {
        // narrow tile (data gradient of the 3-channel first layer, N padded 3 -> 16): main and corr are adjacent
        // 16-column blocks, one 32-column TMEM load fetches both
        uint32_t rm[32];
        tmem_ld_32x32(t_main, rm);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rm[16 + j]);
        for (int k = 1; k <= npeer; ++k) {
          const float4* src = reinterpret_cast<const float4*>(prm.ws + (size_t)(cta + k) * I2_BM * BN + (size_t)row * BN);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 a = src[j];
            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
          }
        }
        if (owner) {
          if (valid) {
            if (prm.ep.out_planar3) {            // (3, H, W) fp32 image gradient
              const int64_t P = (int64_t)prm.H * prm.W;
              prm.ep.out_planar3[p] = v[0];
              prm.ep.out_planar3[P + p] = v[1];
              prm.ep.out_planar3[2 * P + p] = v[2];
            } else {
              epilogue_store<16>(prm.ep, p, n0, prm.N, v);
            }
          }
        } else {
          float4* dst = reinterpret_cast<float4*>(my_slot + (size_t)row * BN);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t rm[32], rc[32];
        tmem_ld_32x32(t_main + (uint32_t)c, rm);
        tmem_ld_32x32(t_corr + (uint32_t)c, rc);
        tmem_ld_wait();
        if (c + 32 >= BN) {                      // last TMEM read of this buffer: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]);
        for (int k = 1; k <= npeer; ++k) {       // fixed order cta+1, cta+2, ... => deterministic sums
          const float4* src = reinterpret_cast<const float4*>(prm.ws + (size_t)(cta + k) * I2_BM * BN +
                                                              (size_t)row * BN + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = src[j];
            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
          }
        }
        if (owner) {
          if (valid && prm.dbg != 3) epilogue_store<32>(prm.ep, p, n0 + c, prm.N, v);   // dbg 3: no epilogue math/stores
        } else {
          float4* dst = reinterpret_cast<float4*>(my_slot + (size_t)row * BN + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      }
      if (!owner) {
        // publish the partial tile: every epilogue thread's stores -> gpu scope, then one release store of the flag
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) st_release_gpu(prm.flags + cta, prm.epoch);
      }
      if (tr_me && seg == 0) tr[TR_CLK_EPI_FIRST] = (unsigned long long)clock64();
      u += (ke - ks);
    }
    if (tr_me) {
      tr[TR_CLK_EPI_END] = (unsigned long long)clock64();
      tr[TR_W_FLAGS] = (unsigned long long)w_flags;
      tr[TR_W_TMEM_FULL] = (unsigned long long)w_tfull;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (tr && threadIdx.x == 0) {
    tr[TR_CLK_OUT] = (unsigned long long)clock64();
    tr[TR_GT_OUT] = global_ns();
  }
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
static void pick_patch2(int H, int W, int& TH, int& TW) {
  int best_th = 8, best_tw = 16, best_sq = 1 << 30;
  int64_t best_area = -1;
  for (int th = 1; th <= 128; th <<= 1) {
    const int tw = 128 / th;
    const int64_t area = (int64_t)ceil_div(H, th) * th * ceil_div(W, tw) * tw;
    const int sq = (tw > 16) ? tw / 16 : 16 / tw;
    if (best_area < 0 || area < best_area || (area == best_area && sq < best_sq)) {
      best_area = area;
      best_th = th;
      best_tw = tw;
      best_sq = sq;
    }
  }
  TH = best_th;
  TW = best_tw;
}

struct StreamKWorkspace {
  float* ws = nullptr;
  unsigned int* flags = nullptr;
  unsigned int epoch = 0;
  int device = -1;
};
static StreamKWorkspace g_sk;
static unsigned long long* g_trace = nullptr;      // device buffer of >= 148 * 16 u64, set by smb_debug_set_igemm_trace
void set_igemm_trace(unsigned long long* buf) { g_trace = buf; }
unsigned long long* get_igemm_trace() { return g_trace; }

static int ensure_workspace() {
  int dev = 0;
  SMB_CUDA_CHECK(cudaGetDevice(&dev));
  if (g_sk.ws && g_sk.device == dev) return SMB_OK;
  SMB_REQUIRE(g_sk.ws == nullptr, "igemm_tc2: the stream-K workspace is per process and bound to device %d", g_sk.device);
  SMB_CUDA_CHECK(cudaMalloc(&g_sk.ws, (size_t)I2_MAX_GRID * I2_BM * 256 * sizeof(float)));
  SMB_CUDA_CHECK(cudaMalloc(&g_sk.flags, I2_MAX_GRID * sizeof(unsigned int)));
  SMB_CUDA_CHECK(cudaMemset(g_sk.flags, 0, I2_MAX_GRID * sizeof(unsigned int)));
  g_sk.device = dev;
  g_sk.epoch = 0;
  return SMB_OK;
}

// shared with the CTA-pair variant (tc_igemm_v3.cu): one workspace / epoch counter per process
int igemm_streamk_workspace(float** ws, unsigned int** flags, unsigned int* epoch) {
  int rc = ensure_workspace();
  if (rc) return rc;
  *ws = g_sk.ws;
  *flags = g_sk.flags;
  unsigned int e = ++g_sk.epoch;
  if (e == 0) e = ++g_sk.epoch;
  *epoch = e;
  return SMB_OK;
}

template <int BN>
static int launch_igemm_tc2_bn(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st) {
  using Cfg = I2Cfg<BN>;
  int rc = ensure_workspace();
  if (rc) return rc;
  IGemm2Params prm;
  prm.H = a.H;
  prm.W = a.W;
  pick_patch2(a.H, a.W, prm.TH, prm.TW);
  prm.tiles_x = ceil_div(a.W, prm.TW);
  const int tiles_y = ceil_div(a.H, prm.TH);
  prm.tiles_n = b.N / BN;
  prm.kchunks = b.K / I2_BK;
  prm.taps = b.taps;
  prm.N = b.N;
  prm.ipt = prm.taps * prm.kchunks;
  const long long tiles = (long long)prm.tiles_x * tiles_y * prm.tiles_n;
  prm.total_units = tiles * prm.ipt;
  prm.ws = g_sk.ws;
  prm.flags = g_sk.flags;
  prm.epoch = ++g_sk.epoch;
  if (prm.epoch == 0) prm.epoch = ++g_sk.epoch;      // 0 is the "never written" value of the flags
  prm.ep = ep;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("SMB_IGEMM_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  prm.dbg = dbg;
  prm.trace = g_trace;

  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  {
    const uint64_t dims[3] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H};
    const uint64_t strides[2] = {(uint64_t)a.C * 2, (uint64_t)a.W * a.C * 2};
    const uint32_t box[3] = {(uint32_t)I2_BK, (uint32_t)prm.TW, (uint32_t)prm.TH};
    rc = make_tmap_bf16(&tmA_hi, a.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmA_lo, a.lo, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)b.K, (uint64_t)b.N, (uint64_t)b.taps};
    const uint64_t strides[2] = {(uint64_t)b.K * 2, (uint64_t)b.N * b.K * 2};
    const uint32_t box[3] = {(uint32_t)I2_BK, (uint32_t)BN, 1u};
    rc = make_tmap_bf16(&tmB_hi, b.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB_lo, b.lo, 3, dims, strides, box);
    if (rc) return rc;
  }
  const int smem_bytes = Cfg::STAGES * Cfg::STAGE_BYTES + I2_SMEM_EXTRA;
  static bool attr_set = false;
  if (!attr_set) {
    SMB_CUDA_CHECK(cudaFuncSetAttribute(igemm_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SMB_CUDA_CHECK(cudaGetDevice(&dev));
    SMB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (num_sms > I2_MAX_GRID) num_sms = I2_MAX_GRID;
  }
  // every CTA must be co-resident (owners wait for higher-indexed CTAs): 1 CTA/SM by shared memory, grid <= #SMs,
  // and the engine issues these launches on one stream with nothing else running on the device
  // small problems: do not shred a handful of tiles over all SMs (every extra split costs a 128 x BN fp32 partial
  // write + read); keep at least ~16 K-iterations per CTA unless that would leave fewer CTAs than tiles
  long long want = std::max<long long>(tiles, prm.total_units / 16);
  const int grid = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(num_sms, prm.total_units), want));
  SMB_LAUNCH(igemm_tc2_kernel<BN>, grid, I2_THREADS, smem_bytes, st, tmA_hi, tmA_lo, tmB_hi, tmB_lo, prm);
  return SMB_OK;
}

int launch_igemm_tc2(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st) {
  SMB_REQUIRE(b.taps == 9 || b.taps == 1, "igemm_tc2: taps must be 1 or 9");
  SMB_REQUIRE(a.C == b.K && b.K % I2_BK == 0, "igemm_tc2: K=%d must equal the activation channels and be a multiple of 64",
              b.K);
  SMB_REQUIRE(b.N % 64 == 0 || b.N == 16, "igemm_tc2: N=%d must be a multiple of 64 (or the padded 16)", b.N);
  if (a.pixels() == 0) return SMB_OK;
  if (b.N == 16) return launch_igemm_tc2_bn<16>(a, b, ep, st);
  // widest N tile: 256 halves the operand traffic per MMA but its two accumulators (main + corr) fill the TMEM,
  // so the epilogue cannot overlap the next tile; 128 keeps a double-buffered TMEM.  SMB_IGEMM_MAX_BN overrides.
  static int max_bn = 0;
  if (!max_bn) {
    const char* e = getenv("SMB_IGEMM_MAX_BN");
    max_bn = e ? atoi(e) : 128;      // measured on B200: 128 -> 3.10 ms / step, 256 -> 3.41 ms / step
    if (max_bn != 64 && max_bn != 128 && max_bn != 256) max_bn = 128;
  }
  if (b.N % 256 == 0 && max_bn >= 256) return launch_igemm_tc2_bn<256>(a, b, ep, st);
  if (b.N % 128 == 0 && max_bn >= 128) return launch_igemm_tc2_bn<128>(a, b, ep, st);
  return launch_igemm_tc2_bn<64>(a, b, ep, st);
}

}  // namespace smb
