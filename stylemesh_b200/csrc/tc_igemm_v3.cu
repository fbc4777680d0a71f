// igemm_tc3_kernel — the CTA-PAIR (tcgen05 cta_group::2) variant of the persistent stream-K implicit GEMM, for the
// wide layers (N multiple of 256: VGG blocks 3-5, 70 % of the conv FLOPs).
//
// Why: igemm_tc2<256> streams 96 KB of operands per 1536 MMA-cycles and only fits a 2-stage ring; ncu shows the
// tensor pipe 55-67 % active with DRAM < 4 % and L2 ~30 % — the SM waits on operand ingest / pipeline depth.
// A CTA pair computes a 256-pixel x 256-output tile with ONE MMA stream (M = 256): each SM loads its own 128 pixel
// rows of A and only HALF of B (128 of the 256 output rows); the tensor cores read the other half from the peer's
// shared memory.  64 KB per SM per K-chunk instead of 96 KB -> a 3-stage ring and 2/3 of the ingest per MMA.
//
// Everything else is igemm_tc2: TMA boxes with out-of-bounds zero fill as conv padding, bf16 hi/lo three-pass
// products into main/corr TMEM accumulators, stream-K over (pair-tile, K-iteration) units with first-K-part
// ownership and epoch-flag fix-ups, fused epilogue.
#include <cstdlib>

#include "tc_common.cuh"
#include "smb_epilogue.cuh"
#include "smb_kernels.h"

namespace smb {
using namespace tc;

constexpr int I3_THREADS = 192;
constexpr int I3_BM = 128;                    // rows per CTA (the pair covers 256)
constexpr int I3_BK = 64;
constexpr int I3_A_BYTES = I3_BM * I3_BK * 2;           // 16 KiB per plane
constexpr int I3_SMEM_EXTRA = 1024 + 256;
constexpr int I3_TMEM_COLS = 512;             // NBUF x (main + corr)

template <int BN>
struct I3Cfg {
  static constexpr int B_BYTES = (BN / 2) * I3_BK * 2;                 // this CTA's half of B, per plane
  static constexpr int STAGE_BYTES = 2 * I3_A_BYTES + 2 * B_BYTES;     // 48 KiB (BN=128) / 64 KiB (BN=256)
  static constexpr int STAGES = 196608 / STAGE_BYTES;                  // 4 / 3
  static constexpr int NBUF = (512 / (2 * BN)) >= 2 ? 2 : 1;           // double-buffered accumulators for BN=128
};
constexpr uint32_t I3_PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair

struct IGemm3Params {
  int H, W, TH, TW, tiles_x, tiles_m, tiles_n, kchunks, taps, N;
  int ipt;
  long long total_units;      // pair_tiles * ipt
  float* ws;                  // [grid][128][256] fp32 partial tiles (indexed by CTA id)
  unsigned int* flags;        // [grid]
  unsigned int epoch;
  Epilogue ep;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  // executed by both CTAs; the transaction bytes are credited to the barrier of CTA 0 of the pair
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & I3_PEER_MASK), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (after all previously issued MMAs completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu3(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu3(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I3_THREADS, 1)
igemm_tc3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const IGemm3Params prm) {
  using Cfg = I3Cfg<BN>;
  constexpr int I3_BN = BN;
  constexpr int I3_B_BYTES = Cfg::B_BYTES;
  constexpr int I3_STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int I3_STAGES = Cfg::STAGES;
  const uint32_t rank = cluster_ctarank();            // 0 = leader (issues the MMAs), 1 = peer
  const long long G = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const long long cta = blockIdx.x;
  const long long u0 = pair * prm.total_units / G, u1 = (pair + 1) * prm.total_units / G;   // G <= total_units

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + I3_STAGES * I3_STAGE_BYTES);   // used in the leader
  uint64_t* empty_bar = full_bar + I3_STAGES;                                           // both CTAs
  uint64_t* tmem_full_bar = empty_bar + I3_STAGES;                                      // [2] both CTAs
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                                         // [2] used in the leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ipt = prm.ipt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < I3_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 8);      // 4 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, I3_TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // peer barriers are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own A rows, own half of B) =====================
    if (elect_one()) {
      long long g = 0;
      for (long long u = u0; u < u1; ++u, ++g) {
        const int stage = (int)(g % I3_STAGES);
        const uint32_t phase = (uint32_t)(g / I3_STAGES) & 1u;
        const int tile = (int)(u / ipt), it = (int)(u % ipt);
        const int pm = tile / prm.tiles_n, n_tile = tile % prm.tiles_n;
        const int m_tile = 2 * pm + (int)rank;
        const int y0 = (m_tile / prm.tiles_x) * prm.TH, x0 = (m_tile % prm.tiles_x) * prm.TW;
        const int tap = it / prm.kchunks, kc = it % prm.kchunks;
        const int dy = (prm.taps == 9) ? tap / 3 - 1 : 0;
        const int dx = (prm.taps == 9) ? tap % 3 - 1 : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1u, 31);
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * I3_STAGE_BYTES);   // bytes of both CTAs
        uint8_t* st = smem + stage * I3_STAGE_BYTES;
        const int nb = n_tile * I3_BN + (int)rank * (I3_BN / 2);
        tma_load_3d_2sm(st, &tmA_hi, &full_bar[stage], kc * I3_BK, x0 + dx, y0 + dy);
        tma_load_3d_2sm(st + I3_A_BYTES, &tmA_lo, &full_bar[stage], kc * I3_BK, x0 + dx, y0 + dy);
        tma_load_3d_2sm(st + 2 * I3_A_BYTES, &tmB_hi, &full_bar[stage], kc * I3_BK, nb, tap);
        tma_load_3d_2sm(st + 2 * I3_A_BYTES + I3_B_BYTES, &tmB_lo, &full_bar[stage], kc * I3_BK, nb, tap);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * I3_BM, I3_BN, 0, 0);
      long long g = 0;
      int seg = 0;
      for (long long u = u0; u < u1; ++seg) {
        const int ks = (int)(u % ipt);
        const long long left = u1 - u;
        const int ke = (left < (long long)(ipt - ks)) ? ks + (int)left : ipt;
        const int buf = seg % Cfg::NBUF;
        const uint32_t use = (uint32_t)(seg / Cfg::NBUF);
        mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u, 32);          // both epilogues drained this buffer
        tc_fence_after();
        const uint32_t t_main = tmem_base + (uint32_t)(buf * 2 * I3_BN);
        const uint32_t t_corr = t_main + (uint32_t)I3_BN;
        for (int it = ks; it < ke; ++it, ++g) {
          const int stage = (int)(g % I3_STAGES);
          const uint32_t phase = (uint32_t)(g / I3_STAGES) & 1u;
          mbar_wait(&full_bar[stage], phase, 33);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * I3_STAGE_BYTES);
          const uint32_t a_lo = a_hi + I3_A_BYTES;
          const uint32_t b_hi = a_hi + 2 * I3_A_BYTES;
          const uint32_t b_lo = b_hi + I3_B_BYTES;
#pragma unroll
          for (int k = 0; k < I3_BK / 16; ++k) {
            const uint64_t dah = make_smem_desc_sw128(a_hi + k * 32, 16, 1024);
            const uint64_t dal = make_smem_desc_sw128(a_lo + k * 32, 16, 1024);
            const uint64_t dbh = make_smem_desc_sw128(b_hi + k * 32, 16, 1024);
            const uint64_t dbl = make_smem_desc_sw128(b_lo + k * 32, 16, 1024);
            const uint32_t acc = (uint32_t)((it > ks) || (k > 0));
            umma_f16_2sm(t_corr, dal, dbh, idesc, acc);
            umma_f16_2sm(t_corr, dah, dbl, idesc, 1u);
            umma_f16_2sm(t_main, dah, dbh, idesc, acc);
          }
          umma_commit_2sm_mc(&empty_bar[stage]);      // frees this stage in BOTH CTAs
        }
        umma_commit_2sm_mc(&tmem_full_bar[buf]);      // accumulators complete, both CTAs
        u += (ke - ks);
      }
    }
  } else {
    // ===================== epilogue warps (each CTA drains its own 128 TMEM lanes) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* my_slot = prm.ws + (size_t)cta * I3_BM * I3_BN;
    int seg = 0;
    for (long long u = u0; u < u1; ++seg) {
      const int tile = (int)(u / ipt);
      const int ks = (int)(u % ipt);
      const long long left = u1 - u;
      const int ke = (left < (long long)(ipt - ks)) ? ks + (int)left : ipt;
      const bool owner = (ks == 0);
      const int buf = seg % Cfg::NBUF;
      const uint32_t use = (uint32_t)(seg / Cfg::NBUF);
      const int pm = tile / prm.tiles_n, n_tile = tile % prm.tiles_n;
      const int m_tile = 2 * pm + (int)rank;
      const int y0 = (m_tile / prm.tiles_x) * prm.TH, x0 = (m_tile % prm.tiles_x) * prm.TW;
      const int yy = y0 + row / prm.TW, xx = x0 + row % prm.TW;
      const bool valid = (m_tile < prm.tiles_m) && (yy < prm.H) && (xx < prm.W);
      const int64_t p = (int64_t)yy * prm.W + xx;
      const int n0 = n_tile * I3_BN;

      int npeer = 0;
      if (owner && ke < ipt) {
        const long long tile_end = (long long)(tile + 1) * ipt;
        long long c = pair + 1;
        while (c < G && c * prm.total_units / G < tile_end) {
          const unsigned int* f = prm.flags + 2 * c + rank;           // same-rank CTA of the later pair
          if (lane == 0) {
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile const unsigned int*>(f) != prm.epoch) {
              __nanosleep(64);
              if (clock64() - t0 > 4000000000LL) {
                printf("[smb] igemm3 stream-K watchdog: CTA %d waiting for partial of CTA %d\n", (int)cta,
                       (int)(2 * c + rank));
                asm volatile("trap;");
              }
            }
          }
          __syncwarp();
          (void)ld_acquire_gpu3(f);
          ++npeer;
          ++c;
        }
      }

      mbar_wait(&tmem_full_bar[buf], use & 1u, 34);
      tc_fence_after();
      const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * I3_BN);
      const uint32_t t_corr = t_main + (uint32_t)I3_BN;
#pragma unroll 1
      for (int c = 0; c < I3_BN; c += 32) {
        uint32_t rm[32], rc[32];
        tmem_ld_32x32(t_main + (uint32_t)c, rm);
        tmem_ld_32x32(t_corr + (uint32_t)c, rc);
        tmem_ld_wait();
        if (c + 32 >= I3_BN) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&tmem_empty_bar[buf], 0);    // the leader's barrier collects 8 arrivals
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]);
        for (int k = 1; k <= npeer; ++k) {
          const float4* src = reinterpret_cast<const float4*>(prm.ws + (size_t)(cta + 2 * k) * I3_BM * I3_BN +
                                                              (size_t)row * I3_BN + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = src[j];
            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
          }
        }
        if (owner) {
          if (valid) epilogue_store<32>(prm.ep, p, n0 + c, prm.N, v);
        } else {
          float4* dst = reinterpret_cast<float4*>(my_slot + (size_t)row * I3_BN + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      if (!owner) {
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) st_release_gpu3(prm.flags + cta, prm.epoch);
      }
      u += (ke - ks);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // the peer's shared memory / TMEM stay alive until both CTAs are done
  if (warp == 1) tmem_dealloc_2sm(tmem_base, I3_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
int igemm_streamk_workspace(float** ws, unsigned int** flags, unsigned int* epoch);   // tc_igemm_v2.cu

static void pick_patch3(int H, int W, int& TH, int& TW) {
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

template <int BN>
static int launch_igemm_tc3_bn(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st) {
  using Cfg = I3Cfg<BN>;
  constexpr int I3_BN = BN;
  constexpr int I3_STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int I3_STAGES = Cfg::STAGES;
  IGemm3Params prm;
  int rc = igemm_streamk_workspace(&prm.ws, &prm.flags, &prm.epoch);
  if (rc) return rc;
  prm.H = a.H;
  prm.W = a.W;
  pick_patch3(a.H, a.W, prm.TH, prm.TW);
  prm.tiles_x = ceil_div(a.W, prm.TW);
  prm.tiles_m = prm.tiles_x * ceil_div(a.H, prm.TH);
  prm.tiles_n = b.N / I3_BN;
  prm.kchunks = b.K / I3_BK;
  prm.taps = b.taps;
  prm.N = b.N;
  prm.ipt = prm.taps * prm.kchunks;
  const long long pair_tiles = (long long)ceil_div(prm.tiles_m, 2) * prm.tiles_n;
  prm.total_units = pair_tiles * prm.ipt;
  prm.ep = ep;

  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  {
    const uint64_t dims[3] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H};
    const uint64_t strides[2] = {(uint64_t)a.C * 2, (uint64_t)a.W * a.C * 2};
    const uint32_t box[3] = {(uint32_t)I3_BK, (uint32_t)prm.TW, (uint32_t)prm.TH};
    rc = make_tmap_bf16(&tmA_hi, a.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmA_lo, a.lo, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)b.K, (uint64_t)b.N, (uint64_t)b.taps};
    const uint64_t strides[2] = {(uint64_t)b.K * 2, (uint64_t)b.N * b.K * 2};
    const uint32_t box[3] = {(uint32_t)I3_BK, (uint32_t)(I3_BN / 2), 1u};
    rc = make_tmap_bf16(&tmB_hi, b.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB_lo, b.lo, 3, dims, strides, box);
    if (rc) return rc;
  }
  const int smem_bytes = I3_STAGES * I3_STAGE_BYTES + I3_SMEM_EXTRA;
  static bool attr_set = false;
  if (!attr_set) {
    SMB_CUDA_CHECK(cudaFuncSetAttribute(igemm_tc3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SMB_CUDA_CHECK(cudaGetDevice(&dev));
    SMB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (num_sms > 148) num_sms = 148;
  }
  long long pairs = std::min<long long>(num_sms / 2, prm.total_units);
  pairs = std::max<long long>(1, std::min<long long>(pairs, std::max<long long>(pair_tiles, prm.total_units / 16)));
  igemm_tc3_kernel<BN><<<(unsigned)(2 * pairs), I3_THREADS, smem_bytes, st>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, prm);
  SMB_LAUNCH_CHECK();
  return SMB_OK;
}

int launch_igemm_tc3(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st) {
  SMB_REQUIRE(b.taps == 9 || b.taps == 1, "igemm_tc3: taps must be 1 or 9");
  SMB_REQUIRE(a.C == b.K && b.K % I3_BK == 0 && b.N % 128 == 0,
              "igemm_tc3: K=%d must be a multiple of 64 and N=%d a multiple of 128", b.K, b.N);
  if (a.pixels() == 0) return SMB_OK;
  static int want_bn = 0;
  if (!want_bn) {
    const char* e = getenv("SMB_IGEMM_PAIR_BN");
    want_bn = e ? atoi(e) : 128;                 // 128: double-buffered accumulators (epilogue overlaps the MMAs)
    if (want_bn != 128 && want_bn != 256) want_bn = 128;
  }
  if (want_bn == 256 && b.N % 256 == 0) return launch_igemm_tc3_bn<256>(a, b, ep, st);
  return launch_igemm_tc3_bn<128>(a, b, ep, st);
}

}  // namespace smb
