// tcgen05 (5th-gen tensor core) kernels for sm_100a:
//
//  (the convolutions live in tc_igemm_v5.cu [pair + halo, the product kernel], tc_igemm_v2.cu [stream-K, 1x1 and odd
//   shapes] and tc_conv_first.cu; all operands are channels-last bf16 hi/lo planes, three bf16 MMAs per product:
//   hi*hi + lo*hi + hi*lo, ~2^-16 relative product error at 1/3 of the bf16 tensor rate)
//
//  gram_tc_kernel — masked Gram partials  G_s[i][j] = sum_{p in split s} Fm[p][i] Fm[p][j]  with both operands
//      MN-major straight out of the channels-last planes (TMA boxes {64 ch, 64 pixels}), split over pixel ranges
//      across CTAs, deterministic per-split partial tiles reduced by gram_mse_kernel.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quarter = warp_id % 4).  One output tile per CTA; smem ring of STAGES stages with full/empty mbarriers.
#include "tc_common.cuh"
#include "smb_epilogue.cuh"
#include "smb_kernels.h"

namespace smb {
using namespace tc;

constexpr int TC_THREADS = 192;
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                       // bf16 elements per K chunk = one 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KiB per plane
constexpr int TC_SMEM_BUDGET = 196608;          // ring bytes (+ alignment slack + barriers below)
constexpr int TC_SMEM_EXTRA = 1024 + 256;

// ------------------------------------------------------------------------------------------------------------
// Gram partials (both operands MN-major)
// ------------------------------------------------------------------------------------------------------------
constexpr int GR_KP = 64;                        // pixels per stage
constexpr int GR_BOX_BYTES = GR_KP * 64 * 2;     // one {64 ch, 64 pixel} box = 8 KiB

struct GramTcParams {
  int C, nsplit;
  int64_t P, pix_per_split;
  float* partial;          // [nsplit][C][C]
  const float* rowmask;    // [P] of {0, 1} or nullptr: pixels with mask 0 do not contribute (cs:136-143 compaction)
};

// ALIAS: the whole Gram matrix is ONE tile (C == BN <= 128), so the A and the B operand are the same channels of the
// same pixels: the stage holds them once and both descriptors point at it.  The r11 / r21 Grams are HBM-bound
// reductions; with separate A / B copies (and the duplicated upper half of A at C = 64) every feature byte crossed
// L2 -> shared memory three times and only a third of the bytes in flight were useful.
template <int BN, bool ALIAS>
struct GramCfg {
  static constexpr int A_BYTES = ALIAS ? (BN / 64) * GR_BOX_BYTES : 2 * GR_BOX_BYTES;   // per plane (non-alias: 128 ch)
  static constexpr int B_BYTES = ALIAS ? 0 : (BN / 64) * GR_BOX_BYTES;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = TC_SMEM_BUDGET / STAGE_BYTES < 8 ? TC_SMEM_BUDGET / STAGE_BYTES : 8;
  // C == 64 runs the M = 128 MMA with rows 64..127 of A read from the 8 KiB after the tile (results dropped): the last
  // stage's lo plane needs that much readable shared memory behind it
  static constexpr int SLACK = (ALIAS && BN == 64) ? GR_BOX_BYTES : 0;
  static constexpr int SMEM = STAGES * STAGE_BYTES + SLACK + TC_SMEM_EXTRA;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN, bool ALIAS>
__global__ void __launch_bounds__(TC_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const GramTcParams prm) {
  using Cfg = GramCfg<BN, ALIAS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::SLACK);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint64_t* masked_bar = tmem_full_bar + 1;              // [STAGES] count 4: masked pixel rows of B are zeroed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(masked_bar + Cfg::STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * TC_BM, j0 = blockIdx.y * BN, split = blockIdx.z;
  const int C = prm.C;
  const int64_t pbeg = (int64_t)split * prm.pix_per_split;
  const int64_t pend = (pbeg + prm.pix_per_split < prm.P) ? pbeg + prm.pix_per_split : prm.P;
  const int num_iters = (pend > pbeg) ? (int)((pend - pbeg + GR_KP - 1) / GR_KP) : 0;
  // rows of the A tile that exist (C == 64 uses the upper half of a 128-row MMA as a duplicate that is dropped)
  const int a_groups = (C - i0 >= 128) ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&masked_bar[s], 4);         // one arrival per mask warp
    }
    mbar_init(tmem_full_bar, 1);
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
  pdl_sync();

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < num_iters; ++it) {
        const int stage = it % Cfg::STAGES;
        const uint32_t phase = (uint32_t)(it / Cfg::STAGES) & 1u;
        mbar_wait(&empty_bar[stage], phase ^ 1u, 11);
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        const int pix = (int)(pbeg + (int64_t)it * GR_KP);
        uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
        for (int g = 0; g < Cfg::A_BYTES / GR_BOX_BYTES; ++g) {
          const int ch = i0 + ((g < a_groups) ? g * 64 : 0);
          tma_load_2d(st + g * GR_BOX_BYTES, &tm_hi, &full_bar[stage], ch, pix);
          tma_load_2d(st + Cfg::A_BYTES + g * GR_BOX_BYTES, &tm_lo, &full_bar[stage], ch, pix);
        }
        if constexpr (!ALIAS) {
          for (int g = 0; g < BN / 64; ++g) {
            tma_load_2d(st + 2 * Cfg::A_BYTES + g * GR_BOX_BYTES, &tm_hi, &full_bar[stage], j0 + g * 64, pix);
            tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES + g * GR_BOX_BYTES, &tm_lo, &full_bar[stage], j0 + g * 64,
                        pix);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN, 1, 1);
      for (int it = 0; it < num_iters; ++it) {
        const int stage = it % Cfg::STAGES;
        const uint32_t phase = (uint32_t)(it / Cfg::STAGES) & 1u;
        mbar_wait(prm.rowmask ? &masked_bar[stage] : &full_bar[stage], phase, 12);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t a_lo = a_hi + Cfg::A_BYTES;
        const uint32_t b_hi = ALIAS ? a_hi : a_hi + 2 * Cfg::A_BYTES;
        const uint32_t b_lo = ALIAS ? a_lo : b_hi + Cfg::B_BYTES;
#pragma unroll
        for (int k = 0; k < GR_KP / 16; ++k) {
          // MN-major SWIZZLE_128B: 64-channel groups GR_BOX_BYTES apart (LBO), 8-pixel groups 1024 B apart (SBO),
          // a K=16 step = 16 pixel rows = 2048 B
          const uint64_t dah = make_smem_desc_sw128(a_hi + k * 2048, GR_BOX_BYTES, 1024);
          const uint64_t dal = make_smem_desc_sw128(a_lo + k * 2048, GR_BOX_BYTES, 1024);
          const uint64_t dbh = make_smem_desc_sw128(b_hi + k * 2048, GR_BOX_BYTES, 1024);
          const uint64_t dbl = make_smem_desc_sw128(b_lo + k * 2048, GR_BOX_BYTES, 1024);
          umma_f16(tmem_base, dal, dbh, idesc, (uint32_t)((it | k) != 0));
          umma_f16(tmem_base, dah, dbl, idesc, 1u);
          umma_f16(tmem_base, dah, dbh, idesc, 1u);
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* out = prm.partial + (int64_t)split * C * C;
    if (prm.rowmask) {
      // While the main loop runs these four warps apply the pixel mask: m in {0, 1}, so G = sum_p m_p F_p F_p^T is the
      // plain product with the masked pixels' rows of ONE operand zeroed.  A pixel is one 128-byte row of every
      // 64-channel box (the swizzle only permutes 16-byte chunks inside the row): thread t owns pixel t % 64 of the
      // stage in plane t / 64 (hi, lo) of B.  Generic-proxy stores -> fence.proxy.async -> masked_bar -> MMA.
      const int t = threadIdx.x - 64, pl = t & 63, plane = t >> 6;
      auto mask_at = [&](int it) -> float {
        const int64_t pix = pbeg + (int64_t)it * GR_KP + pl;
        return (it < num_iters && pix < prm.P) ? __ldg(prm.rowmask + pix) : 1.f;   // rows past P are TMA zero fill
      };
      float m = mask_at(0);
      for (int it = 0; it < num_iters; ++it) {
        const float m_next = mask_at(it + 1);            // in flight while this stage lands
        const int stage = it % Cfg::STAGES;
        const uint32_t phase = (uint32_t)(it / Cfg::STAGES) & 1u;
        mbar_wait(&full_bar[stage], phase, 14);
        // most stages have no masked pixel at all (masks are contiguous regions): no store, no proxy fence then
        if (__any_sync(0xffffffffu, m == 0.f)) {
          if (m == 0.f) {
            // (ALIAS: A and B are the same bytes - zeroing the pixel in both operands is still m_p, m in {0,1})
            uint8_t* rowp = smem + stage * Cfg::STAGE_BYTES + pl * 128 +
                            (ALIAS ? plane * Cfg::A_BYTES : 2 * Cfg::A_BYTES + plane * Cfg::B_BYTES);
#pragma unroll
            for (int g = 0; g < BN / 64; ++g)
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(rowp + g * GR_BOX_BYTES + j * 16) = make_uint4(0u, 0u, 0u, 0u);
          }
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&masked_bar[stage]);
        m = m_next;
      }
    }
    if (num_iters > 0) {
      mbar_wait(tmem_full_bar, 0, 13);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      if (num_iters > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (i0 + row < C) {
        float4* dst = reinterpret_cast<float4*>(out + (int64_t)(i0 + row) * C + j0 + c);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                               __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------
static void pick_patch(int H, int W, int& TH, int& TW) {
  // TH*TW == 128: minimise the padded area; on ties prefer the most square patch (best halo reuse in L2)
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

template <int BN, bool ALIAS>
static int launch_gram_tc_bn(const Act& fm, const float* rowmask, float* partial, int nsplit, cudaStream_t st) {
  using Cfg = GramCfg<BN, ALIAS>;
  GramTcParams prm;
  prm.C = fm.C;
  prm.nsplit = nsplit;
  prm.P = fm.pixels();
  prm.pix_per_split = ceil_div64(ceil_div64(std::max<int64_t>(prm.P, 1), GR_KP), nsplit) * GR_KP;   // whole 64-pixel stages
  prm.partial = partial;
  prm.rowmask = rowmask;
  CUtensorMap tm_hi, tm_lo;
  const uint64_t dims[2] = {(uint64_t)fm.C, (uint64_t)std::max<int64_t>(prm.P, 1)};
  const uint64_t strides[1] = {(uint64_t)fm.C * 2};
  const uint32_t box[2] = {64u, (uint32_t)GR_KP};
  int rc = make_tmap_bf16(&tm_hi, fm.hi, 2, dims, strides, box);
  if (rc) return rc;
  rc = make_tmap_bf16(&tm_lo, fm.lo, 2, dims, strides, box);
  if (rc) return rc;
  const int smem_bytes = Cfg::SMEM;
  static bool attr_set = false;
  if (!attr_set) {
    SMB_CUDA_CHECK(cudaFuncSetAttribute(gram_tc_kernel<BN, ALIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  dim3 grid(ceil_div(fm.C, TC_BM), fm.C / BN, nsplit);
  SMB_LAUNCH((gram_tc_kernel<BN, ALIAS>), grid, TC_THREADS, smem_bytes, st, tm_hi, tm_lo, prm);
  return SMB_OK;
}

// N tile of the Gram kernel.  Every CTA writes a 128 x BN fp32 partial tile and gram_mse reads all of them back, so
// the partial traffic is nsplit * C^2 * 4 bytes with nsplit ~ 2 * 148 / tiles: wide tiles on the small deep layers
// (C = 512, P = 4800) moved 39 MB of partials for a 9.8 MB feature map.  C >= 256 therefore uses BN = 64 (4x the
// tiles, a quarter of the splits and of the partial bytes; the feature map is L2 resident, the tensor pipe is not
// the limit at 2.5 GFLOP per layer).
int gram_tc_bn(int C) { return C == 128 ? 128 : 64; }

int launch_gram_tc(const Act& fm, const float* rowmask, float* partial, int nsplit, cudaStream_t st) {
  SMB_REQUIRE(fm.C % 64 == 0, "gram_tc: C=%d must be a multiple of 64", fm.C);
  SMB_REQUIRE(fm.pixels() > 0, "gram_tc: empty feature map");
  if (fm.C == 128) return launch_gram_tc_bn<128, true>(fm, rowmask, partial, nsplit, st);
  if (fm.C == 64) return launch_gram_tc_bn<64, true>(fm, rowmask, partial, nsplit, st);
  return launch_gram_tc_bn<64, false>(fm, rowmask, partial, nsplit, st);
}

}  // namespace smb
