// igemm_ph_kernel ("pair + halo") — fourth generation of the 3x3 implicit-GEMM conv.
//
// What the per-CTA timelines of igemm_tc2 showed (tools/gpu_trace_probe.py, conv3_2 at 480x640, 1.72 GHz):
//   * the MMA loop takes 77.5 k cycles for 56 k cycles of tensor work: an M=128 x N=128 kind::f16 MMA with both
//     operands in shared memory reads 8 KB per 64 cycles = the whole 128 B/clk of the SM's shared memory, and the
//     same memory takes the 83 B/clk TMA fill of a tap-per-stage pipeline (64 KB per 768 MMA-cycles);
//   * with the MMAs removed the TMA stream alone needs 70 k cycles (68 B/clk/SM from L2), with the TMA removed the
//     MMAs alone need 70 k: both sides sit at the same wall, which is why neither the halo kernel (less L2 traffic,
//     same shared-memory reads) nor the CTA-pair kernel (fewer reads, same traffic) moved the time on their own;
//   * the last tile's epilogue is exposed: 14 k cycles, dominated by 32-way divergent 16-byte global stores.
// This kernel combines the three fixes:
//   1. CTA pair (tcgen05 cta_group::2, M = 256): each SM reads its own 128 pixel rows of A and HALF of B per MMA
//      -> 6 KB per 64 cycles (96 B/clk) at N = 128;
//   2. activation halo: a 16 x 8-pixel patch loads its 18 x 10 halo once per 64-channel chunk and the nine taps are
//      shifted UMMA descriptors on it (start = base + dy * 1280 + dx * 128, SBO = 1280; verified on B200 in
//      igemm_halo_kernel) -> shared-memory fill per SM drops from 64 KB to 21 KB per tap (28 B/clk);
//   3. epilogue through shared memory: each thread writes its pixel's 64 channels as swizzled 16-byte chunks
//      (conflict free), one thread issues a TMA tensor store per plane; out-of-image pixels are clipped by TMA.
// Unchanged: bf16 hi/lo three-pass products into main/corr TMEM accumulators (double buffered), persistent CTAs
// with stream-K over (pair tile, K-chunk) units, first-K-part ownership with epoch-flag fix-ups, watchdogs.
#include <cstdlib>

#include "tc_common.cuh"
#include "smb_epilogue.cuh"
#include "smb_kernels.h"

namespace smb {
using namespace tc;

constexpr int I5_THREADS = 320;                       // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int I5_THREADS_MASK = 384;                  // + warps 10-11: pixel mask of a fused 1x1 term (launched only then)
constexpr int I5_EPI_THREADS = 256;
constexpr int I5_BM = 128;                            // pixel rows per CTA (the pair covers 256)
constexpr int I5_TH = 16, I5_TW = 8;                  // output patch of one CTA: 16 rows x 8 pixels
constexpr int I5_HR = I5_TH + 2, I5_HW = I5_TW + 2;   // halo: 18 rows x 10 pixels
constexpr int I5_PITCH = I5_HW * 128;                 // bytes between halo rows (dense)
constexpr int I5_A_PLANE = (I5_HR * I5_PITCH + 1023) & ~1023;   // 23552
constexpr int I5_A_BUF = 2 * I5_A_PLANE;              // hi + lo
constexpr int I5_A_TX = 2 * I5_HR * I5_HW * 128;      // bytes TMA writes per halo and CTA (hi + lo) = 46080
constexpr int I5_OUT_PLANE = I5_BM * 128;             // 128 pixels x 64 channels bf16 = 16 KiB
constexpr int I5_OUT_BYTES = 2 * I5_OUT_PLANE;        // hi + lo staging of one 64-channel group
constexpr int I5_BAR_BYTES = 512;
constexpr int I5_SMEM_LIMIT = 227 * 1024;
constexpr int I5_MAX_NB = 9;                          // 9 stages hold all taps of a 64-wide layer (resident B)
constexpr uint32_t I5_PEER_MASK = 0xFEFFFFFFu;        // shared::cluster address -> same offset in CTA 0 of the pair

template <int BN>
struct I5Cfg {
  static constexpr int B_PLANE = (BN / 2) * 128;                 // this CTA's half of the B tile of one tap
  static constexpr int B_STAGE = 2 * B_PLANE;                    // hi + lo
  static constexpr int NB_FIT = (I5_SMEM_LIMIT - 1024 - 2 * I5_A_BUF - I5_OUT_BYTES - I5_BAR_BYTES) / B_STAGE;
  // TMEM columns of one tile buffer.  BN = 128 / 16: main + corr.  BN = 64: [main | corr1] interleaved per CTA half
  // (128 columns, written by ONE N = 128 MMA whose B operand is the hi plane followed by the lo plane) + corr2 (64).
  // (WIDE64 trades the third MMA of a K step for 64 more columns; measured 71 -> 68 us on conv1_2, so the 64-channel
  // layers are not bound by shared-memory reads alone.  The default keeps main + corr = 128 columns and uses the
  // freed TMEM for FOUR tile buffers: with 9 taps of N = 64 per tile (1.8 us of MMAs) the commit -> epilogue ->
  // remote-arrive round trip of a buffer is longer than the MMAs of the one other tile a double buffer can hide.)
#ifndef SMB_PH_WIDE64
#define SMB_PH_WIDE64 0
#endif
  static constexpr bool WIDE64 = SMB_PH_WIDE64 != 0;   // compile flag, see DESIGN §4
  static constexpr int TBUF = (BN == 64 && WIDE64) ? 192 : 2 * BN;
  static constexpr int NT = (BN == 64 && !WIDE64) ? 4 : 2;              // tile buffers in TMEM
  static constexpr int TMEM_NEED = NT * TBUF;
  static constexpr int NB = NB_FIT < I5_MAX_NB ? NB_FIT : I5_MAX_NB;
  static constexpr int SMEM = 1024 + 2 * I5_A_BUF + I5_OUT_BYTES + NB * B_STAGE + I5_BAR_BYTES;
  static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 256 ? 256 : 512);
  static_assert(NB >= 4, "B ring too shallow");
};

struct IGemm5Params {
  int H, W, tiles_x, tiles_m, tiles_n, kchunks, N;
  const float* fmask;         // [H*W] {0,1} pixel mask of the fused term (rows of F with mask 0 do not contribute) or nullptr
  int kreg;                   // K-chunks kc < kreg are the 3x3 conv (A = input halo, 9 taps of B); chunks kc >= kreg are a
                              // fused 1x1 term D += F * G^T (Gram backward of the layer receiving the gradient): A = halo of
                              // F (tmF), only the centre tap multiplies, B = G (tmG).  kreg == kchunks: plain conv.
  long long total_units;      // pair_tiles * kchunks * 9 work units (one tap of one K-chunk of one pair tile)
  int align;                  // stream-K range boundaries are multiples of this many units: 9 (whole halo units) or 1
  float* ws;                  // [grid][128][BN] fp32 partial tiles (indexed by CTA id)
  unsigned int* flags;        // [grid]
  unsigned int epoch;
  int resident;               // 1: one K-chunk, one N tile -> the 9 B tiles are loaded once and stay in stages 0..8
  int knob;                   // experiment bits (SMB_PH_KNOB): 1 = request the next halo as early as possible,
                              // 2 = epilogue drains TMEM but computes / stores nothing, 4 = no MMAs, 8 = no TMA loads
  int halo_split;             // 1: the activation halo is requested as three 6-row boxes per plane (SMB_PH_HALO_SPLIT)
  int tma_out;                // 1: hi/lo planes leave through shared memory + TMA tensor stores, 2: fp32 rows do,
                              // 3: only maxpool2x2 of the hi/lo planes is stored (inference-only forward into a pool)
  unsigned long long* trace;  // optional [grid][16] per-CTA timeline (same slots as igemm_tc2), nullptr = off
  Epilogue ep;
};

enum : int { T5_GT_IN = 0, T5_GT_OUT, T5_CLK_IN, T5_CLK_PROLOGUE, T5_CLK_TMA_END, T5_CLK_MMA_FIRST, T5_CLK_MMA_END,
             T5_CLK_EPI_FIRST, T5_CLK_EPI_END, T5_W_FLAGS, T5_W_TMEM_FULL, T5_W_FULL, T5_W_TMEM_EMPTY, T5_W_EMPTY,
             T5_CLK_OUT, T5_SMID };
__device__ __forceinline__ unsigned long long i5_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- cta_group::2 / cluster flavours of the building blocks -------------------------------------------------
__device__ __forceinline__ uint32_t i5_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void i5_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// executed by both CTAs; the transaction bytes are credited to the barrier of CTA 0 of the pair
__device__ __forceinline__ void i5_tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & I5_PEER_MASK), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
// same load, completion credited to a barrier of the executing CTA
__device__ __forceinline__ void i5_tma_load_3d_local(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void i5_tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void i5_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void i5_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void i5_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void i5_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void i5_tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void i5_tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void i5_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void i5_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same MMA with the two shared-memory descriptors given as 32-bit halves (the high halves are compile-time constants,
// the low halves differ by small immediates between the 12 MMAs of a tap)
__device__ __forceinline__ void i5_umma2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// walks halo units (pair tile, K-chunk) in stream-K order without divisions: unit = (pm * tiles_n + n_tile) * ipt + kc
struct UnitWalk {
  int kc, n_tile, pm;
  __device__ __forceinline__ void init(int unit, int ipt, int tiles_n) {
    kc = unit % ipt;
    const int tile = unit / ipt;
    n_tile = tile % tiles_n;
    pm = tile / tiles_n;
  }
  __device__ __forceinline__ void next(int ipt, int tiles_n) {
    if (++kc == ipt) {
      kc = 0;
      if (++n_tile == tiles_n) {
        n_tile = 0;
        ++pm;
      }
    }
  }
};
// arrive (after all previously issued MMAs completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void i5_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void i5_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// cluster-scope release / acquire pair for data written by the generic proxy of either CTA (mask warps -> MMA issuer)
__device__ __forceinline__ void i5_arrive_cta_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void i5_wait_cluster(uint64_t* bar, uint32_t parity, int who) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("[smb] igemm_ph watchdog: wait #%d timed out in block %d thread %d\n", who, (int)blockIdx.x,
             (int)threadIdx.x);
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ unsigned int i5_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void i5_st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// K-major SWIZZLE_128B operand whose 8-row groups are `sbo` bytes apart (1024 for dense tiles, the halo row pitch
// for the activation operand); base_offset 0: the tensor core swizzles on absolute shared-memory address bits
__device__ __forceinline__ uint64_t i5_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;                                   // LBO (unused for swizzled K-major) = 16 B
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void i5_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I5_THREADS_MASK, 1)
igemm_ph_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
                const __grid_constant__ CUtensorMap tmF_hi, const __grid_constant__ CUtensorMap tmF_lo,
                const __grid_constant__ CUtensorMap tmG_hi, const __grid_constant__ CUtensorMap tmG_lo,
                const __grid_constant__ CUtensorMap tmA6_hi, const __grid_constant__ CUtensorMap tmA6_lo,
                const IGemm5Params prm) {
  using Cfg = I5Cfg<BN>;
  constexpr int NB = Cfg::NB;
  const uint32_t rank = i5_ctarank();                 // 0 = leader (issues the MMAs), 1 = peer
  const long long G = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const long long cta = blockIdx.x;
  // stream-K range of pair c: [bound(c), bound(c + 1)), boundaries aligned to prm.align units; G <= total / align
  const long long n_al = prm.total_units / prm.align;
  auto bound = [&](long long c) { return (c * n_al / G) * prm.align; };
  const long long u0 = bound(pair), u1 = bound(pair + 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                         // [2 bufs][hi, lo][18 halo rows x 1280 B]
  uint8_t* sOut = sA + 2 * I5_A_BUF;                          // [hi, lo][128 pixels x 128 B]
  uint8_t* sB = sOut + I5_OUT_BYTES;                          // [NB][hi, lo][BN/2 x 128 B]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + NB * Cfg::B_STAGE);   // [2]  used in the leader
  uint64_t* a_empty = a_full + 2;                             // [2]  both CTAs
  uint64_t* b_full = a_empty + 2;                             // [NB] used in the leader
  uint64_t* b_empty = b_full + I5_MAX_NB;                     // [NB] both CTAs
  uint64_t* tmem_full_bar = b_empty + I5_MAX_NB;              // [4]  both CTAs
  uint64_t* tmem_empty_bar = tmem_full_bar + 4;               // [4]  used in the leader
  uint64_t* f_land = tmem_empty_bar + 4;                      // [2]  both CTAs: own halo of a MASKED fused chunk has landed
  uint64_t* a_masked = f_land + 2;                            // [2]  used in the leader: both halos landed and masked
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_masked + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ipt = prm.kchunks;                                // (K-chunk) halo loads per pair tile
  const int tpt = 9 * ipt;                                    // work units (one tap of one K-chunk) per pair tile
  unsigned long long* tr = prm.trace ? prm.trace + (size_t)cta * 16 : nullptr;
  if (tr && threadIdx.x == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr[T5_GT_IN] = i5_global_ns();
    tr[T5_CLK_IN] = (unsigned long long)clock64();
    tr[T5_SMID] = smid;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    if (prm.halo_split) {
      tma_prefetch_desc(&tmA6_hi);
      tma_prefetch_desc(&tmA6_lo);
    }
    if (prm.tma_out) {
      tma_prefetch_desc(&tmO_hi);
      tma_prefetch_desc(&tmO_lo);
    }
    if (prm.kreg < prm.kchunks) {
      tma_prefetch_desc(&tmF_hi);
      tma_prefetch_desc(&tmF_lo);
      tma_prefetch_desc(&tmG_hi);
      tma_prefetch_desc(&tmG_lo);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&f_land[s], 1);
      mbar_init(&a_masked[s], 4);            // 2 mask warps of each CTA
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 16);     // 8 epilogue warps of each CTA
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    i5_tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    i5_tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  i5_cluster_sync();                         // peer barriers are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();                                // everything above overlaps the tail of the preceding launch
  if (tr && threadIdx.x == 0) tr[T5_CLK_PROLOGUE] = (unsigned long long)clock64();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own halo, own half of B) =====================
    // Program order: halo of the first unit; then one B tile per work unit (tap), with the NEXT unit's halo issued
    // at tap 3 of the current one (it only needs the other A buffer to be drained).  A range may start / end in the
    // middle of a halo unit.  All indices are walked incrementally: this single thread has ~768 cycles per tap.
    if (elect_one()) {
      const int w0 = (int)u0, w1 = (int)u1;
      UnitWalk cur, nxt;
      cur.init(w0 / 9, ipt, prm.tiles_n);
      nxt = cur;
      int tap = w0 % 9, u = w0 / 9, a_next = u;
      const int a_last = (w1 - 1) / 9;
      int ga = 0, bs = 0;
      uint32_t bpar = 1;                       // parity of a never-completed phase: the first NB waits pass at once
      long long w_empty = 0;
      // request the halo of unit `nxt` into the next A buffer; `force` = wait for the buffer, else give up if the
      // MMAs of the unit that used it two units ago have not completed yet (the caller retries at the next tap)
      auto issue_A = [&](bool force) -> bool {
        const int abuf = ga & 1;
        const uint32_t par = (uint32_t)((ga >> 1) & 1) ^ 1u;
        if (!mbar_try_wait(&a_empty[abuf], par)) {
          if (!force) return false;
          mbar_wait(&a_empty[abuf], par, 61);
        }
        ++ga;
        if (prm.knob & 8) {
          if (rank == 0) mbar_arrive(&a_full[abuf]);
        } else {
          const int m_tile = 2 * nxt.pm + (int)rank;
          const int y0 = (m_tile / prm.tiles_x) * I5_TH, x0 = (m_tile % prm.tiles_x) * I5_TW;
          const bool to_mask = prm.fmask && nxt.kc >= prm.kreg;
          if (to_mask) mbar_arrive_expect_tx(&f_land[abuf], I5_A_TX);            // own bytes, own barrier
          else if (rank == 0) mbar_arrive_expect_tx(&a_full[abuf], 2 * I5_A_TX);  // bytes of both CTAs
          uint8_t* ah = sA + abuf * I5_A_BUF;
          if (nxt.kc < prm.kreg && prm.halo_split) {
            // the 18-row halo as three 6-row boxes per plane: a box is fetched as 128-byte rows with a bounded number
            // of requests in flight, so one 180-row box takes ~8 k cycles from request to complete_tx even out of L2 -
            // longer than the 6.9 k cycles of MMAs (N = 128) that a double-buffered halo can hide
#pragma unroll
            for (int r3 = 0; r3 < 3; ++r3) {
              i5_tma_load_3d(ah + r3 * 6 * I5_PITCH, &tmA6_hi, &a_full[abuf], nxt.kc * 64, x0 - 1, y0 - 1 + 6 * r3);
              i5_tma_load_3d(ah + I5_A_PLANE + r3 * 6 * I5_PITCH, &tmA6_lo, &a_full[abuf], nxt.kc * 64, x0 - 1,
                             y0 - 1 + 6 * r3);
            }
          } else if (nxt.kc < prm.kreg) {
            i5_tma_load_3d(ah, &tmA_hi, &a_full[abuf], nxt.kc * 64, x0 - 1, y0 - 1);
            i5_tma_load_3d(ah + I5_A_PLANE, &tmA_lo, &a_full[abuf], nxt.kc * 64, x0 - 1, y0 - 1);
          } else if (!prm.fmask) {               // fused 1x1 term: same halo geometry on the feature tensor
            i5_tma_load_3d(ah, &tmF_hi, &a_full[abuf], (nxt.kc - prm.kreg) * 64, x0 - 1, y0 - 1);
            i5_tma_load_3d(ah + I5_A_PLANE, &tmF_lo, &a_full[abuf], (nxt.kc - prm.kreg) * 64, x0 - 1, y0 - 1);
          } else {                               // ... masked: lands on this CTA's own barrier, the mask warps pass it on
            i5_tma_load_3d_local(ah, &tmF_hi, &f_land[abuf], (nxt.kc - prm.kreg) * 64, x0 - 1, y0 - 1);
            i5_tma_load_3d_local(ah + I5_A_PLANE, &tmF_lo, &f_land[abuf], (nxt.kc - prm.kreg) * 64, x0 - 1, y0 - 1);
          }
        }
        nxt.next(ipt, prm.tiles_n);
        ++a_next;
        return true;
      };
      issue_A(true);
      if (prm.resident) {
        // every tile of this layer multiplies the same nine B tiles (all CTAs would otherwise stream the same 144 KB
        // from a handful of L2 lines): load them once, then only halos move
        for (int t9 = 0; t9 < 9; ++t9) {
          if (prm.knob & 8) {
            if (rank == 0) mbar_arrive(&b_full[t9]);
            continue;
          }
          if (rank == 0) mbar_arrive_expect_tx(&b_full[t9], 2 * Cfg::B_STAGE);
          uint8_t* bh = sB + t9 * Cfg::B_STAGE;
          const int nb0 = (int)rank * (BN / 2);
          i5_tma_load_3d(bh, &tmB_hi, &b_full[t9], 0, nb0, t9);
          i5_tma_load_3d(bh + Cfg::B_PLANE, &tmB_lo, &b_full[t9], 0, nb0, t9);
        }
        while (a_next <= a_last) issue_A(true);
      }
      for (int w = prm.resident ? w1 : w0; w < w1; ++w) {
        if (a_next == u + 1 && a_next <= a_last) {
          if (prm.knob & 1) issue_A(tap == 8);
          else if (tap >= 3) issue_A(true);
        }
        const bool fused = cur.kc >= prm.kreg;
        if (!fused || tap == 4) {                // a fused chunk has one B tile (centre tap), the other taps are no-ops
          const long long tw0 = tr ? clock64() : 0;
          mbar_wait(&b_empty[bs], bpar, 62);
          if (tr) w_empty += clock64() - tw0;
          if (prm.knob & 8) {
            if (rank == 0) mbar_arrive(&b_full[bs]);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&b_full[bs], 2 * Cfg::B_STAGE);
            uint8_t* bh = sB + bs * Cfg::B_STAGE;
            const int nb0 = cur.n_tile * BN + (int)rank * (BN / 2);
            if (!fused) {
              i5_tma_load_3d(bh, &tmB_hi, &b_full[bs], cur.kc * 64, nb0, tap);
              i5_tma_load_3d(bh + Cfg::B_PLANE, &tmB_lo, &b_full[bs], cur.kc * 64, nb0, tap);
            } else {
              i5_tma_load_3d(bh, &tmG_hi, &b_full[bs], (cur.kc - prm.kreg) * 64, nb0, 0);
              i5_tma_load_3d(bh + Cfg::B_PLANE, &tmG_lo, &b_full[bs], (cur.kc - prm.kreg) * 64, nb0, 0);
            }
          }
          if (++bs == NB) {
            bs = 0;
            bpar ^= 1u;
          }
        }
        if (++tap == 9) {
          tap = 0;
          ++u;
          cur.next(ipt, prm.tiles_n);
        }
      }
      if (tr) {
        tr[T5_CLK_TMA_END] = (unsigned long long)clock64();
        tr[T5_W_EMPTY] = (unsigned long long)w_empty;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // One thread issues 12 MMAs per tap (768 tensor cycles at N = 128): the scalar work per tap has to stay well
    // below that, so descriptors are built from precomputed 32-bit halves (low word = address >> 4 | LBO, one add per
    // operand) and every index is a wrapped counter.
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * I5_BM, BN, 0, 0);
      constexpr uint32_t HI_A = (uint32_t)(I5_PITCH >> 4) | (1u << 14) | (2u << 29);   // SBO, version 1, SWIZZLE_128B
      constexpr uint32_t HI_B = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t A_LO_PLANE = I5_A_PLANE >> 4, B_LO_PLANE = Cfg::B_PLANE >> 4;
      const uint32_t a_word0 = (smem_u32(sA) >> 4) | (1u << 16);                      // low descriptor word of A buffer 0 (hi plane)
      const uint32_t b_word0 = (smem_u32(sB) >> 4) | (1u << 16);                      // ... of B stage 0 (hi plane)
      const int w1 = (int)u1;
      int w = (int)u0;
      int tap = w % 9, ga = 0, bs = 0, seg = 0;
      uint32_t bfpar = 0, a_word = 0;
      uint32_t par_full = 0, par_masked = 0;   // bit b = parity of the next phase of a_full[b] / a_masked[b]
      int abuf = 0;
      long long w_full = 0, w_full_a = 0, w_tempty = 0;
      bool first = true;
      while (w < w1) {
        const int ks = w % tpt;
        const int ke = (w1 - w < tpt - ks) ? ks + (w1 - w) : tpt;
        const int buf = seg % Cfg::NT;
        const uint32_t use = (uint32_t)(seg / Cfg::NT);
        const long long tw1 = tr ? clock64() : 0;
        mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u, 63);       // both epilogues drained this buffer
        if (tr) w_tempty += clock64() - tw1;
        tc_fence_after();
        const uint32_t t_main = tmem_base + (uint32_t)(buf * Cfg::TBUF);
        const uint32_t t_corr = t_main + (uint32_t)((BN == 64 && Cfg::WIDE64) ? 128 : BN);
        int kc = ks / 9;                         // K-chunk of unit t (t = kc * 9 + tap)
        uint32_t started = 0;                    // 0 until the first MMA of this segment has initialised the accumulators
#pragma unroll 1
        for (int t = ks; t < ke; ++t) {
          if (t == ks || tap == 0) {             // first tap of a halo unit inside this range: its halo must have landed
            abuf = ga & 1;
            const long long tw2 = tr ? clock64() : 0;
            if (prm.fmask && kc >= prm.kreg) {   // masked fused chunk: both CTAs' mask warps have passed the halo on
              i5_wait_cluster(&a_masked[abuf], (par_masked >> abuf) & 1u, 67);
              par_masked ^= 1u << abuf;
            } else {
              mbar_wait(&a_full[abuf], (par_full >> abuf) & 1u, 64);
              par_full ^= 1u << abuf;
            }
            if (tr) w_full_a += clock64() - tw2;
            a_word = a_word0 + (uint32_t)abuf * (uint32_t)(I5_A_BUF >> 4);
          }
          if (kc < prm.kreg || tap == 4) {       // (a fused 1x1 chunk multiplies its centre tap only)
            if (prm.resident) {                  // stage = tap, loaded once (phase 0 stays complete)
              bs = tap;
              bfpar = 0;
            }
            const long long tw3 = tr ? clock64() : 0;
            mbar_wait(&b_full[bs], bfpar, 65);
            if (tr) {
              const long long now = clock64();
              w_full += now - tw3;
              if (first) tr[T5_CLK_MMA_FIRST] = (unsigned long long)now;
              first = false;
            }
            tc_fence_after();
            const int dy = (tap >= 6) ? 2 : (tap >= 3 ? 1 : 0);
            const uint32_t a_t = a_word + (uint32_t)(dy * (I5_PITCH >> 4) + (tap - 3 * dy) * 8);   // halo coords of the tap
            const uint32_t b_t = b_word0 + (uint32_t)bs * (uint32_t)(Cfg::B_STAGE >> 4);
            if (!(prm.knob & 4)) {
              if constexpr (BN == 64 && Cfg::WIDE64) {
                // An M = 256, N = 64 MMA reads 5 KB of shared memory per SM in 32 tensor cycles (160 B/clk against the
                // 128 B/clk the SM has): three of them per K step were shared-memory bound (57 cycles each, measured).
                // The two products that share the hi plane of A become ONE N = 128 MMA - B = this CTA's 32 hi rows
                // followed by its 32 lo rows, which are adjacent in the stage - so A_hi is read once: 11 KB per K step
                // for 96 tensor cycles.  Columns: [0,32) hi.hi and [32,64) hi.lo of the leader's channels, [64,128) the
                // same of the peer's; lo.hi goes to its own 64 columns.
                constexpr uint32_t idesc_wide = make_idesc_bf16(2 * I5_BM, 128, 0, 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t acc = started | (uint32_t)(k > 0);
                  i5_umma2(t_corr, a_t + A_LO_PLANE + 2 * k, HI_A, b_t + 2 * k, HI_B, idesc, acc);
                  i5_umma2(t_main, a_t + 2 * k, HI_A, b_t + 2 * k, HI_B, idesc_wide, acc);
                }
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t acc = started | (uint32_t)(k > 0);
                  i5_umma2(t_corr, a_t + A_LO_PLANE + 2 * k, HI_A, b_t + 2 * k, HI_B, idesc, acc);
                  i5_umma2(t_corr, a_t + 2 * k, HI_A, b_t + B_LO_PLANE + 2 * k, HI_B, idesc, 1u);
                  i5_umma2(t_main, a_t + 2 * k, HI_A, b_t + 2 * k, HI_B, idesc, acc);
                }
              }
            }
            started = 1u;
            if (!prm.resident) {
              i5_commit_mc(&b_empty[bs]);        // frees this B stage in BOTH CTAs
              if (++bs == NB) {
                bs = 0;
                bfpar ^= 1u;
              }
            }
          }
          if (tap == 8 || t == ke - 1) {         // last tap of this halo unit inside the range
            i5_commit_mc(&a_empty[abuf]);        // frees this halo buffer in BOTH CTAs
            ++ga;
          }
          if (++tap == 9) {
            tap = 0;
            ++kc;
          }
        }
        i5_commit_mc(&tmem_full_bar[buf]);       // accumulators complete, both CTAs
        w += (ke - ks);
        ++seg;
      }
      if (tr) {
        tr[T5_CLK_MMA_END] = (unsigned long long)clock64();
        // low 32 bits: cycles waiting for B stages, high 32 bits: cycles waiting for halos
        tr[T5_W_FULL] = ((unsigned long long)w_full_a << 32) | ((unsigned long long)w_full & 0xffffffffull);
        tr[T5_W_TMEM_EMPTY] = (unsigned long long)w_tempty;
      }
    }
  } else if (warp < 10) {
    // ===================== epilogue warps (each CTA drains its own 128 TMEM lanes) =====================
    // warps 2-5 and 6-9 both cover the four TMEM lane quarters (quarter = warp % 4); set 0 takes the first 32
    // channels of every 64-channel group, set 1 the second 32, and they meet at the staging buffer / named barrier
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int cset = (warp - 2) >> 2;
    const bool epi_leader = (warp == 2 && lane == 0);
    float* my_slot = prm.ws + (size_t)cta * I5_BM * BN;
    const uint32_t s_out = smem_u32(sOut);
    long long w_flags = 0, w_tfull = 0;
    const bool tr_me = tr && epi_leader;
    int seg = 0;
    for (long long u = u0; u < u1; ++seg) {
      const int tile = (int)(u / tpt);
      const int ks = (int)(u % tpt);
      const long long left = u1 - u;
      const int ke = (left < (long long)(tpt - ks)) ? ks + (int)left : tpt;
      const bool owner = (ks == 0);
      const int buf = seg % Cfg::NT;
      const uint32_t use = (uint32_t)(seg / Cfg::NT);
      const int n_tile = tile % prm.tiles_n;
      const int m_tile = 2 * (tile / prm.tiles_n) + (int)rank;
      const int y0 = (m_tile / prm.tiles_x) * I5_TH, x0 = (m_tile % prm.tiles_x) * I5_TW;
      const int yy = y0 + row / I5_TW, xx = x0 + row % I5_TW;
      const bool valid = (m_tile < prm.tiles_m) && (yy < prm.H) && (xx < prm.W);
      const int64_t p = (int64_t)yy * prm.W + xx;
      const int n0 = n_tile * BN;

      int npeer = 0;
      const long long tw4 = tr_me ? clock64() : 0;
      if (owner && ke < tpt) {
        const long long tile_end = (long long)(tile + 1) * tpt;
        long long c = pair + 1;
        while (c < G && bound(c) < tile_end) {
          const unsigned int* f = prm.flags + 2 * c + rank;           // same-rank CTA of the later pair
          if (lane == 0) {
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile const unsigned int*>(f) != prm.epoch) {
              __nanosleep(64);
              if (clock64() - t0 > 4000000000LL) {
                printf("[smb] igemm_ph stream-K watchdog: CTA %d waiting for partial of CTA %d\n", (int)cta,
                       (int)(2 * c + rank));
                asm volatile("trap;");
              }
            }
          }
          __syncwarp();
          (void)i5_ld_acquire(f);
          ++npeer;
          ++c;
        }
      }

      const long long tw5 = tr_me ? clock64() : 0;
      mbar_wait(&tmem_full_bar[buf], use & 1u, 66);
      if (tr_me) {
        w_flags += tw5 - tw4;
        w_tfull += clock64() - tw5;
      }
      tc_fence_after();
      const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Cfg::TBUF);
      const uint32_t t_corr = t_main + (uint32_t)((BN == 64 && Cfg::WIDE64) ? 128 : BN);
      const bool staged = owner && prm.tma_out;
      if constexpr (BN == 16) {
        // narrow tile (data gradient of the 3-channel first layer, N padded 3 -> 16): main and corr are adjacent
        // 16-column blocks, one 32-column TMEM load fetches both; only warp set 0 computes, set 1 just keeps the
        // barrier counts
        float v[16];
        if (cset == 0) {
          uint32_t rm[32];
          tmem_ld_32x32(t_main, rm);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rm[16 + j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) i5_arrive_cta(&tmem_empty_bar[buf], 0);
        if (cset == 0) {
          for (int k = 1; k <= npeer; ++k) {
            const float4* src = reinterpret_cast<const float4*>(prm.ws + (size_t)(cta + 2 * k) * I5_BM * BN) + row;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 a = __ldcg(src + (size_t)j * I5_BM);
              v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
            }
          }
          if (!owner) {
            float4* dst = reinterpret_cast<float4*>(my_slot) + row;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[(size_t)j * I5_BM] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else if (valid) {
            if (prm.ep.out_planar3) {            // (3, H, W) fp32 image gradient
              const int64_t P = (int64_t)prm.H * prm.W;
              prm.ep.out_planar3[p] = v[0];
              prm.ep.out_planar3[P + p] = v[1];
              prm.ep.out_planar3[2 * P + p] = v[2];
            } else {
              epilogue_store<16>(prm.ep, p, n0, prm.N, v);
            }
          }
        }
      } else {
#pragma unroll 1
      for (int c = cset * 32; c < BN; c += 64) {
        if (staged) {
          // the staging buffer is free once the previous group's tensor stores have read it
          if (epi_leader) i5_bulk_wait_read0();
          i5_epi_bar();
        }
        // Partial tiles travel through the workspace in a warp-coalesced layout: float4 number (c/4 + j) * 128 + row
        // holds channels c+4j .. c+4j+3 of pixel `row` (a warp reads / writes 512 contiguous bytes per access).
        // The first peer's partial is requested before the TMEM load so that the L2 round trip overlaps it.
        float4 pv[8];
        if (npeer > 0) {
          const float4* src = reinterpret_cast<const float4*>(prm.ws + (size_t)(cta + 2) * I5_BM * BN) +
                              (size_t)(c >> 2) * I5_BM + row;
#pragma unroll
          for (int j = 0; j < 8; ++j) pv[j] = __ldcg(src + (size_t)j * I5_BM);
        }
        uint32_t rm[32], rc[32];
        float v[32];
        if constexpr (BN == 64 && Cfg::WIDE64) {  // main / corr1 of channels c..c+31 sit at columns 2c / 2c + 32, corr2 at c
          tmem_ld_32x32(t_main + (uint32_t)(2 * c), rm);
          tmem_ld_32x32(t_main + (uint32_t)(2 * c + 32), rc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]);
          tmem_ld_32x32(t_corr + (uint32_t)c, rm);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(rm[j]);
        } else {
          tmem_ld_32x32(t_main + (uint32_t)c, rm);
          tmem_ld_32x32(t_corr + (uint32_t)c, rc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]);
        }
        if (c + 64 >= BN) {                      // this warp's last TMEM read of the buffer: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) i5_arrive_cta(&tmem_empty_bar[buf], 0);     // the leader's barrier collects 16 arrivals
        }
        if ((prm.knob & 2) && owner) continue;
        if (npeer > 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 * j] += pv[j].x; v[4 * j + 1] += pv[j].y; v[4 * j + 2] += pv[j].z; v[4 * j + 3] += pv[j].w;
          }
        }
        for (int k = 2; k <= npeer; ++k) {       // fixed order => deterministic sums
          const float4* src = reinterpret_cast<const float4*>(prm.ws + (size_t)(cta + 2 * k) * I5_BM * BN) +
                              (size_t)(c >> 2) * I5_BM + row;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = __ldcg(src + (size_t)j * I5_BM);
            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
          }
        }
        if (!owner) {
          float4* dst = reinterpret_cast<float4*>(my_slot) + (size_t)(c >> 2) * I5_BM + row;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[(size_t)j * I5_BM] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else if (!staged) {
          if (valid) epilogue_store<32>(prm.ep, p, n0 + c, prm.N, v);
        } else {
          epilogue_apply<32>(prm.ep, p, n0 + c, prm.N, v, valid);
          if (prm.tma_out == 3) {
            // MaxPool2d(2,2) in registers: pixel (py, px) of the 16 x 8 patch is lane (py % 4) * 8 + px of quarter
            // py / 4, so the 2x2 window partners are lane ^ 1 and lane ^ 8.  Lanes with both bits clear hold the
            // window maximum = pooled pixel (py / 2, px / 2) of the 8 x 4 pooled patch; windows that stick out of the
            // image lie outside the pooled tensor (floor mode) and are clipped by the tensor store.
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
              v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 8));
            }
            if ((lane & 9) == 0) {
              const uint32_t prow = (uint32_t)((q * 2 + (lane >> 4)) * 4 + ((lane & 7) >> 1));
              const uint32_t pbase = s_out + prow * 128u, psw = prow & 7u, pch0 = (uint32_t)((c & 63) >> 3);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 h, l;
                split2_pack(v[8 * j], v[8 * j + 1], h.x, l.x);
                split2_pack(v[8 * j + 2], v[8 * j + 3], h.y, l.y);
                split2_pack(v[8 * j + 4], v[8 * j + 5], h.z, l.z);
                split2_pack(v[8 * j + 6], v[8 * j + 7], h.w, l.w);
                const uint32_t a = pbase + (((pch0 + (uint32_t)j) ^ psw) << 4);
                i5_sts128(a, h);
                i5_sts128(a + I5_OUT_PLANE, l);
              }
            }
            fence_proxy_async_smem();
            i5_epi_bar();
            if (epi_leader) {
              i5_tma_store_3d(&tmO_hi, sOut, n0 + (c & ~63), x0 >> 1, y0 >> 1);
              i5_tma_store_3d(&tmO_lo, sOut + I5_OUT_PLANE, n0 + (c & ~63), x0 >> 1, y0 >> 1);
              i5_bulk_commit();
            }
            continue;
          }
          // pixel `row` of the patch is 128-byte row `row` of the staging tile; 16-byte chunk index XOR (row & 7)
          // = the SWIZZLE_128B pattern the tensor store expects (conflict free: 8 lanes cover 8 distinct chunks)
          const uint32_t rbase = s_out + (uint32_t)row * 128u;
          const uint32_t sw = (uint32_t)(row & 7);
          if (prm.tma_out == 2) {
            // fp32 rows: this warp set's 32 channels are one 128-byte row of ITS half of the staging buffer
            const uint32_t hb = rbase + (uint32_t)cset * I5_OUT_PLANE;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 w4;
              w4.x = __float_as_uint(v[4 * j]); w4.y = __float_as_uint(v[4 * j + 1]);
              w4.z = __float_as_uint(v[4 * j + 2]); w4.w = __float_as_uint(v[4 * j + 3]);
              i5_sts128(hb + ((((uint32_t)j) ^ sw) << 4), w4);
            }
            fence_proxy_async_smem();
            i5_epi_bar();
            if (epi_leader) {
              i5_tma_store_3d(&tmO_hi, sOut, n0 + (c & ~63), x0, y0);
              i5_tma_store_3d(&tmO_hi, sOut + I5_OUT_PLANE, n0 + (c & ~63) + 32, x0, y0);
              i5_bulk_commit();
            }
            continue;
          }
          const uint32_t ch0 = (uint32_t)((c & 63) >> 3);
          // side output for a layer that feeds MaxPool2d(2,2): the 2x2 window partners of patch pixel (py, px) =
          // lane (py % 4) * 8 + px of quarter py / 4 are lanes ^1 and ^8; lanes with both bits clear own the window and
          // store its maximum straight to the pooled planes (a quarter of the pixels, 64 contiguous bytes per plane)
          const bool pool_side = prm.ep.pool_hi != nullptr;
          const int pY = yy >> 1, pX = xx >> 1;
          const bool pool_store = pool_side && (lane & 9) == 0 && pY < (prm.H >> 1) && pX < (prm.W >> 1);
          const int64_t poff = ((int64_t)pY * (prm.W >> 1) + pX) * prm.N + n0 + c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 h, l;
            split2_pack(v[8 * j], v[8 * j + 1], h.x, l.x);
            split2_pack(v[8 * j + 2], v[8 * j + 3], h.y, l.y);
            split2_pack(v[8 * j + 4], v[8 * j + 5], h.z, l.z);
            split2_pack(v[8 * j + 6], v[8 * j + 7], h.w, l.w);
            const uint32_t a = rbase + (((ch0 + (uint32_t)j) ^ sw) << 4);
            i5_sts128(a, h);
            i5_sts128(a + I5_OUT_PLANE, l);
            if (pool_side) {                     // (warp-uniform branch: the shuffles need all 32 lanes)
              float m[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float t = v[8 * j + e];
                t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 1));
                m[e] = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 8));
              }
              if (pool_store) {
                uint4 ph, pl;
                split2_pack(m[0], m[1], ph.x, pl.x);
                split2_pack(m[2], m[3], ph.y, pl.y);
                split2_pack(m[4], m[5], ph.z, pl.z);
                split2_pack(m[6], m[7], ph.w, pl.w);
                *reinterpret_cast<uint4*>(prm.ep.pool_hi + poff + 8 * j) = ph;
                *reinterpret_cast<uint4*>(prm.ep.pool_lo + poff + 8 * j) = pl;
              }
            }
          }
          {                                      // 64-channel group complete: one tensor store per plane
            fence_proxy_async_smem();
            i5_epi_bar();
            if (epi_leader) {
              i5_tma_store_3d(&tmO_hi, sOut, n0 + (c & ~63), x0, y0);
              i5_tma_store_3d(&tmO_lo, sOut + I5_OUT_PLANE, n0 + (c & ~63), x0, y0);
              i5_bulk_commit();
            }
          }
        }
      }
      }
      if (!owner) {
        // publish the partial tile: every epilogue thread's stores -> gpu scope, then one release store of the flag
        __threadfence();
        i5_epi_bar();
        if (epi_leader) i5_st_release(prm.flags + cta, prm.epoch);
      }
      if (tr_me && seg == 0) tr[T5_CLK_EPI_FIRST] = (unsigned long long)clock64();
      u += (ke - ks);
    }
    // shared memory may go once the tensor stores have READ it; their global writes complete with the grid
    if (epi_leader) i5_bulk_wait_read0();
    if (tr_me) {
      tr[T5_CLK_EPI_END] = (unsigned long long)clock64();
      tr[T5_W_FLAGS] = (unsigned long long)w_flags;
      tr[T5_W_TMEM_FULL] = (unsigned long long)w_tfull;
    }
  } else if (prm.fmask && prm.kreg < prm.kchunks) {
    // ===================== mask warps (10-11; launched only for a masked fused term) =====================
    // m in {0,1}: m_p * (F_p . G) is the product with the masked pixels' rows of F zeroed.  A pixel is one 128-byte
    // row of the halo (the swizzle permutes 16-byte chunks inside it); only the centre 16 x 8 pixels are read by the
    // centre tap.  Each CTA masks its own halo (landed on its own f_land barrier) and hands it to the leader's MMA
    // thread through a_masked (4 warp arrivals, cluster-scope release after a proxy fence).
    const int tid = (int)threadIdx.x - I5_THREADS;                 // 0..63
    const int h0 = (int)(u0 / 9), h1 = (int)((u1 + 8) / 9);        // fused launches split at whole halo units
    UnitWalk wk;
    wk.init(h0, ipt, prm.tiles_n);
    uint32_t par = 0;                                              // bit b = parity of the next phase of f_land[b]
    for (int h = h0; h < h1; ++h, wk.next(ipt, prm.tiles_n)) {
      if (wk.kc < prm.kreg) continue;
      const int abuf = (h - h0) & 1;
      const int m_tile = 2 * wk.pm + (int)rank;
      const int y0 = (m_tile / prm.tiles_x) * I5_TH, x0 = (m_tile % prm.tiles_x) * I5_TW;
      float m[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = tid + 64 * j, y = y0 + (r >> 3), x = x0 + (r & 7);
        m[j] = (y < prm.H && x < prm.W) ? __ldg(prm.fmask + (int64_t)y * prm.W + x) : 1.f;   // outside: TMA zero fill
      }
      mbar_wait(&f_land[abuf], (par >> abuf) & 1u, 68);
      par ^= 1u << abuf;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (m[j] == 0.f) {
          const int r = tid + 64 * j;
          uint8_t* px = sA + abuf * I5_A_BUF + (1 + (r >> 3)) * I5_PITCH + (1 + (r & 7)) * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<uint4*>(px + c * 16) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(px + I5_A_PLANE + c * 16) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) i5_arrive_cta_cluster(&a_masked[abuf], 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  i5_cluster_sync();                         // the peer's shared memory / TMEM stay alive until both CTAs are done
  if (warp == 1) i5_tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (tr && threadIdx.x == 0) {
    tr[T5_CLK_OUT] = (unsigned long long)clock64();
    tr[T5_GT_OUT] = i5_global_ns();
  }
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
int igemm_streamk_workspace(float** ws, unsigned int** flags, unsigned int* epoch);   // tc_igemm_v2.cu
unsigned long long* get_igemm_trace();                                                // tc_igemm_v2.cu

template <int BN>
static int launch_igemm_ph_bn(const Act& a, const PackedB& b, const Epilogue& ep, const FusedTerm* ft, cudaStream_t st) {
  using Cfg = I5Cfg<BN>;
  IGemm5Params prm;
  int rc = igemm_streamk_workspace(&prm.ws, &prm.flags, &prm.epoch);
  if (rc) return rc;
  prm.H = a.H;
  prm.W = a.W;
  prm.tiles_x = ceil_div(a.W, I5_TW);
  prm.tiles_m = prm.tiles_x * ceil_div(a.H, I5_TH);
  prm.tiles_n = b.N / BN;
  prm.kreg = b.K / 64;
  prm.kchunks = prm.kreg + (ft ? ft->f.C / 64 : 0);
  prm.fmask = ft ? ft->rowmask : nullptr;
  prm.N = b.N;
  const long long pair_tiles = (long long)ceil_div(prm.tiles_m, 2) * prm.tiles_n;
  prm.total_units = pair_tiles * prm.kchunks * 9;
  // Split at whole halo units (no halo is loaded twice, layers with one K-chunk get no partial tiles at all) unless
  // that leaves the busiest pair more than 10 % above the mean (conv5_1: 160 units on 74 pairs = 3 vs 2.16).
  {
    const long long units = pair_tiles * prm.kchunks, g = std::min<long long>(74, units);
    const double mean = (double)units / (double)g;
    prm.align = ((double)((units + g - 1) / g) > 1.10 * mean) ? 1 : 9;
    if (ft) prm.align = 9;      // a fused chunk works on its centre tap only: ranges must not start inside a halo unit
  }
  prm.ep = ep;
  // bf16 planes leave through shared memory + TMA; everything else (fp32 rows, masked copies, planar image gradient)
  // keeps the per-thread stores
  prm.tma_out = 0;
  if (BN >= 64 && !ep.outm_hi && !ep.out_planar3) {
    if (ep.out_hi && ep.out_lo && !ep.out_f32) prm.tma_out = ep.pool2x2 ? 3 : 1;
    else if (ep.out_f32 && !ep.out_hi) prm.tma_out = 2;
  }
  SMB_REQUIRE(!ep.pool2x2 || prm.tma_out == 3, "igemm_ph: the pooled epilogue needs hi/lo outputs only and N %% 64 == 0");
  SMB_REQUIRE(!ep.pool_hi || (prm.tma_out == 1 && ep.pool_lo),
              "igemm_ph: the pooled side output needs hi/lo outputs through the staged epilogue and N %% 64 == 0");
  static int no_tma_out = -1;
  if (no_tma_out < 0) {
    const char* e = getenv("SMB_PH_DIRECT_STORES");     // experiment knob: 1 = per-thread global stores as in igemm_tc2
    no_tma_out = (e && atoi(e)) ? 1 : 0;
  }
  if (no_tma_out && prm.tma_out != 3) prm.tma_out = 0;
  prm.resident = (prm.kchunks == 1 && prm.tiles_n == 1 && Cfg::NB >= 9) ? 1 : 0;
  {
    static int no_resident = -1;
    if (no_resident < 0) {
      const char* e = getenv("SMB_PH_NO_RESIDENT");     // experiment knob
      no_resident = (e && atoi(e)) ? 1 : 0;
    }
    if (no_resident) prm.resident = 0;
  }
  prm.trace = get_igemm_trace();
  static int knob = -1;
  if (knob < 0) {
    const char* e = getenv("SMB_PH_KNOB");
    knob = e ? atoi(e) : 1;      // default: bit 0 set - the next halo is requested as soon as its buffer is free (tried at
                                 // every tap, forced at the last) instead of blocking for it at tap 3, which also held
                                 // back the B tiles of taps 3..8: +0.6 % on the step (profiles/r02o_bench_{product,knob1}.json)
  }
  prm.knob = knob;

  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo, tmO_hi, tmO_lo, tmF_hi, tmF_lo, tmG_hi, tmG_lo, tmA6_hi, tmA6_lo;
  {
    const uint64_t dims[3] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H};
    const uint64_t strides[2] = {(uint64_t)a.C * 2, (uint64_t)a.W * a.C * 2};
    const uint32_t box[3] = {64u, (uint32_t)I5_HW, (uint32_t)I5_HR};   // whole halo: 18 rows x 10 pixels x 64 channels
    rc = make_tmap_bf16(&tmA_hi, a.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmA_lo, a.lo, 3, dims, strides, box);
    if (rc) return rc;
    const uint32_t box6[3] = {64u, (uint32_t)I5_HW, (uint32_t)(I5_HR / 3)};   // a third of it: 6 rows
    rc = make_tmap_bf16(&tmA6_hi, a.hi, 3, dims, strides, box6);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmA6_lo, a.lo, 3, dims, strides, box6);
    if (rc) return rc;
  }
  {
    static int halo_split = -1;
    if (halo_split < 0) {
      const char* e = getenv("SMB_PH_HALO_SPLIT");
      halo_split = (e && atoi(e)) ? 1 : 0;
    }
    prm.halo_split = halo_split;
  }
  {
    const uint64_t dims[3] = {(uint64_t)b.K, (uint64_t)b.N, (uint64_t)b.taps};
    const uint64_t strides[2] = {(uint64_t)b.K * 2, (uint64_t)b.N * b.K * 2};
    const uint32_t box[3] = {64u, (uint32_t)(BN / 2), 1u};
    rc = make_tmap_bf16(&tmB_hi, b.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB_lo, b.lo, 3, dims, strides, box);
    if (rc) return rc;
  }
  if (prm.tma_out == 3) {
    const uint64_t dims[3] = {(uint64_t)b.N, (uint64_t)(a.W / 2), (uint64_t)(a.H / 2)};       // MaxPool2d floor mode
    const uint64_t strides[2] = {(uint64_t)b.N * 2, (uint64_t)(a.W / 2) * b.N * 2};
    const uint32_t box[3] = {64u, (uint32_t)(I5_TW / 2), (uint32_t)(I5_TH / 2)};
    rc = make_tmap_bf16(&tmO_hi, ep.out_hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmO_lo, ep.out_lo, 3, dims, strides, box);
    if (rc) return rc;
  } else if (prm.tma_out == 1) {
    const uint64_t dims[3] = {(uint64_t)b.N, (uint64_t)a.W, (uint64_t)a.H};
    const uint64_t strides[2] = {(uint64_t)b.N * 2, (uint64_t)a.W * b.N * 2};
    const uint32_t box[3] = {64u, (uint32_t)I5_TW, (uint32_t)I5_TH};
    rc = make_tmap_bf16(&tmO_hi, ep.out_hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmO_lo, ep.out_lo, 3, dims, strides, box);
    if (rc) return rc;
  } else if (prm.tma_out == 2) {
    const uint64_t dims[3] = {(uint64_t)b.N, (uint64_t)a.W, (uint64_t)a.H};
    const uint64_t strides[2] = {(uint64_t)b.N * 4, (uint64_t)a.W * b.N * 4};
    const uint32_t box[3] = {32u, (uint32_t)I5_TW, (uint32_t)I5_TH};          // 32 fp32 channels = one 128-byte row
    rc = make_tmap_f32(&tmO_hi, ep.out_f32, 3, dims, strides, box);
    if (rc) return rc;
    tmO_lo = tmO_hi;
  } else {
    tmO_hi = tmA_hi;      // never dereferenced
    tmO_lo = tmA_lo;
  }
  if (ft) {
    const uint64_t dims[3] = {(uint64_t)ft->f.C, (uint64_t)a.W, (uint64_t)a.H};
    const uint64_t strides[2] = {(uint64_t)ft->f.C * 2, (uint64_t)a.W * ft->f.C * 2};
    const uint32_t box[3] = {64u, (uint32_t)I5_HW, (uint32_t)I5_HR};
    rc = make_tmap_bf16(&tmF_hi, ft->f.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmF_lo, ft->f.lo, 3, dims, strides, box);
    if (rc) return rc;
    const uint64_t gdims[3] = {(uint64_t)ft->g.K, (uint64_t)ft->g.N, 1u};
    const uint64_t gstrides[2] = {(uint64_t)ft->g.K * 2, (uint64_t)ft->g.N * ft->g.K * 2};
    const uint32_t gbox[3] = {64u, (uint32_t)(BN / 2), 1u};
    rc = make_tmap_bf16(&tmG_hi, ft->g.hi, 3, gdims, gstrides, gbox);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmG_lo, ft->g.lo, 3, gdims, gstrides, gbox);
    if (rc) return rc;
  } else {
    tmF_hi = tmA_hi;      // never dereferenced
    tmF_lo = tmA_lo;
    tmG_hi = tmB_hi;
    tmG_lo = tmB_lo;
  }
  static bool attr_set = false;
  if (!attr_set) {
    SMB_CUDA_CHECK(cudaFuncSetAttribute(igemm_ph_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SMB_CUDA_CHECK(cudaGetDevice(&dev));
    SMB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (num_sms > 148) num_sms = 148;
  }
  const long long pairs = std::max<long long>(1, std::min<long long>(num_sms / 2, prm.total_units / prm.align));
  const int threads = (prm.fmask && prm.kreg < prm.kchunks) ? I5_THREADS_MASK : I5_THREADS;
  SMB_LAUNCH(igemm_ph_kernel<BN>, (unsigned)(2 * pairs), threads, Cfg::SMEM, st, tmA_hi, tmA_lo, tmB_hi, tmB_lo, tmO_hi,
             tmO_lo, tmF_hi, tmF_lo, tmG_hi, tmG_lo, tmA6_hi, tmA6_lo, prm);
  return SMB_OK;
}

int launch_igemm_ph(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st, const FusedTerm* ft) {
  SMB_REQUIRE(b.taps == 9, "igemm_ph: 3x3 convolutions only");
  if (ft) {
    SMB_REQUIRE(ft->f.H == a.H && ft->f.W == a.W && ft->f.C % 64 == 0 && ft->f.C > 0,
                "igemm_ph: the fused term's features must have the conv's spatial size and a multiple of 64 channels");
    SMB_REQUIRE(ft->g.taps == 1 && ft->g.N == b.N && ft->g.K == ft->f.C && b.N % 64 == 0,
                "igemm_ph: the fused term's matrix must be [N=%d][K=%d]", b.N, ft->f.C);
  }
  SMB_REQUIRE(a.C == b.K && b.K % 64 == 0 && (b.N % 64 == 0 || b.N == 16),
              "igemm_ph: K=%d must be a multiple of 64, N=%d a multiple of 64 (or the padded 16)", b.K, b.N);
  SMB_REQUIRE((long long)ceil_div(a.W, I5_TW) * ceil_div(a.H, I5_TH) * ceil_div(b.N, 64) *
                      (b.K / 64 + (ft ? ft->f.C / 64 : 0)) * 9 < (1LL << 30),
              "igemm_ph: problem too large for 32-bit work indices");
  if (a.pixels() == 0) return SMB_OK;
  if (b.N == 16) return launch_igemm_ph_bn<16>(a, b, ep, nullptr, st);
  if (b.N % 128 == 0) return launch_igemm_ph_bn<128>(a, b, ep, ft, st);
  return launch_igemm_ph_bn<64>(a, b, ep, ft, st);
}

}  // namespace smb
