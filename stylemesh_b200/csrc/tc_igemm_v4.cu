// igemm_halo_kernel — third generation of the 3x3 implicit-GEMM conv: the activation operand is loaded ONCE per
// K-chunk as a halo tile and reused by all nine filter taps.
//
// Why (measured on B200, tools/gpu_conv_probe.py): igemm_tc2 with every tcgen05.mma removed runs the VGG forward in
// 0.640 ms against 0.701 ms with them — the kernel is bound by the L2 -> shared-memory operand stream (7.3 GB per
// forward = 11.4 TB/s, ~39 B/clk/SM), not by the tensor pipe.  A tap-per-stage pipeline fetches the same activation
// pixels nine times (once per tap, shifted by one pixel).  Here a 16 x 8-pixel output tile loads its (16+2) x (8+2)
// halo once per 64-channel chunk — 18 row boxes of 10 pixels, each at a 2 KiB pitch so that every 8-pixel run is a
// 1024-byte-aligned swizzle atom — and the tap (dy, dx) is just a different UMMA descriptor on the same tile:
//     start = base + dy * pitch + dx * 128,  stride between 8-row groups (SBO) = pitch,  base_offset = 0
// (verified on B200: the tensor core applies the 128-byte swizzle on absolute shared-memory address bits, so any
//  128-byte-aligned start and any row pitch — 2048 or the dense 1280 — read back exactly what TMA wrote).
// Operand bytes per (tile, K-chunk) at BN = 128: 46 KB (A halo) + 9 x 32 KB (B) = 334 KB instead of 9 x 64 = 576 KB.
//
// Everything else is igemm_tc2: bf16 hi/lo three-pass products into main/corr TMEM accumulators (double buffered),
// persistent CTAs with stream-K over (tile, K-chunk) units, first-K-part ownership with epoch-flag fix-ups, fused
// epilogue, watchdog on every wait.
#include <cstdlib>

#include "tc_common.cuh"
#include "smb_epilogue.cuh"
#include "smb_kernels.h"

namespace smb {
using namespace tc;

constexpr int I4_THREADS = 192;
constexpr int I4_BM = 128;
constexpr int I4_TH = 16, I4_TW = 8;                  // output patch: 16 rows x 8 pixels
constexpr int I4_HR = I4_TH + 2, I4_HW = I4_TW + 2;   // halo: 18 rows x 10 pixels
constexpr int I4_A_TX = 2 * I4_HR * I4_HW * 128;      // bytes written by TMA per A buffer (hi + lo) = 46080
constexpr int I4_SMEM_EXTRA = 1024 + 256;
constexpr int I4_SMEM_TOTAL = 224 * 1024;             // ring budget + extra must stay below the 227 KiB opt-in limit
constexpr int I4_MAX_NB = 6;

struct IGemm4Params {
  int H, W, tiles_x, tiles_n, kchunks, N;
  long long total_units;      // tiles * kchunks
  float* ws;
  unsigned int* flags;
  unsigned int epoch;
  int desc_mode;              // 1 = descriptor base_offset 0 (correct: verified on B200), 0 = (start >> 7) & 7 (wrong)
  int row_pitch;              // bytes between halo rows in smem: 2048 (1 KiB-aligned 8-pixel atoms) or 1280 (dense)
  int nb;                     // B ring stages (one tap each)
  Epilogue ep;
};

template <int BN>
struct I4Cfg {
  static constexpr int B_STAGE = 2 * BN * 128;                  // hi + lo of one tap
  static constexpr int TMEM_COLS = (2 * 2 * BN) <= 256 ? 256 : 512;
};

__device__ __forceinline__ uint64_t make_smem_desc_sw128_bo(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;                                   // LBO (unused for swizzled K-major) = 16 B
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
  d |= (uint64_t)(base_off & 7u) << 49;                     // matrix base offset: (start_address >> 7) & 7
  d |= (uint64_t)2 << 61;                                   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu4(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu4(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(I4_THREADS, 1)
igemm_halo_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmA_row_hi, const __grid_constant__ CUtensorMap tmA_row_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const IGemm4Params prm) {
  using Cfg = I4Cfg<BN>;
  const long long G = gridDim.x, cta = blockIdx.x;
  const long long u0 = cta * prm.total_units / G, u1 = (cta + 1) * prm.total_units / G;
  if (u0 >= u1) return;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int I4_ROW_PITCH = prm.row_pitch;
  const int I4_A_PLANE = (I4_HR * I4_ROW_PITCH + 1023) & ~1023;   // keep every plane 1 KiB aligned
  const int I4_A_BUF = 2 * I4_A_PLANE;
  const int NB = prm.nb;
  uint8_t* sA = smem;                                         // [2 bufs][hi, lo][18 halo rows]
  uint8_t* sB = smem + 2 * I4_A_BUF;                          // [NB][hi, lo][BN x 128 B]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + NB * Cfg::B_STAGE);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + I4_MAX_NB;
  uint64_t* tmem_full_bar = b_empty + I4_MAX_NB;              // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ipt = prm.kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmA_row_hi);
    tma_prefetch_desc(&tmA_row_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Program order: A(u0); then per unit u: B taps 0..8, with A(u+1) issued after tap 2 so that the halo of the next
    // K-chunk is in flight long before the tensor pipe needs it (it only needs the other A buffer to be drained).
    if (elect_one()) {
      long long ga_issue = 0, gb = 0;
      auto issue_A = [&](long long u) {
        const int tile = (int)(u / ipt), kc = (int)(u % ipt);
        const int m_tile = tile / prm.tiles_n;
        const int y0 = (m_tile / prm.tiles_x) * I4_TH, x0 = (m_tile % prm.tiles_x) * I4_TW;
        const int abuf = (int)(ga_issue & 1);
        mbar_wait(&a_empty[abuf], (uint32_t)((ga_issue >> 1) & 1) ^ 1u, 51);
        mbar_arrive_expect_tx(&a_full[abuf], I4_A_TX);
        uint8_t* ah = sA + abuf * I4_A_BUF;
        uint8_t* al = ah + I4_A_PLANE;
        if (I4_ROW_PITCH == I4_HW * 128) {      // dense rows: the whole 18 x 10-pixel halo is ONE box per plane
          tma_load_3d(ah, &tmA_hi, &a_full[abuf], kc * 64, x0 - 1, y0 - 1);
          tma_load_3d(al, &tmA_lo, &a_full[abuf], kc * 64, x0 - 1, y0 - 1);
        } else {
#pragma unroll 1
          for (int r = 0; r < I4_HR; ++r) {     // halo row r = image row y0 - 1 + r, pixels x0 - 1 .. x0 + 8
            tma_load_3d(ah + r * I4_ROW_PITCH, &tmA_row_hi, &a_full[abuf], kc * 64, x0 - 1, y0 - 1 + r);
            tma_load_3d(al + r * I4_ROW_PITCH, &tmA_row_lo, &a_full[abuf], kc * 64, x0 - 1, y0 - 1 + r);
          }
        }
        ++ga_issue;
      };
      issue_A(u0);
      for (long long u = u0; u < u1; ++u) {
        const int tile = (int)(u / ipt), kc = (int)(u % ipt);
        const int n_tile = tile % prm.tiles_n;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, ++gb) {
          if (tap == 3 && u + 1 < u1) issue_A(u + 1);
          const int bs = (int)(gb % NB);
          mbar_wait(&b_empty[bs], (uint32_t)((gb / NB) & 1) ^ 1u, 52);
          mbar_arrive_expect_tx(&b_full[bs], Cfg::B_STAGE);
          uint8_t* bh = sB + bs * Cfg::B_STAGE;
          tma_load_3d(bh, &tmB_hi, &b_full[bs], kc * 64, n_tile * BN, tap);
          tma_load_3d(bh + BN * 128, &tmB_lo, &b_full[bs], kc * 64, n_tile * BN, tap);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(I4_BM, BN, 0, 0);
      long long ga = 0, gb = 0;
      int seg = 0;
      for (long long u = u0; u < u1; ++seg) {
        const int ks = (int)(u % ipt);
        const long long left = u1 - u;
        const int ke = (left < (long long)(ipt - ks)) ? ks + (int)left : ipt;
        const int buf = seg & 1;
        const uint32_t use = (uint32_t)(seg >> 1);
        mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u, 53);
        tc_fence_after();
        const uint32_t t_main = tmem_base + (uint32_t)(buf * 2 * BN);
        const uint32_t t_corr = t_main + (uint32_t)BN;
        for (int kc = ks; kc < ke; ++kc, ++ga) {
          const int abuf = (int)(ga & 1);
          mbar_wait(&a_full[abuf], (uint32_t)((ga >> 1) & 1), 54);
          const uint32_t a_hi = smem_u32(sA + abuf * I4_A_BUF);
          const uint32_t a_lo = a_hi + I4_A_PLANE;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap, ++gb) {
            const int bs = (int)(gb % NB);
            mbar_wait(&b_full[bs], (uint32_t)((gb / NB) & 1), 55);
            tc_fence_after();
            const int dy = tap / 3, dx = tap % 3;            // halo coordinates of the tap's top-left pixel
            const uint32_t a_off = (uint32_t)(dy * I4_ROW_PITCH + dx * 128);
            const uint32_t bo = (prm.desc_mode == 0) ? (((a_hi + a_off) >> 7) & 7u) : 0u;   // a_lo has the same phase
            const uint32_t b_hi = smem_u32(sB + bs * Cfg::B_STAGE);
            const uint32_t b_lo = b_hi + BN * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t dah = make_smem_desc_sw128_bo(a_hi + a_off + k * 32, I4_ROW_PITCH, bo);
              const uint64_t dal = make_smem_desc_sw128_bo(a_lo + a_off + k * 32, I4_ROW_PITCH, bo);
              const uint64_t dbh = make_smem_desc_sw128(b_hi + k * 32, 16, 1024);
              const uint64_t dbl = make_smem_desc_sw128(b_lo + k * 32, 16, 1024);
              const uint32_t acc = (uint32_t)((kc > ks) || (tap > 0) || (k > 0));
              umma_f16(t_corr, dal, dbh, idesc, acc);
              umma_f16(t_corr, dah, dbl, idesc, 1u);
              umma_f16(t_main, dah, dbh, idesc, acc);
            }
            umma_commit(&b_empty[bs]);
          }
          umma_commit(&a_empty[abuf]);
        }
        umma_commit(&tmem_full_bar[buf]);
        u += (ke - ks);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* my_slot = prm.ws + (size_t)cta * I4_BM * BN;
    int seg = 0;
    for (long long u = u0; u < u1; ++seg) {
      const int tile = (int)(u / ipt);
      const int ks = (int)(u % ipt);
      const long long left = u1 - u;
      const int ke = (left < (long long)(ipt - ks)) ? ks + (int)left : ipt;
      const int buf = seg & 1;
      const uint32_t use = (uint32_t)(seg >> 1);
      const bool owner = (ks == 0);
      const int m_tile = tile / prm.tiles_n, n_tile = tile % prm.tiles_n;
      const int y0 = (m_tile / prm.tiles_x) * I4_TH, x0 = (m_tile % prm.tiles_x) * I4_TW;
      const int yy = y0 + row / I4_TW, xx = x0 + row % I4_TW;
      const bool valid = (yy < prm.H) && (xx < prm.W);
      const int64_t p = (int64_t)yy * prm.W + xx;
      const int n0 = n_tile * BN;

      int npeer = 0;
      if (owner && ke < ipt) {
        const long long tile_end = (long long)(tile + 1) * ipt;
        long long c = cta + 1;
        while (c < G && c * prm.total_units / G < tile_end) {
          if (lane == 0) {
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile const unsigned int*>(prm.flags + c) != prm.epoch) {
              __nanosleep(64);
              if (clock64() - t0 > 4000000000LL) {
                printf("[smb] igemm_halo stream-K watchdog: CTA %d waiting for partial of CTA %d\n", (int)cta, (int)c);
                asm volatile("trap;");
              }
            }
          }
          __syncwarp();
          (void)ld_acquire_gpu4(prm.flags + c);
          ++npeer;
          ++c;
        }
      }

      mbar_wait(&tmem_full_bar[buf], use & 1u, 56);
      tc_fence_after();
      const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN);
      const uint32_t t_corr = t_main + (uint32_t)BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t rm[32], rc[32];
        tmem_ld_32x32(t_main + (uint32_t)c, rm);
        tmem_ld_32x32(t_corr + (uint32_t)c, rc);
        tmem_ld_wait();
        if (c + 32 >= BN) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]);
        for (int k = 1; k <= npeer; ++k) {
          const float4* src = reinterpret_cast<const float4*>(prm.ws + (size_t)(cta + k) * I4_BM * BN +
                                                              (size_t)row * BN + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = src[j];
            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
          }
        }
        if (owner) {
          if (valid) epilogue_store<32>(prm.ep, p, n0 + c, prm.N, v);
        } else {
          float4* dst = reinterpret_cast<float4*>(my_slot + (size_t)row * BN + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      if (!owner) {
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) st_release_gpu4(prm.flags + cta, prm.epoch);
      }
      u += (ke - ks);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
int igemm_streamk_workspace(float** ws, unsigned int** flags, unsigned int* epoch);   // tc_igemm_v2.cu

template <int BN>
static int launch_igemm_halo_bn(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st) {
  using Cfg = I4Cfg<BN>;
  IGemm4Params prm;
  int rc = igemm_streamk_workspace(&prm.ws, &prm.flags, &prm.epoch);
  if (rc) return rc;
  prm.H = a.H;
  prm.W = a.W;
  prm.tiles_x = ceil_div(a.W, I4_TW);
  const int tiles_y = ceil_div(a.H, I4_TH);
  prm.tiles_n = b.N / BN;
  prm.kchunks = b.K / 64;
  prm.N = b.N;
  const long long tiles = (long long)prm.tiles_x * tiles_y * prm.tiles_n;
  prm.total_units = tiles * prm.kchunks;
  prm.ep = ep;
  static int desc_mode = -1;
  if (desc_mode < 0) {
    const char* e = getenv("SMB_HALO_DESC_MODE");
    desc_mode = e ? atoi(e) : 1;     // measured: the tensor core swizzles on absolute smem address bits -> base_offset 0
  }
  prm.desc_mode = desc_mode;
  static int pitch = 0;
  if (!pitch) {
    const char* e = getenv("SMB_HALO_PITCH");
    pitch = e ? atoi(e) : 1280;      // dense halo rows work too (SBO = 1280) and leave room for a deeper B ring
    if (pitch != 2048 && pitch != 1280) pitch = 1280;
  }
  prm.row_pitch = pitch;
  const int a_plane = (I4_HR * pitch + 1023) & ~1023;
  prm.nb = std::min(I4_MAX_NB, (I4_SMEM_TOTAL - 4 * a_plane) / Cfg::B_STAGE);
  SMB_REQUIRE(prm.nb >= 2, "igemm_halo: shared memory budget leaves %d B stages", prm.nb);

  CUtensorMap tmA_hi, tmA_lo, tmA_row_hi, tmA_row_lo, tmB_hi, tmB_lo;
  {
    const uint64_t dims[3] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H};
    const uint64_t strides[2] = {(uint64_t)a.C * 2, (uint64_t)a.W * a.C * 2};
    const uint32_t box[3] = {64u, (uint32_t)I4_HW, (uint32_t)I4_HR};   // whole halo: 18 rows x 10 pixels x 64 channels
    rc = make_tmap_bf16(&tmA_hi, a.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmA_lo, a.lo, 3, dims, strides, box);
    if (rc) return rc;
    const uint32_t rbox[3] = {64u, (uint32_t)I4_HW, 1u};               // one halo row (2048-byte pitch variant)
    rc = make_tmap_bf16(&tmA_row_hi, a.hi, 3, dims, strides, rbox);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmA_row_lo, a.lo, 3, dims, strides, rbox);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)b.K, (uint64_t)b.N, (uint64_t)b.taps};
    const uint64_t strides[2] = {(uint64_t)b.K * 2, (uint64_t)b.N * b.K * 2};
    const uint32_t box[3] = {64u, (uint32_t)BN, 1u};
    rc = make_tmap_bf16(&tmB_hi, b.hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB_lo, b.lo, 3, dims, strides, box);
    if (rc) return rc;
  }
  const int smem_bytes = 4 * a_plane + prm.nb * Cfg::B_STAGE + I4_SMEM_EXTRA;
  static bool attr_set = false;
  if (!attr_set) {
    SMB_CUDA_CHECK(cudaFuncSetAttribute(igemm_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        I4_SMEM_TOTAL + I4_SMEM_EXTRA));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SMB_CUDA_CHECK(cudaGetDevice(&dev));
    SMB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (num_sms > 148) num_sms = 148;
  }
  const int grid = (int)std::max<long long>(1, std::min<long long>(num_sms, prm.total_units));
  igemm_halo_kernel<BN><<<grid, I4_THREADS, smem_bytes, st>>>(tmA_hi, tmA_lo, tmA_row_hi, tmA_row_lo, tmB_hi, tmB_lo,
                                                              prm);
  SMB_LAUNCH_CHECK();
  return SMB_OK;
}

int launch_igemm_halo(const Act& a, const PackedB& b, const Epilogue& ep, cudaStream_t st) {
  SMB_REQUIRE(b.taps == 9, "igemm_halo: 3x3 convolutions only");
  SMB_REQUIRE(a.C == b.K && b.K % 64 == 0 && b.N % 64 == 0, "igemm_halo: K=%d, N=%d must be multiples of 64", b.K, b.N);
  if (a.pixels() == 0) return SMB_OK;
  if (b.N % 128 == 0) return launch_igemm_halo_bn<128>(a, b, ep, st);
  return launch_igemm_halo_bn<64>(a, b, ep, st);
}

}  // namespace smb
