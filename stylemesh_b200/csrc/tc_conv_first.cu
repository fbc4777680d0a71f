// conv_first_tc_kernel — VGG conv1_1 (3 -> 64 channels) on tcgen05 with a software im2col producer.
//
// The first layer has K = 27 (3 channels x 3x3): no 64-channel rows for TMA to fetch, and on CUDA cores it is
// weight-LDS / store bound (77 us per 640x480 image for 1 GFLOP).  Here four producer warps build the
// [128 pixels][K = 32 (27 + 5 zero)] bf16 hi/lo operand tiles directly in shared memory in the 128-byte-swizzled
// K-major layout UMMA expects (generic-proxy stores + fence.proxy.async), one thread issues the 2 x 3 MMAs per tile
// (N = 64, main/corr accumulators as in igemm_tc2), and four epilogue warps apply bias + ReLU and write the hi/lo
// activation planes.  Persistent CTAs, double-buffered TMEM accumulators: build(i+1) | MMA(i) | store(i-1); the A tile
// is single-buffered (the producers hold tile i+1 in registers while the ~200-cycle MMAs of tile i drain it).
// The planes leave through a swizzled shared-memory tile and one TMA tensor store per plane: per-thread 16-byte
// stores at a 128-byte stride (32 L1 requests per instruction) made this kernel L1-throughput bound (ncu: L1/TEX 69 %,
// DRAM 9 %, issue slots 28 % busy).
#include "tc_common.cuh"
#include "smb_epilogue.cuh"
#include "smb_kernels.h"

namespace smb {
using namespace tc;

constexpr int CF_THREADS = 288;          // warps 0-3 producers, warp 4 MMA, warps 5-8 epilogue
constexpr int CF_BM = 128;
constexpr int CF_N = 64;
constexpr int CF_K = 32;                 // 27 real + 5 zero
constexpr int CF_A_BYTES = CF_BM * 128;  // one plane: 128 rows x 128-byte swizzle rows (first 64 bytes carry K = 32)
constexpr int CF_B_BYTES = CF_N * 128;
constexpr int CF_OUT_BYTES = CF_BM * 128; // one plane of the staged output tile: 128 pixels x 64 channels bf16
constexpr int CF_SMEM = 2 * CF_A_BYTES + 2 * CF_B_BYTES + 2 * CF_OUT_BYTES + 1024 + 256;
constexpr int CF_TMEM_COLS = 256;        // 2 buffers x (main 64 + corr 64)

struct ConvFirstParams {
  const float* img;                      // (3, H, W) fp32
  const __nv_bfloat16* w_hi;             // [64][32] (k = ci*9 + r*3 + s, zero padded)
  const __nv_bfloat16* w_lo;
  int H, W, TH, TW, tiles_x, tiles;
  int tma_out;                           // 1: hi/lo planes leave through shared memory + TMA tensor stores
  Epilogue ep;
};

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// byte offset of 16-byte chunk `j` of row `r` inside a 1024-byte-aligned SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

__global__ void __launch_bounds__(CF_THREADS, 2)
conv_first_tc_kernel(const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
                     const ConvFirstParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // [plane][CF_A_BYTES]
  uint8_t* sB = smem + 2 * CF_A_BYTES;                  // [plane][CF_B_BYTES]
  uint8_t* sOut = sB + 2 * CF_B_BYTES;                  // [plane][CF_OUT_BYTES]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sOut + 2 * CF_OUT_BYTES);   // [2] count 128 (producer threads)
  uint64_t* a_empty = a_full + 2;                       // [2] count 1 (tcgen05.commit)
  uint64_t* t_full = a_empty + 2;                       // [2] count 1 (tcgen05.commit)
  uint64_t* t_empty = t_full + 2;                       // [2] count 4 (epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t P = (int64_t)prm.H * prm.W;

  // weights -> swizzled K-major smem tiles (generic proxy), once per CTA
  for (int i = threadIdx.x; i < CF_N * 4 * 2; i += blockDim.x) {
    const int plane = i / (CF_N * 4), rem = i % (CF_N * 4), n = rem >> 2, j = rem & 3;
    const uint4 v = *reinterpret_cast<const uint4*>((plane ? prm.w_lo : prm.w_hi) + n * CF_K + j * 8);
    *reinterpret_cast<uint4*>(sB + plane * CF_B_BYTES + sw128_off(n, j)) = v;
  }
  if (threadIdx.x == 0) {
    if (prm.tma_out) {
      tma_prefetch_desc(&tmO_hi);
      tma_prefetch_desc(&tmO_lo);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_full[b], 128);
      mbar_init(&a_empty[b], 1);
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, CF_TMEM_COLS);
    tmem_relinquish();
  }
  fence_proxy_async_smem();                // weight tiles visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();                              // prologue (weights are constants) overlaps the preceding launch's tail

  if (warp < 4) {
    // ===================== im2col producers: thread r builds row r of the A tile =====================
    const int r = threadIdx.x;             // 0..127
    int it = 0;
    for (int tile = blockIdx.x; tile < prm.tiles; tile += gridDim.x, ++it) {
      const int y0 = (tile / prm.tiles_x) * prm.TH, x0 = (tile % prm.tiles_x) * prm.TW;
      const int y = y0 + r / prm.TW, x = x0 + r % prm.TW;
      float in[CF_K];
#pragma unroll
      for (int k = 27; k < CF_K; ++k) in[k] = 0.f;
      const bool inside = (y < prm.H) && (x < prm.W);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int ss = 0; ss < 3; ++ss) {
            const int yy = y + rr - 1, xx = x + ss - 1;
            in[ci * 9 + rr * 3 + ss] = (inside && yy >= 0 && yy < prm.H && xx >= 0 && xx < prm.W)
                                           ? __ldg(prm.img + (int64_t)ci * P + (int64_t)yy * prm.W + xx)
                                           : 0.f;
          }
      mbar_wait(&a_empty[0], ((uint32_t)it & 1u) ^ 1u, 41);  // the MMAs of the previous tile have read the buffer
      uint8_t* a_hi = sA;
      uint8_t* a_lo = sA + CF_A_BYTES;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 h, l;
        split2_pack(in[8 * j + 0], in[8 * j + 1], h.x, l.x);
        split2_pack(in[8 * j + 2], in[8 * j + 3], h.y, l.y);
        split2_pack(in[8 * j + 4], in[8 * j + 5], h.z, l.z);
        split2_pack(in[8 * j + 6], in[8 * j + 7], h.w, l.w);
        *reinterpret_cast<uint4*>(a_hi + sw128_off(r, j)) = h;
        *reinterpret_cast<uint4*>(a_lo + sw128_off(r, j)) = l;
      }
      fence_proxy_async_smem();            // generic-proxy stores -> visible to tcgen05 (async proxy)
      mbar_arrive(&a_full[0]);
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(CF_BM, CF_N, 0, 0);
      const uint32_t b_hi = smem_u32(sB), b_lo = b_hi + CF_B_BYTES;
      int it = 0;
      for (int tile = blockIdx.x; tile < prm.tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&t_empty[buf], (use & 1u) ^ 1u, 42);
        mbar_wait(&a_full[0], (uint32_t)it & 1u, 43);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + CF_A_BYTES;
        const uint32_t t_main = tmem_base + (uint32_t)(buf * 2 * CF_N), t_corr = t_main + CF_N;
#pragma unroll
        for (int k = 0; k < CF_K / 16; ++k) {
          const uint64_t dah = make_smem_desc_sw128(a_hi + k * 32, 16, 1024);
          const uint64_t dal = make_smem_desc_sw128(a_lo + k * 32, 16, 1024);
          const uint64_t dbh = make_smem_desc_sw128(b_hi + k * 32, 16, 1024);
          const uint64_t dbl = make_smem_desc_sw128(b_lo + k * 32, 16, 1024);
          umma_f16(t_corr, dal, dbh, idesc, (uint32_t)(k > 0));
          umma_f16(t_corr, dah, dbl, idesc, 1u);
          umma_f16(t_main, dah, dbh, idesc, (uint32_t)(k > 0));
        }
        umma_commit(&a_empty[0]);
        umma_commit(&t_full[buf]);
      }
    }
  } else {
    // ===================== epilogue: bias + ReLU + hi/lo planes =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool epi_leader = (warp == 5 && lane == 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < prm.tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int y0 = (tile / prm.tiles_x) * prm.TH, x0 = (tile % prm.tiles_x) * prm.TW;
      const int yy = y0 + row / prm.TW, xx = x0 + row % prm.TW;
      const bool valid = (yy < prm.H) && (xx < prm.W);
      const int64_t p = (int64_t)yy * prm.W + xx;
      mbar_wait(&t_full[buf], use & 1u, 44);
      tc_fence_after();
      const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * CF_N);
      if (prm.tma_out) {                   // staging tile free once the previous tile's tensor stores have read it
        if (epi_leader) bulk_wait_read0();
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
#pragma unroll 1
      for (int c = 0; c < CF_N; c += 32) {
        uint32_t rm[32], rc[32];
        tmem_ld_32x32(t_main + (uint32_t)c, rm);
        tmem_ld_32x32(t_main + (uint32_t)(CF_N + c), rc);
        tmem_ld_wait();
        if (c + 32 >= CF_N) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rm[j]) + __uint_as_float(rc[j]);
        if (!prm.tma_out) {
          if (valid) epilogue_store<32>(prm.ep, p, c, CF_N, v);
          continue;
        }
        epilogue_apply<32>(prm.ep, p, c, CF_N, v, valid);
        // pixel `row` = 128-byte row `row` of the staging tile; 16-byte chunk index XOR (row & 7) is the SWIZZLE_128B
        // pattern of the tensor store (conflict free: 8 lanes cover 8 distinct chunks)
        const uint32_t rbase = smem_u32(sOut) + (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 h, l;
          split2_pack(v[8 * j], v[8 * j + 1], h.x, l.x);
          split2_pack(v[8 * j + 2], v[8 * j + 3], h.y, l.y);
          split2_pack(v[8 * j + 4], v[8 * j + 5], h.z, l.z);
          split2_pack(v[8 * j + 6], v[8 * j + 7], h.w, l.w);
          const uint32_t a = rbase + ((((uint32_t)(c >> 3) + (uint32_t)j) ^ sw) << 4);
          sts128(a, h);
          sts128(a + CF_OUT_BYTES, l);
        }
      }
      if (prm.tma_out) {                   // one tensor store per plane; pixels outside the image are clipped by TMA
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_leader) {
          tma_store_3d(&tmO_hi, sOut, 0, x0, y0);
          tma_store_3d(&tmO_lo, sOut + CF_OUT_BYTES, 0, x0, y0);
          bulk_commit();
        }
      }
    }
    if (prm.tma_out && epi_leader) bulk_wait_read0();   // shared memory may go once the stores have read it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, CF_TMEM_COLS);
}

int launch_conv_first_tc(const float* img, int H, int W, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo,
                         const float* bias, const Epilogue& ep_in, cudaStream_t st) {
  if ((int64_t)H * W == 0) return SMB_OK;
  ConvFirstParams prm;
  prm.img = img;
  prm.w_hi = w_hi;
  prm.w_lo = w_lo;
  prm.H = H;
  prm.W = W;
  // 128-pixel patch: minimise padding, prefer wide patches (coalesced image reads along x)
  int best_th = 8, best_tw = 16;
  int64_t best_area = -1;
  for (int th = 1; th <= 16; th <<= 1) {
    const int tw = 128 / th;
    const int64_t area = (int64_t)ceil_div(H, th) * th * ceil_div(W, tw) * tw;
    if (best_area < 0 || area < best_area) {
      best_area = area;
      best_th = th;
      best_tw = tw;
    }
  }
  prm.TH = best_th;
  prm.TW = best_tw;
  prm.tiles_x = ceil_div(W, prm.TW);
  prm.tiles = prm.tiles_x * ceil_div(H, prm.TH);
  prm.ep = ep_in;
  prm.ep.bias = bias;
  const Epilogue& e = prm.ep;
  prm.tma_out = (e.out_hi && e.out_lo && !e.out_f32 && !e.outm_hi && !e.out_planar3) ? 1 : 0;
  CUtensorMap tmO_hi, tmO_lo;
  memset(&tmO_hi, 0, sizeof(tmO_hi));
  memset(&tmO_lo, 0, sizeof(tmO_lo));
  if (prm.tma_out) {
    const uint64_t dims[3] = {(uint64_t)CF_N, (uint64_t)W, (uint64_t)H};
    const uint64_t strides[2] = {(uint64_t)CF_N * 2, (uint64_t)W * CF_N * 2};
    const uint32_t box[3] = {64u, (uint32_t)prm.TW, (uint32_t)prm.TH};
    int rc = make_tmap_bf16(&tmO_hi, e.out_hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmO_lo, e.out_lo, 3, dims, strides, box);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    SMB_CUDA_CHECK(cudaFuncSetAttribute(conv_first_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM));
    attr_set = true;
  }
  const int grid = std::min(prm.tiles, 2 * 148);     // two CTAs per SM (83 KB smem, 256 TMEM columns each)
  SMB_LAUNCH(conv_first_tc_kernel, grid, CF_THREADS, CF_SMEM, st, tmO_hi, tmO_lo, prm);
  return SMB_OK;
}

}  // namespace smb
