// stylemesh_b200 — common device/host helpers for the sm_100a kernels.
//
// Storage convention used by every VGG-side kernel ("Act" planes):
//   an activation / gradient map of P = H*W pixels and C channels is kept channels-last as two
//   bf16 planes  hi[P][C], lo[P][C]  with  x ~= float(hi) + float(lo)   (|err| <= 2^-17 |x|).
//   Same bytes as fp32, but directly consumable by tcgen05 kind::f16 MMAs as a 3-pass
//   (hi*hi + lo*hi + hi*lo) fp32-grade product, and by TMA as 128-byte channel rows.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace smb {

// ------------------------------------------------------------------------------------------
// error plumbing (C-ABI returns int status; message retrievable through smb_last_error()).
// ------------------------------------------------------------------------------------------
enum : int {
  SMB_OK = 0,
  SMB_ERR_CUDA = -1,
  SMB_ERR_ARG = -2,
  SMB_ERR_UNSUPPORTED = -3,
  SMB_ERR_STATE = -4,
};

void set_error(const char* fmt, ...);
const char* get_error();

#define SMB_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::smb::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,              \
                       cudaGetErrorString(_e));                                           \
      return ::smb::SMB_ERR_CUDA;                                                         \
    }                                                                                     \
  } while (0)

#define SMB_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::smb::set_error(__VA_ARGS__);                                                      \
      return ::smb::SMB_ERR_ARG;                                                          \
    }                                                                                     \
  } while (0)

// every kernel launch of this library goes through SMB_LAUNCH_CHECK: it also feeds the launch counter that
// bench.py reports as "gpu_launches"
void count_launch();
long long launch_count();
#define SMB_LAUNCH_CHECK()                 \
  do {                                     \
    ::smb::count_launch();                 \
    SMB_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A step is ~80 short launches on one stream; with the attribute below the grid of
// launch i+1 is scheduled as soon as every CTA of launch i has started (SMB_LAUNCH kernels trigger at their top),
// so its CTAs take SMs as they drain and run their prologue (barrier init, TMEM alloc, descriptor prefetch) under
// the tail of launch i.  Contract: every kernel launched through SMB_LAUNCH executes pdl_sync() before its first
// global-memory access and before any early return (griddepcontrol.wait returns once the preceding grid has
// completed and its writes are visible; it is a no-op for a launch without the attribute).  SMB_PDL=0 disables.
// ------------------------------------------------------------------------------------------
bool pdl_enabled();
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define SMB_LAUNCH(kernel, grid, block, smem, st, ...)                                               \
  do {                                                                                               \
    ::smb::count_launch();                                                                           \
    SMB_CUDA_CHECK(::smb::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)); \
  } while (0)
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------
// bf16 hi/lo split
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ float merge2(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}
// pack two floats' hi parts / lo parts into 32-bit words (element 0 in the low half).  Same values as split2() on
// each element (round-to-nearest-even both times), but with the packed two-element conversion: one F2FP per pair
// instead of two F2F (the scalar conversion issues at a quarter of the rate and dominated the conv epilogues).
__device__ __forceinline__ void split2_pack(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(b), "f"(a));          // d.hi = first source, d.lo = second
  const float ah = __uint_as_float(hi2 << 16), bh = __uint_as_float(hi2 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(b - bh), "f"(a - ah));
}
__device__ __forceinline__ float bf16lo_to_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// host-side bf16 round-to-nearest-even (weights are packed once at load time on the host)
static inline uint16_t host_f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);  // inf/nan passthrough
  uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}
static inline float host_bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// ------------------------------------------------------------------------------------------
// warp / block reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; result valid in thread 0. blockDim.x must be a multiple of 32 and <= 1024.
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float s_part[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? s_part[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

}  // namespace smb
