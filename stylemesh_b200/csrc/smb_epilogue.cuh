// Shared accumulator epilogue (see struct Epilogue in smb_kernels.h for the semantics).
#pragma once
#include "smb_common.cuh"
#include "smb_kernels.h"

namespace smb {

// The arithmetic half of the epilogue on NV (multiple of 4) consecutive channels n..n+NV-1 of pixel p:
//   v *= rowscale[p]; v += bias[n]; v += addend[p][n]; v = sign_hi[p][n] > 0 ? v : 0; v = relu ? max(v, 0) : v
// `in_image` = p is a real pixel (per-pixel operands are only read then; bias is per channel and always applied).
template <int NV>
__device__ __forceinline__ void epilogue_apply(const Epilogue& ep, int64_t p, int n, int N, float (&v)[NV],
                                               bool in_image = true) {
  static_assert(NV % 4 == 0, "NV must be a multiple of 4");
  const int64_t off = p * (int64_t)N + n;
  if (ep.rowscale && in_image) {
    const float rs = __ldg(ep.rowscale + p);
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] *= rs;
  }
  if (ep.bias) {
#pragma unroll
    for (int j = 0; j < NV; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (ep.addend && in_image) {
#pragma unroll
    for (int j = 0; j < NV; j += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ep.addend + off + j));
      v[j] += a.x; v[j + 1] += a.y; v[j + 2] += a.z; v[j + 3] += a.w;
    }
  }
  if (ep.sign_hi && in_image) {
#pragma unroll
    for (int j = 0; j < NV; j += 4) {
      const uint2 s = __ldg(reinterpret_cast<const uint2*>(ep.sign_hi + off + j));
      // bf16 > 0  <=>  sign bit clear and magnitude non-zero
      const uint32_t e0 = s.x & 0xffffu, e1 = s.x >> 16, e2 = s.y & 0xffffu, e3 = s.y >> 16;
      if (!((e0 & 0x8000u) == 0 && (e0 & 0x7fffu) != 0)) v[j] = 0.f;
      if (!((e1 & 0x8000u) == 0 && (e1 & 0x7fffu) != 0)) v[j + 1] = 0.f;
      if (!((e2 & 0x8000u) == 0 && (e2 & 0x7fffu) != 0)) v[j + 2] = 0.f;
      if (!((e3 & 0x8000u) == 0 && (e3 & 0x7fffu) != 0)) v[j + 3] = 0.f;
    }
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

// Apply the epilogue to NV (multiple of 4) consecutive channels n..n+NV-1 of pixel p and store.
// All row pointers are [P][N] with N a multiple of 4 and n a multiple of 4 => 16-byte fp32 / 8-byte bf16 accesses
// (NV a multiple of 8 and n a multiple of 8 upgrades bf16 accesses to 16 bytes).
template <int NV>
__device__ __forceinline__ void epilogue_store(const Epilogue& ep, int64_t p, int n, int N, float (&v)[NV]) {
  const int64_t off = p * (int64_t)N + n;
  epilogue_apply<NV>(ep, p, n, N, v);
  if (ep.out_f32) {
#pragma unroll
    for (int j = 0; j < NV; j += 4)
      *reinterpret_cast<float4*>(ep.out_f32 + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  if (ep.out_hi) {
    if constexpr (NV % 8 == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 8) {
        uint4 h, l;
        split2_pack(v[j], v[j + 1], h.x, l.x);
        split2_pack(v[j + 2], v[j + 3], h.y, l.y);
        split2_pack(v[j + 4], v[j + 5], h.z, l.z);
        split2_pack(v[j + 6], v[j + 7], h.w, l.w);
        *reinterpret_cast<uint4*>(ep.out_hi + off + j) = h;
        *reinterpret_cast<uint4*>(ep.out_lo + off + j) = l;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        uint2 h, l;
        split2_pack(v[j], v[j + 1], h.x, l.x);
        split2_pack(v[j + 2], v[j + 3], h.y, l.y);
        *reinterpret_cast<uint2*>(ep.out_hi + off + j) = h;
        *reinterpret_cast<uint2*>(ep.out_lo + off + j) = l;
      }
    }
  }
  if (ep.outm_hi) {
    const float rm = __ldg(ep.rowmask + p);
    if constexpr (NV % 8 == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 8) {
        uint4 h, l;
        split2_pack(v[j] * rm, v[j + 1] * rm, h.x, l.x);
        split2_pack(v[j + 2] * rm, v[j + 3] * rm, h.y, l.y);
        split2_pack(v[j + 4] * rm, v[j + 5] * rm, h.z, l.z);
        split2_pack(v[j + 6] * rm, v[j + 7] * rm, h.w, l.w);
        *reinterpret_cast<uint4*>(ep.outm_hi + off + j) = h;
        *reinterpret_cast<uint4*>(ep.outm_lo + off + j) = l;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {
        uint2 h, l;
        split2_pack(v[j] * rm, v[j + 1] * rm, h.x, l.x);
        split2_pack(v[j + 2] * rm, v[j + 3] * rm, h.y, l.y);
        *reinterpret_cast<uint2*>(ep.outm_hi + off + j) = h;
        *reinterpret_cast<uint2*>(ep.outm_lo + off + j) = l;
      }
    }
  }
}

}  // namespace smb
