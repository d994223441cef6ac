// Split-fp16 representation of fp32 values used by the precise tensor-core mode (precision 2, kernels_gemm_x3.cu):
//     x  ~=  hi + lo * 2^-11,   hi = fp16(x),   lo = fp16((x - hi) * 2^11)
// x - hi is exact in fp32, so |x - (hi + lo 2^-11)| <= 2^-22 |x| for |x| >= 2^-14 (below that the absolute error is
// bounded by ~2^-36).  An activation matrix [M][K] that feeds a Linear is stored as two K-major fp16 planes (hi, lo)
// -- the same 4 bytes per element as fp32 -- by the kernel that produces it, so the GEMM's operands arrive by TMA.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace scb {

constexpr float X3_SCALE = 2048.0f;          // 2^11
constexpr float X3_INV_SCALE = 1.0f / 2048.0f;

// saturating instead of producing inf (|x| > 65504 cannot occur for LayerNorm / ReLU / attention outputs of this model;
// saturation keeps a stray value finite)
__device__ __forceinline__ void x3_split(float x, __half& hi, __half& lo) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  hi = __ushort_as_half(h);
  const float r = (x - __half2float(hi)) * X3_SCALE;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(r));
  lo = __ushort_as_half(h);
}

// eight consecutive values -> one 16-byte chunk of the hi plane and one of the lo plane
__device__ __forceinline__ void x3_split8(const float* x, uint4& uh, uint4& ul) {
  __half hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x3_split(x[j], hi[j], lo[j]);
  uh.x = (uint32_t)__half_as_ushort(hi[0]) | ((uint32_t)__half_as_ushort(hi[1]) << 16);
  uh.y = (uint32_t)__half_as_ushort(hi[2]) | ((uint32_t)__half_as_ushort(hi[3]) << 16);
  uh.z = (uint32_t)__half_as_ushort(hi[4]) | ((uint32_t)__half_as_ushort(hi[5]) << 16);
  uh.w = (uint32_t)__half_as_ushort(hi[6]) | ((uint32_t)__half_as_ushort(hi[7]) << 16);
  ul.x = (uint32_t)__half_as_ushort(lo[0]) | ((uint32_t)__half_as_ushort(lo[1]) << 16);
  ul.y = (uint32_t)__half_as_ushort(lo[2]) | ((uint32_t)__half_as_ushort(lo[3]) << 16);
  ul.z = (uint32_t)__half_as_ushort(lo[4]) | ((uint32_t)__half_as_ushort(lo[5]) << 16);
  ul.w = (uint32_t)__half_as_ushort(lo[6]) | ((uint32_t)__half_as_ushort(lo[7]) << 16);
}

// Destination of a kernel that emits split planes: element (row, col) of plane p lives at base[p * plane + row * ld + col]
struct SplitOut {
  __half* base = nullptr;
  size_t plane = 0;      // elements between the hi and the lo plane
  int ld = 0;
  __device__ __forceinline__ void put(size_t row, int col, float v) const {
    __half h, l;
    x3_split(v, h, l);
    base[row * ld + col] = h;
    base[plane + row * ld + col] = l;
  }
};

}  // namespace scb
