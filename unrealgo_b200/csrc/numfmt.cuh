// 16-bit operand formats of the tcgen05 path: bf16 (8-bit mantissa, fp32 range) or fp16 (11-bit mantissa,
// |x| <= 65504; results are clamped, never inf).  kind::f16 MMAs run at the same rate for both; the choice is a
// run-time flag so one set of kernels serves UG_PREC_BF16 and UG_PREC_FP16.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace ug {

__device__ __forceinline__ float2 unpack2(uint32_t w, int f16) {
  if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack2(float a, float b, int f16) {
  if (f16) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float unpack1(uint16_t w, int f16) {
  if (f16) return __half2float(__ushort_as_half(w));
  return __uint_as_float((uint32_t)w << 16);
}
__host__ __device__ constexpr uint32_t one16(int f16) { return f16 ? 0x3C00u : 0x3F80u; }

}  // namespace ug
