// 16-bit operand formats of the tensor-core path: fp16 (10-bit mantissa = tf32's, half the bytes; the
// fp32-tolerance configurations) and bf16 (BASELINE config 5).  Every conversion is round-to-nearest-even
// and saturates to the largest finite value, so a gradient outlier can never become inf.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace mmt {

// low 16 bits = first element, high 16 bits = second element (memory order of two consecutive elements)
__device__ __forceinline__ uint32_t pack2(float lo, float hi, bool bf16) {
  uint32_t r;
  if (bf16) asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack2(uint32_t v, bool bf16) {
  if (bf16) return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ uint2 pack4(float4 v, bool bf16) {
  return make_uint2(pack2(v.x, v.y, bf16), pack2(v.z, v.w, bf16));
}
__device__ __forceinline__ float4 unpack4(uint2 v, bool bf16) {
  const float2 a = unpack2(v.x, bf16), b = unpack2(v.y, bf16);
  return make_float4(a.x, a.y, b.x, b.y);
}
// one element
__device__ __forceinline__ uint16_t pack1(float x, bool bf16) { return (uint16_t)(pack2(x, 0.f, bf16) & 0xffffu); }
__device__ __forceinline__ float unpack1(uint16_t x, bool bf16) { return unpack2((uint32_t)x, bf16).x; }

// warp-per-row layout of rowvec.cuh (lane l owns columns 4*(l + 32*i)): 16-bit copies of a row
template <int VEC>
__device__ __forceinline__ void store_row16(void* __restrict__ p, int lane, const float4 (&v)[VEC], float scale, bool bf16) {
  uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
  for (int i = 0; i < VEC; ++i)
    q[lane + 32 * i] = pack4(make_float4(v[i].x * scale, v[i].y * scale, v[i].z * scale, v[i].w * scale), bf16);
}
template <int VEC>
__device__ __forceinline__ void load_row16(const void* __restrict__ p, int lane, float4 (&v)[VEC], bool bf16) {
  const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = unpack4(q[lane + 32 * i], bf16);
}

}  // namespace mmt
