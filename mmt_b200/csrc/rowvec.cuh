// Warp-per-row register helpers shared by the HBM-bound kernels (rowops.cu, head.cu).
#pragma once
#include "common.cuh"

namespace mmt {

constexpr int WARPS = 8;

template <int VEC>
struct RowVec {
  float4 v[VEC];
};

template <int VEC>
__device__ __forceinline__ void load_row(const float* __restrict__ p, int lane, float4 (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = *reinterpret_cast<const float4*>(p + 4 * (lane + 32 * i));
}
template <int VEC>
__device__ __forceinline__ void store_row(float* __restrict__ p, int lane, const float4 (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) *reinterpret_cast<float4*>(p + 4 * (lane + 32 * i)) = v[i];
}
template <int VEC>
__device__ __forceinline__ float row_sum(const float4 (&v)[VEC]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  return warp_sum(s);
}
template <int VEC>
__device__ __forceinline__ float row_dot(const float4 (&a)[VEC], const float4 (&b)[VEC]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i)
    s += (a[i].x * b[i].x + a[i].y * b[i].y) + (a[i].z * b[i].z + a[i].w * b[i].w);
  return warp_sum(s);
}

#define F4_OP(dst, expr_x, expr_y, expr_z, expr_w) \
  do { (dst).x = (expr_x); (dst).y = (expr_y); (dst).z = (expr_z); (dst).w = (expr_w); } while (0)

__device__ __forceinline__ void atomic_add4(float* p, float4 v) {
  atomicAdd(reinterpret_cast<float4*>(p), v);   // red.global.add.v4.f32 (sm_90+)
}

// LayerNorm statistics of a register-resident row (two-pass, fp32).
template <int VEC>
__device__ __forceinline__ void ln_stats(const float4 (&e)[VEC], int d, float eps, float& mean,
                                         float& rstd) {
  mean = row_sum<VEC>(e) / d;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float a = e[i].x - mean, b = e[i].y - mean, c = e[i].z - mean, dd = e[i].w - mean;
    s += (a * a + b * b) + (c * c + dd * dd);
  }
  s = warp_sum(s) / d;
  rstd = 1.0f / sqrtf(s + eps);
}

// Block-level flush of per-lane column partials: acc[VEC] float4 per lane, summed over the
// block's warps, then atomically added to out[d].
template <int VEC>
__device__ __forceinline__ void flush_cols(float4 (&acc)[VEC], float* __restrict__ out, int lane,
                                           int warp, float4* smem /* [WARPS][32*VEC] */) {
  if (out == nullptr) return;
#pragma unroll
  for (int i = 0; i < VEC; ++i) smem[(warp * VEC + i) * 32 + lane] = acc[i];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 s = smem[i * 32 + lane];
      for (int w = 1; w < WARPS; ++w) {
        float4 t = smem[(w * VEC + i) * 32 + lane];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      atomic_add4(out + 4 * (lane + 32 * i), s);
    }
  }
  __syncthreads();
}


#define DISPATCH_VEC(d, CALL)                                                       \
  switch ((d) / 128) {                                                              \
    case 1: { constexpr int V = 1; CALL; } break;                                   \
    case 2: { constexpr int V = 2; CALL; } break;                                   \
    case 4: { constexpr int V = 4; CALL; } break;                                   \
    case 6: { constexpr int V = 6; CALL; } break;                                   \
    case 8: { constexpr int V = 8; CALL; } break;                                   \
    default: ::mmt::set_error("row width d=%d unsupported (need 128,256,512,768,1024)", (int)(d)); \
             return MMT_E_SHAPE;                                                    \
  }

#define CHECK_D(d) MMT_ARG_CHECK((d) % 128 == 0 && (d) <= 1024 && (d) > 0, MMT_E_SHAPE, "d=%d must be a multiple of 128 <= 1024", (int)(d))
#define CHECK_P(p) MMT_ARG_CHECK((p) >= 0.f && (p) < 1.f, MMT_E_ARG, "dropout p=%f out of [0,1)", (double)(p))

inline int row_grid(int64_t rows) {
  int64_t blocks = (rows + WARPS - 1) / WARPS;
  int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace mmt
