// Shared device/host helpers for the mmt_b200 sm_100a kernels.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmt_b200.h"

#include <atomic>

namespace mmt {

extern std::atomic<int64_t> g_launches;
// Optional device-resident step counter (mmt_set_step_counter): kernels add *ctr to their dropout
// seed / Adam step so a captured CUDA graph draws fresh masks on every replay.
extern const uint64_t* g_step_ctr;

// ---- error plumbing (no exceptions cross the C ABI) -----------------------------------------
void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);

#define MMT_ARG_CHECK(cond, code, ...)                 \
  do {                                                 \
    if (!(cond)) {                                     \
      ::mmt::set_error(__VA_ARGS__);                   \
      return (code);                                   \
    }                                                  \
  } while (0)

#define MMT_LAUNCH_CHECK(what)                                      \
  do {                                                              \
    ::mmt::g_launches.fetch_add(1, std::memory_order_relaxed);      \
    cudaError_t e__ = cudaGetLastError();                           \
    if (e__ != cudaSuccess) return ::mmt::cuda_status(e__, what);   \
  } while (0)

int num_sms();
int ensure_dynamic_smem(const void* fn, size_t bytes, const char* what);   // once per (kernel, device)

// ---- warp helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Philox4x32-10, stateless: (seed, 4x32 counter) -> 4x32 random bits ------------------------
// Dropout masks are a pure function of (seed, site, row, col/4) so that the backward kernels
// regenerate exactly the mask the forward used without storing it.
struct Philox {
  static __device__ __forceinline__ uint4 gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                              uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// Every kernel of this library starts with pdl_trigger() (the next kernel in the stream may begin
// launching: its CTAs take SMs as this grid's CTAs retire and run their prologue -- barrier init,
// TMEM allocation, descriptor prefetch) and calls pdl_wait() before its first access to global
// memory (returns once the preceding grid has completed and its writes are visible).  A train step
// is ~150 dependent launches; the launch / drain gap between two of them is a few microseconds.
// Both instructions are no-ops for a kernel launched without the attribute (MMT_PDL=0).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MMT_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// the same with a run-time cluster shape (cluster_x CTAs along x; 1 = no cluster attribute)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster_x;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cluster_x > 1 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// keep-mask scale for 4 consecutive columns [col4*4, col4*4+4) of `row` at dropout `site`.
// Returns 0 or 1/(1-p) per element.  p == 0 -> all ones (callers skip the call).
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint32_t site, uint32_t row,
                                                 uint32_t col4, float p, float inv_keep) {
  uint4 r = Philox::gen(seed, row, col4, site, 0x6d6d7462u);
  // keep iff u >= p with u uniform in [0,1): compare on the 32-bit integer
  uint32_t thr = (uint32_t)fminf(p * 4294967296.0f, 4294967295.0f);
  float4 o;
  o.x = r.x >= thr ? inv_keep : 0.f;
  o.y = r.y >= thr ? inv_keep : 0.f;
  o.z = r.z >= thr ? inv_keep : 0.f;
  o.w = r.w >= thr ? inv_keep : 0.f;
  return o;
}

// Attention-probability dropout (the S x S matrices: by far the most dropout decisions of a step) spends
// 16 random bits per element: one Philox call covers 8 consecutive key columns [col8*8, col8*8+8).
// P(drop) = floor(p * 65536) / 65536 (p = 0.1 -> 0.09999).  Element order: x.lo x.hi y.lo y.hi z.lo z.hi w.lo w.hi.
__device__ __forceinline__ void dropout_scale8_h16(uint64_t seed, uint32_t site, uint32_t row, uint32_t col8,
                                                   float p, float inv_keep, float (&o)[8]) {
  const uint4 r = Philox::gen(seed, row, col8, site, 0x6d6d7468u);
  const uint32_t thr = (uint32_t)(p * 65536.0f);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = (w[i] & 0xffffu) >= thr ? inv_keep : 0.f;
    o[2 * i + 1] = (w[i] >> 16) >= thr ? inv_keep : 0.f;
  }
}
// the same decisions for the 4 columns [col4*4, col4*4+4) (row-per-warp kernels, one float4 per lane)
__device__ __forceinline__ float4 dropout_scale4_h16(uint64_t seed, uint32_t site, uint32_t row, uint32_t col4,
                                                     float p, float inv_keep) {
  const uint4 r = Philox::gen(seed, row, col4 >> 1, site, 0x6d6d7468u);
  const uint32_t thr = (uint32_t)(p * 65536.0f);
  const uint32_t w0 = (col4 & 1) ? r.z : r.x, w1 = (col4 & 1) ? r.w : r.y;
  float4 o;
  o.x = (w0 & 0xffffu) >= thr ? inv_keep : 0.f;
  o.y = (w0 >> 16) >= thr ? inv_keep : 0.f;
  o.z = (w1 & 0xffffu) >= thr ? inv_keep : 0.f;
  o.w = (w1 >> 16) >= thr ? inv_keep : 0.f;
  return o;
}

// ---- fast dropout decisions of the 16-bit operand path ----------------------------------------------
// One 32-bit integer hash ("lowbias32" finaliser) per PAIR of decisions, 16 random bits each
// (P(drop) = floor(p * 65536) / 65536), keyed by (seed, site) and a per-element counter.  Dropout only needs
// a reproducible, well-mixed mask; Philox4x32-10 (the fp32 / tf32 path above) costs ~80 integer
// instructions per 4 decisions, which made the GEMM epilogues and the attention softmax ALU-bound.
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_key(uint64_t seed, uint32_t site) {
  return hash32((uint32_t)seed ^ hash32((uint32_t)(seed >> 32) + site * 0x9E3779B9u + 0x6d6d7461u));
}
// keep-mask scale of the 4 columns [col4*4, col4*4+4) of `row` (rows < 2^21, columns < 4096)
__device__ __forceinline__ float4 dropout_scale4_fast(uint32_t key32, uint32_t row, uint32_t col4, uint32_t thr16,
                                                      float inv_keep) {
  const uint32_t idx = row * 2048u + col4 * 2u;
  const uint32_t w0 = hash32(idx ^ key32), w1 = hash32((idx + 1u) ^ key32);
  float4 o;
  o.x = (w0 & 0xffffu) >= thr16 ? inv_keep : 0.f;
  o.y = (w0 >> 16) >= thr16 ? inv_keep : 0.f;
  o.z = (w1 & 0xffffu) >= thr16 ? inv_keep : 0.f;
  o.w = (w1 >> 16) >= thr16 ? inv_keep : 0.f;
  return o;
}

__device__ __forceinline__ float gelu_erf(float x) {
  // model/bert.py:53: x * 0.5 * (1 + erf(x / sqrt(2)))
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float dgelu_erf(float x) {
  // d/dx [x Phi(x)] = Phi(x) + x phi(x)
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Branch-free erf via Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7), one MUFU.EX2 + one MUFU.RCP.
// The tensor-core epilogues use it: with only four epilogue warps per SM the long dependent chain
// (and the divergent range split) of libdevice's erff made GELU the bottleneck of the FFN GEMMs.
// e = exp(-u^2/2) is shared between erf(u/sqrt2) and the Gaussian pdf needed by GELU'.
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));      // one MUFU.RCP, 1 ulp; no slow-path call
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));      // one MUFU.EX2, 2 ulp
  return r;
}
__device__ __forceinline__ void erf_parts(float u, float& erf_v, float& e) {
  const float x = u * 0.70710678118654752440f;
  const float ax = fabsf(x);
  const float t = fast_rcp(fmaf(0.3275911f, ax, 1.0f));
  e = fast_ex2(-ax * ax * 1.44269504088896340736f);
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  erf_v = copysignf(fmaf(-p * t, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_fast(float u) {
  float er, e;
  erf_parts(u, er, e);
  return 0.5f * u * (1.0f + er);
}
// gelu(u) and gelu'(u) from one erf / exp evaluation (FFN-up epilogue: the derivative is stored for the backward)
__device__ __forceinline__ void gelu_both(float u, float& g, float& dg) {
  float er, e;
  erf_parts(u, er, e);
  const float cdf = fmaf(0.5f, er, 0.5f);
  g = u * cdf;
  dg = fmaf(u * 0.39894228040143267794f, e, cdf);
}
__device__ __forceinline__ float dgelu_fast(float u) {
  float er, e;
  erf_parts(u, er, e);
  return fmaf(u * 0.39894228040143267794f, e, 0.5f * (1.0f + er));
}

}  // namespace mmt
