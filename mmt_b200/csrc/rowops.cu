// HBM-bound row kernels of the video encoder: token assembly + BertEmbeddings, residual +
// LayerNorm, masked softmax, expert read-out, bias-gradient column sums, dropout.
//
// Layout: one warp per row of d = 128*VEC floats; lane l owns the float4s at columns
// 4*(l + 32*v), v < VEC, so every warp-wide access is a fully coalesced 512 B segment and all
// row reductions are warp shuffles.  Column reductions (dgamma/dbeta/dbias) are kept in
// registers across a grid-stride loop over rows, reduced across the block's warps through shared
// memory and flushed with one vector atomic per lane per block.
#include "cvt16.cuh"
#include "rowvec.cuh"

namespace mmt {
namespace {

// ------------------------------------------------------------------------------------------
// Token assembly + BertEmbeddings forward (model/model.py:485-567, model/bert.py:87-105)
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) embed_ln_fwd_kernel(
    const float* __restrict__ proj, const float* __restrict__ ft, const float* __restrict__ ind,
    const int32_t* __restrict__ type_idx, const float* __restrict__ pos_emb,
    const float* __restrict__ type_emb, const float* __restrict__ gamma,
    const float* __restrict__ beta, int B, int M, int T, int max_pos, float eps, float p_drop,
    uint64_t seed, uint32_t site, const uint64_t* __restrict__ ctr, float* __restrict__ h, float* __restrict__ mask,
    int32_t* __restrict__ pos_ids, int32_t* __restrict__ type_ids, float* __restrict__ inv_norm,
    float* __restrict__ mean_o, float* __restrict__ rstd_o, void* __restrict__ h16, int bf16) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;        // device-side step counter (CUDA-graph replays)
  constexpr int d = 128 * VEC;
  const int S = 1 + M * (T + 1);
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)B * S;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float4 g[VEC], bt[VEC];
  load_row<VEC>(gamma, lane, g);
  load_row<VEC>(beta, lane, bt);
  for (int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows;
       r += (int64_t)gridDim.x * WARPS) {
    const int b = (int)(r / S), s = (int)(r % S);
    int type = 0, pos = 0;
    float mk = 1.f, invn = 0.f;
    float4 e[VEC];
    if (s == 0) {                                        // [CLS]: model.py:496-504
#pragma unroll
      for (int i = 0; i < VEC; ++i) e[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const int k = (s - 1) / (T + 1), j = (s - 1) % (T + 1);
      type = type_idx[k];
      const float* indp = ind + ((int64_t)k * B + b) * T;
      if (j == 0) {                                      // [AGG]: model.py:526-541, 329-330
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, indp[t]);
        mx = warp_max(mx);
        mk = (float)(long long)mx;
      } else {                                           // temporal token: model.py:543-558
        float tv = ft[((int64_t)k * B + b) * T + (j - 1)];
        tv = fminf(fmaxf(tv, 0.f), (float)(max_pos - 1));   // clamp_ (model.py:516-518)
        pos = (int)tv;                                       // .long(): truncation
        mk = (float)(long long)indp[j - 1];
      }
      load_row<VEC>(proj + (((int64_t)k * B + b) * (T + 1) + j) * d, lane, e);   // [M, B, T+1, d]
      float ss = row_dot<VEC>(e, e);
      invn = 1.0f / fmaxf(sqrtf(ss), 1e-12f);            // F.normalize eps (model.py:725)
#pragma unroll
      for (int i = 0; i < VEC; ++i) F4_OP(e[i], e[i].x * invn, e[i].y * invn, e[i].z * invn, e[i].w * invn);
    }
    float4 pe[VEC], te[VEC];
    load_row<VEC>(pos_emb + (int64_t)pos * d, lane, pe);
    load_row<VEC>(type_emb + (int64_t)type * d, lane, te);
#pragma unroll
    for (int i = 0; i < VEC; ++i)                        // bert.py:99: pos + type + features
      F4_OP(e[i], (pe[i].x + te[i].x) + e[i].x, (pe[i].y + te[i].y) + e[i].y,
            (pe[i].z + te[i].z) + e[i].z, (pe[i].w + te[i].w) + e[i].w);
    float mean, rstd;
    ln_stats<VEC>(e, d, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      F4_OP(e[i], (e[i].x - mean) * rstd * g[i].x + bt[i].x, (e[i].y - mean) * rstd * g[i].y + bt[i].y,
            (e[i].z - mean) * rstd * g[i].z + bt[i].z, (e[i].w - mean) * rstd * g[i].w + bt[i].w);
      if (p_drop > 0.f) {
        float4 sc = dropout_scale4(seed, site, (uint32_t)r, lane + 32 * i, p_drop, inv_keep);
        F4_OP(e[i], e[i].x * sc.x, e[i].y * sc.y, e[i].z * sc.z, e[i].w * sc.w);
      }
    }
    store_row<VEC>(h + r * d, lane, e);
    if (h16 != nullptr) store_row16<VEC>(reinterpret_cast<uint16_t*>(h16) + r * d, lane, e, 1.0f, bf16 != 0);
    if (lane == 0) {
      mask[r] = mk; pos_ids[r] = pos; type_ids[r] = type; inv_norm[r] = invn;
      mean_o[r] = mean; rstd_o[r] = rstd;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) embed_ln_bwd_kernel(
    const float* __restrict__ dh, const float* __restrict__ proj,
    const int32_t* __restrict__ pos_ids, const int32_t* __restrict__ type_ids,
    const float* __restrict__ inv_norm, const float* __restrict__ mean_i,
    const float* __restrict__ rstd_i, const float* __restrict__ pos_emb,
    const float* __restrict__ type_emb, const float* __restrict__ gamma, int B, int M, int T,
    float p_drop, uint64_t seed, uint32_t site, const uint64_t* __restrict__ ctr, float* __restrict__ dproj,
    float* __restrict__ dpos_emb, float* __restrict__ dtype_emb, float* __restrict__ dgamma,
    float* __restrict__ dbeta, void* __restrict__ dproj16, float scale16, int bf16) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;        // device-side step counter (CUDA-graph replays)
  constexpr int d = 128 * VEC;
  __shared__ float4 red[WARPS * VEC * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = 1 + M * (T + 1);
  const int64_t rows = (int64_t)B * S;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float4 g[VEC], ag[VEC], ab[VEC];
  load_row<VEC>(gamma, lane, g);
#pragma unroll
  for (int i = 0; i < VEC; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // Rows are walked position-major (j = s * B + b) in one contiguous chunk per warp: consecutive rows then share the token
  // type and -- with regular frame sampling -- the position id, so their embedding gradients are summed in registers and
  // flushed once per run.  (One pair of 512-float atomic rows per TOKEN was 3.6 M float4 atomics on 51 table rows.)
  const int64_t warps_total = (int64_t)gridDim.x * WARPS;
  const int64_t chunk = (rows + warps_total - 1) / warps_total;
  const int64_t j0 = ((int64_t)blockIdx.x * WARPS + warp) * chunk;
  const int64_t j1 = j0 + chunk < rows ? j0 + chunk : rows;
  int cur_pos = -1, cur_type = -1;
  float4 accp[VEC], acct[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) accp[i] = acct[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t j = j0; j < j1; ++j) {
    const int s = (int)(j / B), b = (int)(j % B);
    const int64_t r = (int64_t)b * S + s;
    // expert-major row of the projection buffers [M, B, T+1, d] (CLS has none)
    const int64_t prow = s == 0 ? 0 : (((int64_t)((s - 1) / (T + 1)) * B + b) * (T + 1) + (s - 1) % (T + 1));
    const int pos = pos_ids[r], type = type_ids[r];
    const float invn = inv_norm[r], mean = mean_i[r], rstd = rstd_i[r];
    float4 gy[VEC], f[VEC], xh[VEC], pe[VEC], te[VEC];
    load_row<VEC>(dh + r * d, lane, gy);
    if (p_drop > 0.f) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float4 sc = dropout_scale4(seed, site, (uint32_t)r, lane + 32 * i, p_drop, inv_keep);
        F4_OP(gy[i], gy[i].x * sc.x, gy[i].y * sc.y, gy[i].z * sc.z, gy[i].w * sc.w);
      }
    }
    if (s == 0) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      load_row<VEC>(proj + prow * d, lane, f);
#pragma unroll
      for (int i = 0; i < VEC; ++i) F4_OP(f[i], f[i].x * invn, f[i].y * invn, f[i].z * invn, f[i].w * invn);
    }
    load_row<VEC>(pos_emb + (int64_t)pos * d, lane, pe);
    load_row<VEC>(type_emb + (int64_t)type * d, lane, te);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      F4_OP(xh[i], (((pe[i].x + te[i].x) + f[i].x) - mean) * rstd,
            (((pe[i].y + te[i].y) + f[i].y) - mean) * rstd,
            (((pe[i].z + te[i].z) + f[i].z) - mean) * rstd,
            (((pe[i].w + te[i].w) + f[i].w) - mean) * rstd);
      ag[i].x += gy[i].x * xh[i].x; ag[i].y += gy[i].y * xh[i].y;
      ag[i].z += gy[i].z * xh[i].z; ag[i].w += gy[i].w * xh[i].w;
      ab[i].x += gy[i].x; ab[i].y += gy[i].y; ab[i].z += gy[i].z; ab[i].w += gy[i].w;
      F4_OP(gy[i], gy[i].x * g[i].x, gy[i].y * g[i].y, gy[i].z * g[i].z, gy[i].w * g[i].w);  // d xhat
    }
    const float m1 = row_sum<VEC>(gy) / d;
    const float m2 = row_dot<VEC>(gy, xh) / d;
    float4 de[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      F4_OP(de[i], rstd * (gy[i].x - m1 - xh[i].x * m2), rstd * (gy[i].y - m1 - xh[i].y * m2),
            rstd * (gy[i].z - m1 - xh[i].z * m2), rstd * (gy[i].w - m1 - xh[i].w * m2));
    if (pos != cur_pos) {                                // warp-uniform
      if (cur_pos >= 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) atomic_add4(dpos_emb + (int64_t)cur_pos * d + 4 * (lane + 32 * i), accp[i]);
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) accp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      cur_pos = pos;
    }
    if (type != cur_type) {
      if (cur_type >= 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) atomic_add4(dtype_emb + (int64_t)cur_type * d + 4 * (lane + 32 * i), acct[i]);
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) acct[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      cur_type = type;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      accp[i].x += de[i].x; accp[i].y += de[i].y; accp[i].z += de[i].z; accp[i].w += de[i].w;
      acct[i].x += de[i].x; acct[i].y += de[i].y; acct[i].z += de[i].z; acct[i].w += de[i].w;
    }
    if (s == 0) continue;                                // [CLS] carries no projected feature
    if (invn < 1e12f) {                           // normalize backward: (I - f f^T) de / ||y||
      const float dot = row_dot<VEC>(f, de);
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        F4_OP(de[i], invn * (de[i].x - f[i].x * dot), invn * (de[i].y - f[i].y * dot),
              invn * (de[i].z - f[i].z * dot), invn * (de[i].w - f[i].w * dot));
    } else {                                             // ||y|| <= eps: y / eps is linear
#pragma unroll
      for (int i = 0; i < VEC; ++i) F4_OP(de[i], de[i].x * invn, de[i].y * invn, de[i].z * invn, de[i].w * invn);
    }
    store_row<VEC>(dproj + prow * d, lane, de);
    if (dproj16 != nullptr) store_row16<VEC>(reinterpret_cast<uint16_t*>(dproj16) + prow * d, lane, de, scale16, bf16 != 0);
  }
  if (cur_pos >= 0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) atomic_add4(dpos_emb + (int64_t)cur_pos * d + 4 * (lane + 32 * i), accp[i]);
  }
  if (cur_type >= 0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) atomic_add4(dtype_emb + (int64_t)cur_type * d + 4 * (lane + 32 * i), acct[i]);
  }
  flush_cols<VEC>(ag, dgamma, lane, warp, red);
  flush_cols<VEC>(ab, dbeta, lane, warp, red);
}

// ------------------------------------------------------------------------------------------
// y = LN(dropout(t) + r)   (model/bert.py:186-188, 234-236)
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) res_ln_fwd_kernel(
    float* __restrict__ t, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, int64_t rows, float eps, float p_drop, uint64_t seed,
    uint32_t site, const uint64_t* __restrict__ ctr, float* __restrict__ y, float* __restrict__ mean_o, float* __restrict__ rstd_o) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;        // device-side step counter (CUDA-graph replays)
  constexpr int d = 128 * VEC;
  const int lane = threadIdx.x & 31;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float4 g[VEC], bt[VEC];
  load_row<VEC>(gamma, lane, g);
  load_row<VEC>(beta, lane, bt);
  for (int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows;
       r += (int64_t)gridDim.x * WARPS) {
    float4 z[VEC], rr[VEC];
    load_row<VEC>(t + r * d, lane, z);
    load_row<VEC>(res + r * d, lane, rr);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (p_drop > 0.f) {
        float4 sc = dropout_scale4(seed, site, (uint32_t)r, lane + 32 * i, p_drop, inv_keep);
        F4_OP(z[i], z[i].x * sc.x, z[i].y * sc.y, z[i].z * sc.z, z[i].w * sc.w);
      }
      F4_OP(z[i], z[i].x + rr[i].x, z[i].y + rr[i].y, z[i].z + rr[i].z, z[i].w + rr[i].w);
    }
    store_row<VEC>(t + r * d, lane, z);
    float mean, rstd;
    ln_stats<VEC>(z, d, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      F4_OP(z[i], (z[i].x - mean) * rstd * g[i].x + bt[i].x, (z[i].y - mean) * rstd * g[i].y + bt[i].y,
            (z[i].z - mean) * rstd * g[i].z + bt[i].z, (z[i].w - mean) * rstd * g[i].w + bt[i].w);
    store_row<VEC>(y + r * d, lane, z);
    if (lane == 0) { mean_o[r] = mean; rstd_o[r] = rstd; }
  }
}

template <int VEC>
__global__ void __launch_bounds__(WARPS * 32, VEC <= 4 ? 2 : 1) res_ln_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ dy2, const float* __restrict__ z,
    const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
    const float* __restrict__ gamma, int64_t rows, float p_drop, uint64_t seed, uint32_t site, const uint64_t* __restrict__ ctr,
    float* __restrict__ dz, float* __restrict__ dt, float* __restrict__ dgamma,
    float* __restrict__ dbeta, float* __restrict__ dbias, void* __restrict__ dt16, float scale16, int bf16) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;        // device-side step counter (CUDA-graph replays)
  const bool fast = dt16 != nullptr;       // 16-bit path: the hash mask the GEMM epilogue applied
  const uint32_t key32 = drop_key(seed, site), thr16 = (uint32_t)(p_drop * 65536.0f);
  constexpr int d = 128 * VEC;
  __shared__ float4 red[WARPS * VEC * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float4 g[VEC], ag[VEC], ab[VEC], abias[VEC];
  load_row<VEC>(gamma, lane, g);
#pragma unroll
  for (int i = 0; i < VEC; ++i) ag[i] = ab[i] = abias[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < rows; r += (int64_t)gridDim.x * WARPS) {
    const float mean = mean_i[r], rstd = rstd_i[r];
    float4 gy[VEC], xh[VEC];
    load_row<VEC>(dy + r * d, lane, gy);
    if (dy2 != nullptr) {
      float4 g2[VEC];
      load_row<VEC>(dy2 + r * d, lane, g2);
#pragma unroll
      for (int i = 0; i < VEC; ++i) F4_OP(gy[i], gy[i].x + g2[i].x, gy[i].y + g2[i].y, gy[i].z + g2[i].z, gy[i].w + g2[i].w);
    }
    load_row<VEC>(z + r * d, lane, xh);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      F4_OP(xh[i], (xh[i].x - mean) * rstd, (xh[i].y - mean) * rstd, (xh[i].z - mean) * rstd, (xh[i].w - mean) * rstd);
      ag[i].x += gy[i].x * xh[i].x; ag[i].y += gy[i].y * xh[i].y;
      ag[i].z += gy[i].z * xh[i].z; ag[i].w += gy[i].w * xh[i].w;
      ab[i].x += gy[i].x; ab[i].y += gy[i].y; ab[i].z += gy[i].z; ab[i].w += gy[i].w;
      F4_OP(gy[i], gy[i].x * g[i].x, gy[i].y * g[i].y, gy[i].z * g[i].z, gy[i].w * g[i].w);
    }
    const float m1 = row_sum<VEC>(gy) / d;
    const float m2 = row_dot<VEC>(gy, xh) / d;
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      F4_OP(gy[i], rstd * (gy[i].x - m1 - xh[i].x * m2), rstd * (gy[i].y - m1 - xh[i].y * m2),
            rstd * (gy[i].z - m1 - xh[i].z * m2), rstd * (gy[i].w - m1 - xh[i].w * m2));
    store_row<VEC>(dz + r * d, lane, gy);
    if (p_drop > 0.f) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float4 sc = fast ? dropout_scale4_fast(key32, (uint32_t)r, lane + 32 * i, thr16, inv_keep)
                         : dropout_scale4(seed, site, (uint32_t)r, lane + 32 * i, p_drop, inv_keep);
        F4_OP(gy[i], gy[i].x * sc.x, gy[i].y * sc.y, gy[i].z * sc.z, gy[i].w * sc.w);
      }
      if (dt != nullptr) store_row<VEC>(dt + r * d, lane, gy);
    }
    if (dt16 != nullptr) store_row16<VEC>(reinterpret_cast<uint16_t*>(dt16) + r * d, lane, gy, scale16, bf16 != 0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      abias[i].x += gy[i].x; abias[i].y += gy[i].y; abias[i].z += gy[i].z; abias[i].w += gy[i].w;
    }
  }
  flush_cols<VEC>(ag, dgamma, lane, warp, red);
  flush_cols<VEC>(ab, dbeta, lane, warp, red);
  flush_cols<VEC>(abias, dbias, lane, warp, red);
}

// ------------------------------------------------------------------------------------------
// masked softmax over materialised scores (model/bert.py:147-164), one warp per (b,h,i) row
// ------------------------------------------------------------------------------------------
constexpr int SM_MAXC = 4;   // float4 chunks per lane -> S <= 512

__global__ void __launch_bounds__(WARPS * 32) softmax_fwd_kernel(
    const float* __restrict__ scores, const float* __restrict__ mask, int B, int H, int S, int ld,
    float scale, float p_drop, uint64_t seed, uint32_t site, const uint64_t* __restrict__ ctr, float* __restrict__ Psoft,
    float* __restrict__ Pdrop) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;        // device-side step counter (CUDA-graph replays)
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)B * H * S;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows;
       r += (int64_t)gridDim.x * WARPS) {
    const int b = (int)(r / ((int64_t)H * S));
    const float* __restrict__ mrow = mask + (int64_t)b * S;
    float4 x[SM_MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < SM_MAXC; ++c) {
      const int j = 4 * (lane + 32 * c);
      if (j < ld) {
        float4 v = *reinterpret_cast<const float4*>(scores + r * ld + j);
        float* pv = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // bert.py:149-152: scores / sqrt(dh) + (1 - mask) * -10000
          pv[q] = (j + q < S) ? pv[q] * scale + (1.0f - mrow[j + q]) * -10000.0f : -INFINITY;
          mx = fmaxf(mx, pv[q]);
        }
        x[c] = v;
      } else {
        x[c] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < SM_MAXC; ++c) {
      F4_OP(x[c], expf(x[c].x - mx), expf(x[c].y - mx), expf(x[c].z - mx), expf(x[c].w - mx));
      sum += (x[c].x + x[c].y) + (x[c].z + x[c].w);
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c = 0; c < SM_MAXC; ++c) {
      const int j = 4 * (lane + 32 * c);
      if (j < ld) {
        float4 p;
        F4_OP(p, x[c].x * inv, x[c].y * inv, x[c].z * inv, x[c].w * inv);
        *reinterpret_cast<float4*>(Psoft + r * ld + j) = p;
        if (p_drop > 0.f) {
          float4 sc = dropout_scale4_h16(seed, site, (uint32_t)r, lane + 32 * c, p_drop, inv_keep);
          F4_OP(p, p.x * sc.x, p.y * sc.y, p.z * sc.z, p.w * sc.w);
          *reinterpret_cast<float4*>(Pdrop + r * ld + j) = p;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(WARPS * 32) softmax_bwd_kernel(
    float* __restrict__ dP, const float* __restrict__ Psoft, int64_t rows, int S, int ld,
    float scale, float p_drop, uint64_t seed, uint32_t site, const uint64_t* __restrict__ ctr) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;        // device-side step counter (CUDA-graph replays)
  const int lane = threadIdx.x & 31;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  // each lane owns 8 consecutive columns per chunk: one Philox call = exactly its 8 dropout decisions
  constexpr int NCH = SM_MAXC / 2;
  for (int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows;
       r += (int64_t)gridDim.x * WARPS) {
    float da[NCH][8], a[NCH][8];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = 8 * (lane + 32 * c);
#pragma unroll
      for (int q = 0; q < 8; ++q) { da[c][q] = 0.f; a[c][q] = 0.f; }
      if (j < ld) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (j + 4 * h < ld) {
            const float4 d4 = *reinterpret_cast<const float4*>(dP + r * ld + j + 4 * h);
            const float4 a4 = *reinterpret_cast<const float4*>(Psoft + r * ld + j + 4 * h);
            da[c][4 * h] = d4.x; da[c][4 * h + 1] = d4.y; da[c][4 * h + 2] = d4.z; da[c][4 * h + 3] = d4.w;
            a[c][4 * h] = a4.x; a[c][4 * h + 1] = a4.y; a[c][4 * h + 2] = a4.z; a[c][4 * h + 3] = a4.w;
          }
        }
        if (p_drop > 0.f) {
          float sc[8];
          dropout_scale8_h16(seed, site, (uint32_t)r, (uint32_t)(j >> 3), p_drop, inv_keep, sc);
#pragma unroll
          for (int q = 0; q < 8; ++q) da[c][q] *= sc[q];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (j + q >= S) { a[c][q] = 0.f; da[c][q] = 0.f; }
          dot = fmaf(a[c][q], da[c][q], dot);
        }
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = 8 * (lane + 32 * c);
      if (j < ld) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (j + 4 * h < ld) {
            float4 o;
            o.x = a[c][4 * h] * (da[c][4 * h] - dot) * scale;
            o.y = a[c][4 * h + 1] * (da[c][4 * h + 1] - dot) * scale;
            o.z = a[c][4 * h + 2] * (da[c][4 * h + 2] - dot) * scale;
            o.w = a[c][4 * h + 3] * (da[c][4 * h + 3] - dot) * scale;
            *reinterpret_cast<float4*>(dP + r * ld + j + 4 * h) = o;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// expert read-out + L2 normalise (model/model.py:583-587, 621-623)
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) readout_fwd_kernel(
    const float* __restrict__ h, int B, int S, int M, int T, float* __restrict__ v,
    float* __restrict__ inv_norm) {
  pdl_trigger();
  pdl_wait();
  constexpr int d = 128 * VEC;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= B * M) return;
  const int b = r / M, k = r % M;
  float4 x[VEC];
  load_row<VEC>(h + ((int64_t)b * S + 1 + (int64_t)k * (T + 1)) * d, lane, x);
  const float invn = 1.0f / fmaxf(sqrtf(row_dot<VEC>(x, x)), 1e-12f);
#pragma unroll
  for (int i = 0; i < VEC; ++i) F4_OP(x[i], x[i].x * invn, x[i].y * invn, x[i].z * invn, x[i].w * invn);
  store_row<VEC>(v + (int64_t)r * d, lane, x);
  if (lane == 0) inv_norm[r] = invn;
}

template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) readout_bwd_kernel(
    const float* __restrict__ dv, const float* __restrict__ v, const float* __restrict__ inv_norm,
    int B, int S, int M, int T, float* __restrict__ dh) {
  pdl_trigger();
  pdl_wait();
  constexpr int d = 128 * VEC;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= B * M) return;
  const int b = r / M, k = r % M;
  float4 g[VEC], f[VEC];
  load_row<VEC>(dv + (int64_t)r * d, lane, g);
  load_row<VEC>(v + (int64_t)r * d, lane, f);
  const float invn = inv_norm[r];
  if (invn < 1e12f) {
    const float dot = row_dot<VEC>(f, g);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      F4_OP(g[i], invn * (g[i].x - f[i].x * dot), invn * (g[i].y - f[i].y * dot),
            invn * (g[i].z - f[i].z * dot), invn * (g[i].w - f[i].w * dot));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) F4_OP(g[i], g[i].x * invn, g[i].y * invn, g[i].z * invn, g[i].w * invn);
  }
  store_row<VEC>(dh + ((int64_t)b * S + 1 + (int64_t)k * (T + 1)) * d, lane, g);
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradients) and elementwise dropout
// ------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t rows, int n,
                                                     int64_t ld, int rb, int64_t rbs,
                                                     float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 red[8][32];
  const int c4 = blockIdx.x * 32 + threadIdx.x;        // float4 column index
  const int ty = threadIdx.y;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < n) {
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < rows; r += (int64_t)gridDim.y * 8) {
      const float* p = X + (rb > 0 ? (r / rb) * rbs + (r % rb) * ld : r * ld) + 4 * c4;
      if (VEC4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      } else {
        acc.x += p[0];
        if (4 * c4 + 1 < n) acc.y += p[1];
        if (4 * c4 + 2 < n) acc.z += p[2];
        if (4 * c4 + 3 < n) acc.w += p[3];
      }
    }
  }
  red[ty][threadIdx.x] = acc;
  __syncthreads();
  if (ty == 0 && c4 * 4 < n) {
    for (int w = 1; w < 8; ++w) {
      float4 t = red[w][threadIdx.x];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (VEC4) atomic_add4(out + 4 * c4, acc);
    else {
      atomicAdd(out + 4 * c4, acc.x);
      if (4 * c4 + 1 < n) atomicAdd(out + 4 * c4 + 1, acc.y);
      if (4 * c4 + 2 < n) atomicAdd(out + 4 * c4 + 2, acc.z);
      if (4 * c4 + 3 < n) atomicAdd(out + 4 * c4 + 3, acc.w);
    }
  }
}

__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ in,
                                                      float* __restrict__ out, int64_t rows, int n4,
                                                      float p, uint64_t seed, uint32_t site, const uint64_t* __restrict__ ctr) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;        // device-side step counter (CUDA-graph replays)
  const float inv_keep = 1.f / (1.f - p);
  const int64_t total = rows * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n4;
    const int c4 = (int)(i % n4);
    float4 v = reinterpret_cast<const float4*>(in)[i];
    float4 sc = dropout_scale4(seed, site, (uint32_t)r, c4, p, inv_keep);
    F4_OP(v, v.x * sc.x, v.y * sc.y, v.z * sc.z, v.w * sc.w);
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// fp32 -> bf16 round-to-nearest-even, 4 elements per thread (operands of the experimental MMT_PREC_BF16 GEMM)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, int64_t n4) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    uint32_t lo, hi;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v.y), "f"(v.x));     // d = {hi: first src, lo: second src}
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.w), "f"(v.z));
    out[i] = make_uint2(lo, hi);
  }
}


// ------------------------------------------------------------------------------------------
// 16-bit operand path: plain LayerNorm of the GEMM-epilogue output, operand casts, input packing
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) ln16_fwd_kernel(
    const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows,
    float eps, float* __restrict__ y, void* __restrict__ y16, float* __restrict__ mean_o, float* __restrict__ rstd_o,
    int bf16) {
  pdl_trigger();
  pdl_wait();
  constexpr int d = 128 * VEC;
  const int lane = threadIdx.x & 31;
  float4 g[VEC], bt[VEC];
  load_row<VEC>(gamma, lane, g);
  load_row<VEC>(beta, lane, bt);
  for (int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * WARPS) {
    float4 x[VEC];
    load_row<VEC>(z + r * d, lane, x);
    float mean, rstd;
    ln_stats<VEC>(x, d, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      F4_OP(x[i], (x[i].x - mean) * rstd * g[i].x + bt[i].x, (x[i].y - mean) * rstd * g[i].y + bt[i].y,
            (x[i].z - mean) * rstd * g[i].z + bt[i].z, (x[i].w - mean) * rstd * g[i].w + bt[i].w);
    if (y != nullptr) store_row<VEC>(y + r * d, lane, x);
    if (y16 != nullptr) store_row16<VEC>(reinterpret_cast<uint16_t*>(y16) + r * d, lane, x, 1.0f, bf16 != 0);
    if (lane == 0) { mean_o[r] = mean; rstd_o[r] = rstd; }
  }
}

// out[r, c4*4 .. +3] = rn16(mask * in * scale); columns >= cols are written as zero up to out_cols
__global__ void __launch_bounds__(256) cast16_kernel(const float* __restrict__ in, int64_t rows, int cols, int64_t in_ld,
                                                     uint16_t* __restrict__ out, uint16_t* __restrict__ out_lo, int out_cols,
                                                     int64_t out_ld, float scale, float p_drop, uint64_t seed,
                                                     const uint64_t* __restrict__ ctr, uint32_t site, int bf16, int vec) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int n4 = (out_cols + 3) >> 2;
  const int64_t total = rows * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n4;
    const int c = (int)(i % n4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = in + r * in_ld + c;
    if (vec && c + 4 <= cols) v = *reinterpret_cast<const float4*>(p);
    else {
      if (c < cols) v.x = p[0];
      if (c + 1 < cols) v.y = p[1];
      if (c + 2 < cols) v.z = p[2];
      if (c + 3 < cols) v.w = p[3];
    }
    if (p_drop > 0.f) {
      const float4 sc = dropout_scale4_fast(drop_key(seed, site), (uint32_t)r, (uint32_t)(c >> 2),
                                            (uint32_t)(p_drop * 65536.0f), inv_keep);
      F4_OP(v, v.x * sc.x, v.y * sc.y, v.z * sc.z, v.w * sc.w);
    }
    F4_OP(v, v.x * scale, v.y * scale, v.z * scale, v.w * scale);
    uint16_t* q = out + r * out_ld + c;
    const uint2 hi = pack4(v, bf16 != 0);
    if (vec && c + 4 <= out_cols) *reinterpret_cast<uint2*>(q) = hi;
    else {
      const float vv[4] = {v.x, v.y, v.z, v.w};
      for (int t = 0; t < 4; ++t) if (c + t < out_cols) q[t] = pack1(vv[t], bf16 != 0);
    }
    if (out_lo != nullptr) {
      // second term of the two-term split x ~= hi + lo / 2048 (lo carries the next 11 bits of the significand)
      const float4 h = unpack4(hi, bf16 != 0);
      const float4 lo = make_float4((v.x - h.x) * 2048.f, (v.y - h.y) * 2048.f, (v.z - h.z) * 2048.f, (v.w - h.w) * 2048.f);
      uint16_t* ql = out_lo + r * out_ld + c;
      if (vec && c + 4 <= out_cols) *reinterpret_cast<uint2*>(ql) = pack4(lo, bf16 != 0);
      else {
        const float vv[4] = {lo.x, lo.y, lo.z, lo.w};
        for (int t = 0; t < 4; ++t) if (c + t < out_cols) ql[t] = pack1(vv[t], bf16 != 0);
      }
    }
  }
}

// blockIdx.y = expert; rows (b, j) of [B, T+1]: j == 0 -> maxp[b], else feats[b, j-1]
__global__ void __launch_bounds__(256) pack_inputs16_kernel(const mmt_pack_desc pd) {
  pdl_trigger();
  pdl_wait();
  const int k = blockIdx.y;
  const int in = pd.in[k], ld = pd.ld[k], T = pd.T;
  const bool bf16 = pd.dtype == MMT_DT_BF16;
  const float* __restrict__ feats = pd.feats[k];
  const float* __restrict__ maxp = pd.maxp[k];
  uint16_t* __restrict__ out = reinterpret_cast<uint16_t*>(pd.out[k]);
  const int n4 = ld >> 2;                                  // ld % 8 == 0
  const int64_t total = (int64_t)pd.B * (T + 1) * n4;
  const bool vec = (in & 3) == 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n4;
    const int c = (int)(i % n4) * 4;
    const int b = (int)(r / (T + 1)), j = (int)(r % (T + 1));
    const float* p = (j == 0 ? maxp + (int64_t)b * in : feats + ((int64_t)b * T + (j - 1)) * in) + c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && c + 4 <= in) v = *reinterpret_cast<const float4*>(p);
    else {
      if (c < in) v.x = p[0];
      if (c + 1 < in) v.y = p[1];
      if (c + 2 < in) v.z = p[2];
      if (c + 3 < in) v.w = p[3];
    }
    *reinterpret_cast<uint2*>(out + r * ld + c) = pack4(v, bf16);
  }
}

}  // namespace
}  // namespace mmt

using namespace mmt;

extern "C" {

int mmt_embed_ln_fwd(const float* proj, const float* ft, const float* ind, const int32_t* type_idx,
                     const float* pos_emb, const float* type_emb, const float* gamma,
                     const float* beta, int32_t B, int32_t M, int32_t T, int32_t d,
                     int32_t max_pos, float eps, float p_drop, uint64_t seed, uint32_t site,
                     float* h, float* mask, int32_t* pos_ids, int32_t* type_ids, float* inv_norm,
                     float* mean, float* rstd, void* stream) {
  MMT_ARG_CHECK(proj && ft && ind && type_idx && pos_emb && type_emb && gamma && beta && h && mask &&
                pos_ids && type_ids && inv_norm && mean && rstd, MMT_E_ARG, "mmt_embed_ln_fwd: null pointer");
  MMT_ARG_CHECK(B > 0 && M > 0 && T > 0 && max_pos > 0, MMT_E_SHAPE, "mmt_embed_ln_fwd: bad shape B=%d M=%d T=%d", B, M, T);
  CHECK_D(d); CHECK_P(p_drop);
  const int64_t rows = (int64_t)B * (1 + M * (T + 1));
  DISPATCH_VEC(d, (launch_pdl(embed_ln_fwd_kernel<V>, dim3(row_grid(rows)), dim3(WARPS * 32), 0, (cudaStream_t)stream, 
      proj, ft, ind, type_idx, pos_emb, type_emb, gamma, beta, B, M, T, max_pos, eps, p_drop, seed,
      site, g_step_ctr, h, mask, pos_ids, type_ids, inv_norm, mean, rstd, (void*)nullptr, 0)));
  MMT_LAUNCH_CHECK("embed_ln_fwd");
  return 0;
}

int mmt_embed_ln_bwd(const float* dh, const float* proj, const int32_t* pos_ids,
                     const int32_t* type_ids, const float* inv_norm, const float* mean,
                     const float* rstd, const float* pos_emb, const float* type_emb,
                     const float* gamma, int32_t B, int32_t M, int32_t T, int32_t d, float p_drop,
                     uint64_t seed, uint32_t site, float* dproj, float* dpos_emb,
                     float* dtype_emb, float* dgamma, float* dbeta, void* stream) {
  MMT_ARG_CHECK(dh && proj && pos_ids && type_ids && inv_norm && mean && rstd && pos_emb && type_emb &&
                gamma && dproj && dpos_emb && dtype_emb && dgamma && dbeta, MMT_E_ARG, "mmt_embed_ln_bwd: null pointer");
  CHECK_D(d); CHECK_P(p_drop);
  const int64_t rows = (int64_t)B * (1 + M * (T + 1));
  int grid = row_grid(rows);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  DISPATCH_VEC(d, (launch_pdl(embed_ln_bwd_kernel<V>, dim3(grid), dim3(WARPS * 32), 0, (cudaStream_t)stream, 
      dh, proj, pos_ids, type_ids, inv_norm, mean, rstd, pos_emb, type_emb, gamma, B, M, T, p_drop,
      seed, site, g_step_ctr, dproj, dpos_emb, dtype_emb, dgamma, dbeta, (void*)nullptr, 1.0f, 0)));
  MMT_LAUNCH_CHECK("embed_ln_bwd");
  return 0;
}

int mmt_res_ln_fwd(float* t, const float* r, const float* gamma, const float* beta, int64_t rows,
                   int32_t d, float eps, float p_drop, uint64_t seed, uint32_t site, float* y,
                   float* mean, float* rstd, void* stream) {
  MMT_ARG_CHECK(t && r && gamma && beta && y && mean && rstd, MMT_E_ARG, "mmt_res_ln_fwd: null pointer");
  CHECK_D(d); CHECK_P(p_drop);
  if (rows == 0) return 0;
  DISPATCH_VEC(d, (launch_pdl(res_ln_fwd_kernel<V>, dim3(row_grid(rows)), dim3(WARPS * 32), 0, (cudaStream_t)stream, 
      t, r, gamma, beta, rows, eps, p_drop, seed, site, g_step_ctr, y, mean, rstd)));
  MMT_LAUNCH_CHECK("res_ln_fwd");
  return 0;
}

int mmt_res_ln_bwd(const float* dy, const float* dy2, const float* z, const float* mean,
                   const float* rstd, const float* gamma, int64_t rows, int32_t d, float p_drop,
                   uint64_t seed, uint32_t site, float* dz, float* dt, float* dgamma, float* dbeta,
                   float* dbias, void* stream) {
  MMT_ARG_CHECK(dy && z && mean && rstd && gamma && dz && dgamma && dbeta, MMT_E_ARG, "mmt_res_ln_bwd: null pointer");
  MMT_ARG_CHECK(p_drop == 0.f || dt != nullptr, MMT_E_ARG, "mmt_res_ln_bwd: dt required when p_drop > 0");
  CHECK_D(d); CHECK_P(p_drop);
  if (rows == 0) return 0;
  int grid = row_grid(rows);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  DISPATCH_VEC(d, (launch_pdl(res_ln_bwd_kernel<V>, dim3(grid), dim3(WARPS * 32), 0, (cudaStream_t)stream, 
      dy, dy2, z, mean, rstd, gamma, rows, p_drop, seed, site, g_step_ctr, dz, dt, dgamma, dbeta, dbias,
      (void*)nullptr, 1.0f, 0)));
  MMT_LAUNCH_CHECK("res_ln_bwd");
  return 0;
}

int mmt_softmax_mask_fwd(const float* scores, const float* mask, int32_t B, int32_t H, int32_t S,
                         int32_t ld, float scale, float p_drop, uint64_t seed, uint32_t site,
                         float* Psoft, float* Pdrop, void* stream) {
  MMT_ARG_CHECK(scores && mask && Psoft, MMT_E_ARG, "mmt_softmax_mask_fwd: null pointer");
  MMT_ARG_CHECK(p_drop == 0.f || Pdrop != nullptr, MMT_E_ARG, "mmt_softmax_mask_fwd: Pdrop required when p_drop > 0");
  MMT_ARG_CHECK(S > 0 && S <= 128 * SM_MAXC && ld >= S && ld % 4 == 0 && ld <= 128 * SM_MAXC, MMT_E_SHAPE,
                "mmt_softmax_mask_fwd: S=%d ld=%d unsupported (S <= %d, ld %% 4 == 0)", S, ld, 128 * SM_MAXC);
  CHECK_P(p_drop);
  launch_pdl(softmax_fwd_kernel, dim3(row_grid((int64_t)B * H * S)), dim3(WARPS * 32), 0, (cudaStream_t)stream, 
      scores, mask, B, H, S, ld, scale, p_drop, seed, site, g_step_ctr, Psoft, Pdrop);
  MMT_LAUNCH_CHECK("softmax_fwd");
  return 0;
}

int mmt_softmax_mask_bwd(float* dP, const float* Psoft, int32_t B, int32_t H, int32_t S, int32_t ld,
                         float scale, float p_drop, uint64_t seed, uint32_t site, void* stream) {
  MMT_ARG_CHECK(dP && Psoft, MMT_E_ARG, "mmt_softmax_mask_bwd: null pointer");
  MMT_ARG_CHECK(S > 0 && ld >= S && ld % 4 == 0 && ld <= 128 * SM_MAXC, MMT_E_SHAPE, "mmt_softmax_mask_bwd: S=%d ld=%d unsupported", S, ld);
  CHECK_P(p_drop);
  const int64_t rows = (int64_t)B * H * S;
  launch_pdl(softmax_bwd_kernel, dim3(row_grid(rows)), dim3(WARPS * 32), 0, (cudaStream_t)stream, dP, Psoft, rows, S, ld, scale, p_drop, seed, site, g_step_ctr);
  MMT_LAUNCH_CHECK("softmax_bwd");
  return 0;
}

int mmt_readout_norm_fwd(const float* h, int32_t B, int32_t S, int32_t M, int32_t T, int32_t d,
                         float* v, float* inv_norm, void* stream) {
  MMT_ARG_CHECK(h && v && inv_norm, MMT_E_ARG, "mmt_readout_norm_fwd: null pointer");
  MMT_ARG_CHECK(S == 1 + M * (T + 1), MMT_E_SHAPE, "mmt_readout_norm_fwd: S=%d != 1+M*(T+1)", S);
  CHECK_D(d);
  DISPATCH_VEC(d, (launch_pdl(readout_fwd_kernel<V>, dim3((B * M + WARPS - 1) / WARPS), dim3(WARPS * 32), 0, (cudaStream_t)stream, h, B, S, M, T, v, inv_norm)));
  MMT_LAUNCH_CHECK("readout_fwd");
  return 0;
}

int mmt_readout_norm_bwd(const float* dv, const float* v, const float* inv_norm, int32_t B,
                         int32_t S, int32_t M, int32_t T, int32_t d, float* dh, void* stream) {
  MMT_ARG_CHECK(dv && v && inv_norm && dh, MMT_E_ARG, "mmt_readout_norm_bwd: null pointer");
  MMT_ARG_CHECK(S == 1 + M * (T + 1), MMT_E_SHAPE, "mmt_readout_norm_bwd: S=%d != 1+M*(T+1)", S);
  CHECK_D(d);
  cudaError_t e = cudaMemsetAsync(dh, 0, sizeof(float) * (size_t)B * S * d, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_status(e, "readout_bwd memset");
  DISPATCH_VEC(d, (launch_pdl(readout_bwd_kernel<V>, dim3((B * M + WARPS - 1) / WARPS), dim3(WARPS * 32), 0, (cudaStream_t)stream, dv, v, inv_norm, B, S, M, T, dh)));
  MMT_LAUNCH_CHECK("readout_bwd");
  return 0;
}

int mmt_colsum(const float* X, int64_t rows, int32_t n, int64_t ld, int32_t rb, int64_t rbs,
               float* out, int accumulate, void* stream) {
  MMT_ARG_CHECK(X && out, MMT_E_ARG, "mmt_colsum: null pointer");
  MMT_ARG_CHECK(n > 0 && rb >= 0, MMT_E_SHAPE, "mmt_colsum: n=%d rb=%d", n, rb);
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_status(e, "colsum memset");
  }
  if (rows == 0) return 0;
  const bool vec = (n % 4 == 0) && (ld % 4 == 0) && (rbs % 4 == 0) && ((uintptr_t)X % 16 == 0) &&
                   ((uintptr_t)out % 16 == 0);
  int gx = ((n + 3) / 4 + 31) / 32;
  int64_t want = (rows + 63) / 64;
  int64_t cap = 4 * num_sms() / gx + 1;
  int gy = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  if (vec) launch_pdl(colsum_kernel<true>, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, X, rows, n, ld, rb, rbs, out);
  else launch_pdl(colsum_kernel<false>, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, X, rows, n, ld, rb, rbs, out);
  MMT_LAUNCH_CHECK("colsum");
  return 0;
}

int mmt_dropout(const float* in, float* out, int64_t rows, int32_t n, float p, uint64_t seed,
                uint32_t site, void* stream) {
  MMT_ARG_CHECK(in && out, MMT_E_ARG, "mmt_dropout: null pointer");
  MMT_ARG_CHECK(n % 4 == 0, MMT_E_ALIGN, "mmt_dropout: n=%d must be a multiple of 4", n);
  MMT_ARG_CHECK(p > 0.f && p < 1.f, MMT_E_ARG, "mmt_dropout: p=%f out of (0,1)", (double)p);
  const int64_t total = rows * (n / 4);
  if (total == 0) return 0;
  int64_t blocks = (total + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  launch_pdl(dropout_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, in, out, rows, n / 4, p, seed, site, g_step_ctr);
  MMT_LAUNCH_CHECK("dropout");
  return 0;
}

int mmt_cast_bf16(const float* in, void* out_bf16, int64_t n, void* stream) {
  MMT_ARG_CHECK(in && out_bf16, MMT_E_ARG, "mmt_cast_bf16: null pointer");
  MMT_ARG_CHECK(n >= 0 && (n & 3) == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out_bf16 % 8) == 0, MMT_E_ALIGN,
                "mmt_cast_bf16: n must be a multiple of 4 and the buffers aligned");
  if (n == 0) return 0;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  launch_pdl(cast_bf16_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(in),
             reinterpret_cast<uint2*>(out_bf16), n / 4);
  MMT_LAUNCH_CHECK("cast_bf16");
  return 0;
}

int mmt_cast16(const float* in, int64_t rows, int32_t cols, int64_t in_ld, void* out, void* out_lo, int32_t out_cols,
               int64_t out_ld, float scale, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site,
               int32_t dtype, void* stream) {
  MMT_ARG_CHECK(in && out, MMT_E_ARG, "mmt_cast16: null pointer");
  MMT_ARG_CHECK(rows >= 0 && cols >= 0 && out_cols >= cols && in_ld >= cols && out_ld >= out_cols, MMT_E_SHAPE,
                "mmt_cast16: bad shape rows=%lld cols=%d out_cols=%d", (long long)rows, cols, out_cols);
  MMT_ARG_CHECK(dtype == MMT_DT_F16 || dtype == MMT_DT_BF16, MMT_E_ARG, "mmt_cast16: bad dtype %d", dtype);
  CHECK_P(p_drop);
  const int64_t total = rows * ((out_cols + 3) / 4);
  if (total == 0) return 0;
  const int vec = ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 8 == 0) && ((uintptr_t)out_lo % 8 == 0) &&
                  (in_ld % 4 == 0) && (out_ld % 4 == 0);
  int64_t blocks = (total + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  launch_pdl(cast16_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, in, rows, cols, in_ld,
             reinterpret_cast<uint16_t*>(out), reinterpret_cast<uint16_t*>(out_lo), out_cols, out_ld, scale, p_drop, seed,
             seed_ctr, site,
             dtype == MMT_DT_BF16 ? 1 : 0, vec);
  MMT_LAUNCH_CHECK("cast16");
  return 0;
}

int mmt_pack_inputs16(const mmt_pack_desc* d, void* stream) {
  MMT_ARG_CHECK(d != nullptr, MMT_E_ARG, "mmt_pack_inputs16: null descriptor");
  MMT_ARG_CHECK(d->n >= 1 && d->n <= MMT_MAX_EXPERTS && d->B > 0 && d->T > 0, MMT_E_SHAPE,
                "mmt_pack_inputs16: n=%d B=%d T=%d", d->n, d->B, d->T);
  int max_ld = 0;
  for (int k = 0; k < d->n; ++k) {
    MMT_ARG_CHECK(d->feats[k] && d->maxp[k] && d->out[k], MMT_E_ARG, "mmt_pack_inputs16: null pointer (expert %d)", k);
    MMT_ARG_CHECK(d->ld[k] >= d->in[k] && d->ld[k] % 8 == 0 && ((uintptr_t)d->out[k] % 16) == 0, MMT_E_ALIGN,
                  "mmt_pack_inputs16: expert %d pitch %d (in %d) must be a multiple of 8", k, d->ld[k], d->in[k]);
    MMT_ARG_CHECK(((uintptr_t)d->feats[k] % 16) == 0 && ((uintptr_t)d->maxp[k] % 16) == 0, MMT_E_ALIGN,
                  "mmt_pack_inputs16: expert %d inputs must be 16-byte aligned", k);
    if (d->ld[k] > max_ld) max_ld = d->ld[k];
  }
  const int64_t total = (int64_t)d->B * (d->T + 1) * (max_ld / 4);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16 / d->n + 1;
  if (blocks > cap) blocks = cap;
  launch_pdl(pack_inputs16_kernel, dim3((int)blocks, d->n), dim3(256), 0, (cudaStream_t)stream, *d);
  MMT_LAUNCH_CHECK("pack_inputs16");
  return 0;
}

int mmt_embed_ln16_fwd(const float* proj, const float* ft, const float* ind, const int32_t* type_idx,
                       const float* pos_emb, const float* type_emb, const float* gamma, const float* beta,
                       int32_t B, int32_t M, int32_t T, int32_t d, int32_t max_pos, float eps, float p_drop,
                       uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float* h, void* h16, float* mask,
                       int32_t* pos_ids, int32_t* type_ids, float* inv_norm, float* mean, float* rstd,
                       int32_t dtype, void* stream) {
  MMT_ARG_CHECK(proj && ft && ind && type_idx && pos_emb && type_emb && gamma && beta && h && h16 && mask &&
                pos_ids && type_ids && inv_norm && mean && rstd, MMT_E_ARG, "mmt_embed_ln16_fwd: null pointer");
  MMT_ARG_CHECK(B > 0 && M > 0 && T > 0 && max_pos > 0, MMT_E_SHAPE, "mmt_embed_ln16_fwd: bad shape B=%d M=%d T=%d", B, M, T);
  CHECK_D(d); CHECK_P(p_drop);
  const int64_t rows = (int64_t)B * (1 + M * (T + 1));
  DISPATCH_VEC(d, (launch_pdl(embed_ln_fwd_kernel<V>, dim3(row_grid(rows)), dim3(WARPS * 32), 0, (cudaStream_t)stream,
      proj, ft, ind, type_idx, pos_emb, type_emb, gamma, beta, B, M, T, max_pos, eps, p_drop, seed,
      site, seed_ctr, h, mask, pos_ids, type_ids, inv_norm, mean, rstd, h16, dtype == MMT_DT_BF16 ? 1 : 0)));
  MMT_LAUNCH_CHECK("embed_ln16_fwd");
  return 0;
}

int mmt_embed_ln16_bwd(const float* dh, const float* proj, const int32_t* pos_ids, const int32_t* type_ids,
                       const float* inv_norm, const float* mean, const float* rstd, const float* pos_emb,
                       const float* type_emb, const float* gamma, int32_t B, int32_t M, int32_t T, int32_t d,
                       float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float* dproj,
                       void* dproj16, float scale16, float* dpos_emb, float* dtype_emb, float* dgamma,
                       float* dbeta, int32_t dtype, void* stream) {
  MMT_ARG_CHECK(dh && proj && pos_ids && type_ids && inv_norm && mean && rstd && pos_emb && type_emb &&
                gamma && dproj && dproj16 && dpos_emb && dtype_emb && dgamma && dbeta, MMT_E_ARG, "mmt_embed_ln16_bwd: null pointer");
  CHECK_D(d); CHECK_P(p_drop);
  const int64_t rows = (int64_t)B * (1 + M * (T + 1));
  int grid = row_grid(rows);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  DISPATCH_VEC(d, (launch_pdl(embed_ln_bwd_kernel<V>, dim3(grid), dim3(WARPS * 32), 0, (cudaStream_t)stream,
      dh, proj, pos_ids, type_ids, inv_norm, mean, rstd, pos_emb, type_emb, gamma, B, M, T, p_drop,
      seed, site, seed_ctr, dproj, dpos_emb, dtype_emb, dgamma, dbeta, dproj16, scale16, dtype == MMT_DT_BF16 ? 1 : 0)));
  MMT_LAUNCH_CHECK("embed_ln16_bwd");
  return 0;
}

int mmt_ln16_fwd(const float* z, const float* gamma, const float* beta, int64_t rows, int32_t d, float eps,
                 float* y, void* y16, float* mean, float* rstd, int32_t dtype, void* stream) {
  MMT_ARG_CHECK(z && gamma && beta && (y || y16) && mean && rstd, MMT_E_ARG, "mmt_ln16_fwd: null pointer");
  CHECK_D(d);
  if (rows == 0) return 0;
  DISPATCH_VEC(d, (launch_pdl(ln16_fwd_kernel<V>, dim3(row_grid(rows)), dim3(WARPS * 32), 0, (cudaStream_t)stream,
      z, gamma, beta, rows, eps, y, y16, mean, rstd, dtype == MMT_DT_BF16 ? 1 : 0)));
  MMT_LAUNCH_CHECK("ln16_fwd");
  return 0;
}

int mmt_ln16_bwd(const float* dy, const float* dy2, const float* z, const float* mean, const float* rstd,
                 const float* gamma, int64_t rows, int32_t d, float p_drop, uint64_t seed,
                 const uint64_t* seed_ctr, uint32_t site, float* dz, void* dt16, float scale16, float* dgamma,
                 float* dbeta, float* dbias, int32_t dtype, void* stream) {
  MMT_ARG_CHECK(dy && z && mean && rstd && gamma && dz && dt16 && dgamma && dbeta, MMT_E_ARG, "mmt_ln16_bwd: null pointer");
  CHECK_D(d); CHECK_P(p_drop);
  if (rows == 0) return 0;
  int grid = row_grid(rows);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  DISPATCH_VEC(d, (launch_pdl(res_ln_bwd_kernel<V>, dim3(grid), dim3(WARPS * 32), 0, (cudaStream_t)stream,
      dy, dy2, z, mean, rstd, gamma, rows, p_drop, seed, site, seed_ctr, dz, (float*)nullptr, dgamma, dbeta, dbias,
      dt16, scale16, dtype == MMT_DT_BF16 ? 1 : 0)));
  MMT_LAUNCH_CHECK("ln16_bwd");
  return 0;
}

}  // extern "C"
