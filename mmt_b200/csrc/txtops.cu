// Text-encoder specific kernels (BERT-base geometry: d = 768, 12 heads of dh = 64, W = 30 .. 100 tokens):
// word / position / type embedding gather + LayerNorm (transformers BertEmbeddings; reference model/model.py:
// 350-387 feeds token ids, positions 0..W-1 and token type 0), its backward (scatter-add into the embedding
// tables), and self-attention over short sequences.
//
// Attention here is tiny -- R*H*W^2*dh*4 = 0.18 GFLOP forward at R = 64, W = 30, against 326 GFLOP of GEMMs in the
// same encoder -- and its tiles (30 x 30 scores, dh = 64) are far below a tensor-core tile, so it is a shared-memory
// SIMT kernel: one CTA per (caption, head) keeps Q, K, V (and dO) in shared memory as fp32; probabilities are
// recomputed in the backward (nothing of size W x W is saved).  Dropout decisions use the same pair hash as
// attention16.cu.  The encoder's GEMMs, LayerNorms and epilogue fusions are the video encoder's kernels.
#include "cvt16.cuh"
#include "rowvec.cuh"

namespace mmt {
namespace {

constexpr int TDH = 64;                  // head dim
constexpr int TPITCH = TDH + 1;          // fp32 row pitch in shared memory (conflict-free column walks)
constexpr int TMAXW = 128;               // longest sequence (keys per lane: up to 4)
constexpr float LOG2E_T = 1.44269504088896340736f;

__device__ __forceinline__ uint32_t drop_word_t(uint32_t key32, uint32_t prow, uint32_t half_pitch, uint32_t kpair) {
  return hash32((prow * half_pitch + kpair) ^ key32);
}

// ------------------------------------------------------------------------------------------
// embeddings: h[r*W + w] = dropout(LN(word[ids[r, w]] + pos[w] + type[0]))
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) txt_embed_ln_fwd_kernel(
    const int32_t* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
    const float* __restrict__ type0, const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int W,
    int vocab, float eps, float p_drop, uint64_t seed, const uint64_t* __restrict__ ctr, uint32_t site,
    float* __restrict__ h, void* __restrict__ h16, float* __restrict__ mean_o, float* __restrict__ rstd_o, int bf16) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;
  constexpr int d = 128 * VEC;
  const int lane = threadIdx.x & 31;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t key32 = drop_key(seed, site), thr16 = (uint32_t)(p_drop * 65536.0f);
  float4 g[VEC], bt[VEC], tt[VEC];
  load_row<VEC>(gamma, lane, g);
  load_row<VEC>(beta, lane, bt);
  load_row<VEC>(type0, lane, tt);
  for (int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * WARPS) {
    int id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int w = (int)(r % W);
    float4 e[VEC], pe[VEC];
    load_row<VEC>(word + (int64_t)id * d, lane, e);
    load_row<VEC>(pos + (int64_t)w * d, lane, pe);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      F4_OP(e[i], (e[i].x + tt[i].x) + pe[i].x, (e[i].y + tt[i].y) + pe[i].y, (e[i].z + tt[i].z) + pe[i].z,
            (e[i].w + tt[i].w) + pe[i].w);
    float mean, rstd;
    ln_stats<VEC>(e, d, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      F4_OP(e[i], (e[i].x - mean) * rstd * g[i].x + bt[i].x, (e[i].y - mean) * rstd * g[i].y + bt[i].y,
            (e[i].z - mean) * rstd * g[i].z + bt[i].z, (e[i].w - mean) * rstd * g[i].w + bt[i].w);
      if (p_drop > 0.f) {
        const float4 sc = dropout_scale4_fast(key32, (uint32_t)r, lane + 32 * i, thr16, inv_keep);
        F4_OP(e[i], e[i].x * sc.x, e[i].y * sc.y, e[i].z * sc.z, e[i].w * sc.w);
      }
    }
    store_row<VEC>(h + r * d, lane, e);
    store_row16<VEC>(reinterpret_cast<uint16_t*>(h16) + r * d, lane, e, 1.0f, bf16 != 0);
    if (lane == 0) { mean_o[r] = mean; rstd_o[r] = rstd; }
  }
}

template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) txt_embed_ln_bwd_kernel(
    const float* __restrict__ dh, const int32_t* __restrict__ ids, const float* __restrict__ word,
    const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ mean_i,
    const float* __restrict__ rstd_i, const float* __restrict__ gamma, int64_t rows, int W, int vocab, float p_drop,
    uint64_t seed, const uint64_t* __restrict__ ctr, uint32_t site, float* __restrict__ dword, float* __restrict__ dpos,
    float* __restrict__ dtype0, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_trigger();
  pdl_wait();
  if (ctr != nullptr) seed += *ctr;
  constexpr int d = 128 * VEC;
  __shared__ float4 red[WARPS * VEC * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint32_t key32 = drop_key(seed, site), thr16 = (uint32_t)(p_drop * 65536.0f);
  float4 g[VEC], tt[VEC], ag[VEC], ab[VEC], at[VEC];
  load_row<VEC>(gamma, lane, g);
  load_row<VEC>(type0, lane, tt);
#pragma unroll
  for (int i = 0; i < VEC; ++i) ag[i] = ab[i] = at[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < rows; r += (int64_t)gridDim.x * WARPS) {
    int id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int w = (int)(r % W);
    const float mean = mean_i[r], rstd = rstd_i[r];
    float4 gy[VEC], xh[VEC], pe[VEC];
    load_row<VEC>(dh + r * d, lane, gy);
    if (p_drop > 0.f) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float4 sc = dropout_scale4_fast(key32, (uint32_t)r, lane + 32 * i, thr16, inv_keep);
        F4_OP(gy[i], gy[i].x * sc.x, gy[i].y * sc.y, gy[i].z * sc.z, gy[i].w * sc.w);
      }
    }
    load_row<VEC>(word + (int64_t)id * d, lane, xh);
    load_row<VEC>(pos + (int64_t)w * d, lane, pe);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      F4_OP(xh[i], (((xh[i].x + tt[i].x) + pe[i].x) - mean) * rstd, (((xh[i].y + tt[i].y) + pe[i].y) - mean) * rstd,
            (((xh[i].z + tt[i].z) + pe[i].z) - mean) * rstd, (((xh[i].w + tt[i].w) + pe[i].w) - mean) * rstd);
      ag[i].x += gy[i].x * xh[i].x; ag[i].y += gy[i].y * xh[i].y; ag[i].z += gy[i].z * xh[i].z; ag[i].w += gy[i].w * xh[i].w;
      ab[i].x += gy[i].x; ab[i].y += gy[i].y; ab[i].z += gy[i].z; ab[i].w += gy[i].w;
      F4_OP(gy[i], gy[i].x * g[i].x, gy[i].y * g[i].y, gy[i].z * g[i].z, gy[i].w * g[i].w);
    }
    const float m1 = row_sum<VEC>(gy) / d;
    const float m2 = row_dot<VEC>(gy, xh) / d;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 de;
      F4_OP(de, rstd * (gy[i].x - m1 - xh[i].x * m2), rstd * (gy[i].y - m1 - xh[i].y * m2),
            rstd * (gy[i].z - m1 - xh[i].z * m2), rstd * (gy[i].w - m1 - xh[i].w * m2));
      if (dword != nullptr) atomic_add4(dword + (int64_t)id * d + 4 * (lane + 32 * i), de);
      if (dpos != nullptr) atomic_add4(dpos + (int64_t)w * d + 4 * (lane + 32 * i), de);
      at[i].x += de.x; at[i].y += de.y; at[i].z += de.z; at[i].w += de.w;
    }
  }
  flush_cols<VEC>(ag, dgamma, lane, warp, red);
  flush_cols<VEC>(ab, dbeta, lane, warp, red);
  flush_cols<VEC>(at, dtype0, lane, warp, red);
}

// ------------------------------------------------------------------------------------------
// self-attention over short sequences, one CTA per (caption r, head h)
// ------------------------------------------------------------------------------------------
struct TAttArgs {
  const uint16_t* qkv16;  // [R*W, 3*H*64]
  const float* mask;      // [R, W] 1 = attend
  int R, H, W;
  float scale, p_drop, inv_keep;
  uint64_t seed;
  const uint64_t* ctr;
  uint32_t site;
  int bf16;
};

// loads rows [W][64] of column block `blk` (0 = Q, 1 = K, 2 = V) of head h into fp32 shared memory
__device__ __forceinline__ void load_head(float* dst, const uint16_t* src, int64_t row0, int W, int ld, int col0, bool bf16) {
  for (int i = threadIdx.x; i < W * (TDH / 8); i += blockDim.x) {
    const int j = i / (TDH / 8), c8 = (i % (TDH / 8)) * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(src + (row0 + j) * ld + col0 + c8);
    const float2 a = unpack2(u.x, bf16), b = unpack2(u.y, bf16), c = unpack2(u.z, bf16), d = unpack2(u.w, bf16);
    float* p = dst + j * TPITCH + c8;
    p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y; p[4] = c.x; p[5] = c.y; p[6] = d.x; p[7] = d.y;
  }
}

__global__ void __launch_bounds__(128) txt_attention_fwd_kernel(const TAttArgs a, uint16_t* __restrict__ ctx16) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int W = a.W, H = a.H, h = blockIdx.x, r = blockIdx.y;
  const int d_model = H * TDH, ld = 3 * d_model;
  float* sK = sm;                        // [W][65]
  float* sV = sK + W * TPITCH;
  float* sQ = sV + W * TPITCH;
  float* sMask = sQ + W * TPITCH;        // [W] additive mask (log2 domain)
  const bool bf16 = a.bf16 != 0;
  const int64_t row0 = (int64_t)r * W;
  load_head(sQ, a.qkv16, row0, W, ld, h * TDH, bf16);
  load_head(sK, a.qkv16, row0, W, ld, d_model + h * TDH, bf16);
  load_head(sV, a.qkv16, row0, W, ld, 2 * d_model + h * TDH, bf16);
  for (int j = threadIdx.x; j < W; j += blockDim.x) sMask[j] = (1.0f - a.mask[row0 + j]) * (-10000.0f * LOG2E_T);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t seed = a.seed + (a.ctr ? *a.ctr : 0);
  const uint32_t key32 = drop_key(seed, a.site), thr = (uint32_t)(a.p_drop * 65536.0f);
  const uint32_t half_pitch = (uint32_t)((W + 1) >> 1);
  const float sl2 = a.scale * LOG2E_T;
  for (int i = warp; i < W; i += 4) {
    const float* q = sQ + i * TPITCH;
    float s[4], mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      s[t] = -INFINITY;
      if (j < W) {
        const float* k = sK + j * TPITCH;
        float acc = 0.f;
#pragma unroll 16
        for (int c = 0; c < TDH; ++c) acc = fmaf(q[c], k[c], acc);
        s[t] = fmaf(acc, sl2, sMask[j]);
      }
      mx = fmaxf(mx, s[t]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) { s[t] = fast_ex2(s[t] - mx); sum += s[t]; }     // 2^-inf = 0 past W
    sum = warp_sum(sum);
    const float inv = a.inv_keep / sum;
    if (a.p_drop > 0.f) {
      const uint32_t prow = (uint32_t)(((int64_t)r * H + h) * W + i);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = lane + 32 * t;
        const uint32_t w = drop_word_t(key32, prow, half_pitch, (uint32_t)j >> 1);
        if (((w >> ((j & 1) * 16)) & 0xffffu) < thr) s[t] = 0.f;
      }
    }
    float o0 = 0.f, o1 = 0.f;                                // channels 2*lane, 2*lane + 1
    for (int j = 0; j < W; ++j) {
      const float p = __shfl_sync(0xffffffffu, s[j >> 5], j & 31);
      o0 = fmaf(p, sV[j * TPITCH + 2 * lane], o0);
      o1 = fmaf(p, sV[j * TPITCH + 2 * lane + 1], o1);
    }
    *reinterpret_cast<uint32_t*>(ctx16 + (row0 + i) * d_model + h * TDH + 2 * lane) = pack2(o0 * inv, o1 * inv, bf16);
  }
}

// backward: recomputes the probabilities; dO carries scale16, so do dQ / dK / dV
__global__ void __launch_bounds__(128) txt_attention_bwd_kernel(const TAttArgs a, const uint16_t* __restrict__ dctx16,
                                                                uint16_t* __restrict__ dqkv16) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int W = a.W, H = a.H, h = blockIdx.x, r = blockIdx.y;
  const int d_model = H * TDH, ld = 3 * d_model, WP = W + 1;
  float* sK = sm;
  float* sV = sK + W * TPITCH;
  float* sQ = sV + W * TPITCH;
  float* sdO = sQ + W * TPITCH;
  float* sP = sdO + W * TPITCH;          // [W][W+1]  Pd  (dropped, scaled probabilities)
  float* sdS = sP + W * WP;              // [W][W+1]  dS
  float* sMask = sdS + W * WP;
  const bool bf16 = a.bf16 != 0;
  const int64_t row0 = (int64_t)r * W;
  load_head(sQ, a.qkv16, row0, W, ld, h * TDH, bf16);
  load_head(sK, a.qkv16, row0, W, ld, d_model + h * TDH, bf16);
  load_head(sV, a.qkv16, row0, W, ld, 2 * d_model + h * TDH, bf16);
  load_head(sdO, dctx16, row0, W, d_model, h * TDH, bf16);
  for (int j = threadIdx.x; j < W; j += blockDim.x) sMask[j] = (1.0f - a.mask[row0 + j]) * (-10000.0f * LOG2E_T);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t seed = a.seed + (a.ctr ? *a.ctr : 0);
  const uint32_t key32 = drop_key(seed, a.site), thr = (uint32_t)(a.p_drop * 65536.0f);
  const uint32_t half_pitch = (uint32_t)((W + 1) >> 1);
  const float sl2 = a.scale * LOG2E_T;
  for (int i = warp; i < W; i += 4) {
    const float* q = sQ + i * TPITCH;
    const float* dO = sdO + i * TPITCH;
    float s[4], dp[4], mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      s[t] = -INFINITY; dp[t] = 0.f;
      if (j < W) {
        const float* k = sK + j * TPITCH;
        const float* v = sV + j * TPITCH;
        float acc = 0.f, acc2 = 0.f;
#pragma unroll 16
        for (int c = 0; c < TDH; ++c) { acc = fmaf(q[c], k[c], acc); acc2 = fmaf(dO[c], v[c], acc2); }
        s[t] = fmaf(acc, sl2, sMask[j]);
        dp[t] = acc2;
      }
      mx = fmaxf(mx, s[t]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) { s[t] = fast_ex2(s[t] - mx); sum += s[t]; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    const uint32_t prow = (uint32_t)(((int64_t)r * H + h) * W + i);
    float delta = 0.f, keep[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      keep[t] = a.inv_keep;
      if (a.p_drop > 0.f) {
        const uint32_t w = drop_word_t(key32, prow, half_pitch, (uint32_t)j >> 1);
        if (((w >> ((j & 1) * 16)) & 0xffffu) < thr) keep[t] = 0.f;
      }
      s[t] *= inv;                                           // P
      delta = fmaf(s[t] * keep[t], dp[t], delta);
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      if (j < W) {
        sP[i * WP + j] = s[t] * keep[t];
        sdS[i * WP + j] = s[t] * fmaf(dp[t], keep[t], -delta) * a.scale;
      }
    }
  }
  __syncthreads();
  // dQ_i = sum_j dS_ij K_j ; dK_j = sum_i dS_ij Q_i ; dV_j = sum_i Pd_ij dO_i   (lane = channel pair)
  for (int i = warp; i < W; i += 4) {
    float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int j = 0; j < W; ++j) {
      const float ds_ij = sdS[i * WP + j], ds_ji = sdS[j * WP + i], p_ji = sP[j * WP + i];
      q0 = fmaf(ds_ij, sK[j * TPITCH + 2 * lane], q0);   q1 = fmaf(ds_ij, sK[j * TPITCH + 2 * lane + 1], q1);
      k0 = fmaf(ds_ji, sQ[j * TPITCH + 2 * lane], k0);   k1 = fmaf(ds_ji, sQ[j * TPITCH + 2 * lane + 1], k1);
      v0 = fmaf(p_ji, sdO[j * TPITCH + 2 * lane], v0);   v1 = fmaf(p_ji, sdO[j * TPITCH + 2 * lane + 1], v1);
    }
    uint16_t* o = dqkv16 + (row0 + i) * ld + h * TDH + 2 * lane;
    *reinterpret_cast<uint32_t*>(o) = pack2(q0, q1, bf16);
    *reinterpret_cast<uint32_t*>(o + d_model) = pack2(k0, k1, bf16);
    *reinterpret_cast<uint32_t*>(o + 2 * d_model) = pack2(v0, v1, bf16);
  }
}

// out[n] += scale * sum_r X16[r*ld + n]   (bias gradients from a 16-bit gradient tensor)
__global__ void __launch_bounds__(256) colsum16_kernel(const uint16_t* __restrict__ X, int64_t rows, int n, int64_t ld,
                                                       float scale, float* __restrict__ out, int bf16) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 red[8][32];
  const int c4 = blockIdx.x * 32 + threadIdx.x;            // group of 4 columns
  const int ty = threadIdx.y;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < n) {
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < rows; r += (int64_t)gridDim.y * 8) {
      const float4 v = unpack4(*reinterpret_cast<const uint2*>(X + r * ld + 4 * c4), bf16 != 0);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[ty][threadIdx.x] = acc;
  __syncthreads();
  if (ty == 0 && c4 * 4 < n) {
    for (int w = 1; w < 8; ++w) {
      const float4 t = red[w][threadIdx.x];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    atomic_add4(out + 4 * c4, make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale));
  }
}

}  // namespace
}  // namespace mmt

using namespace mmt;

extern "C" {

int mmt_txt_embed_ln_fwd(const int32_t* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                         const float* beta, int64_t rows, int32_t W, int32_t vocab, int32_t d, float eps, float p_drop,
                         uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float* h, void* h16, float* mean,
                         float* rstd, int32_t dtype, void* stream) {
  MMT_ARG_CHECK(ids && word && pos && type0 && gamma && beta && h && h16 && mean && rstd, MMT_E_ARG, "mmt_txt_embed_ln_fwd: null pointer");
  MMT_ARG_CHECK(rows >= 0 && W > 0 && vocab > 0 && rows % W == 0, MMT_E_SHAPE, "mmt_txt_embed_ln_fwd: rows=%lld W=%d", (long long)rows, W);
  CHECK_D(d); CHECK_P(p_drop);
  if (rows == 0) return 0;
  DISPATCH_VEC(d, (launch_pdl(txt_embed_ln_fwd_kernel<V>, dim3(row_grid(rows)), dim3(WARPS * 32), 0, (cudaStream_t)stream, ids,
                              word, pos, type0, gamma, beta, rows, W, vocab, eps, p_drop, seed, seed_ctr, site, h, h16, mean,
                              rstd, dtype == MMT_DT_BF16 ? 1 : 0)));
  MMT_LAUNCH_CHECK("txt_embed_ln_fwd");
  return 0;
}

int mmt_txt_embed_ln_bwd(const float* dh, const int32_t* ids, const float* word, const float* pos, const float* type0,
                         const float* mean, const float* rstd, const float* gamma, int64_t rows, int32_t W, int32_t vocab,
                         int32_t d, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float* dword,
                         float* dpos, float* dtype0, float* dgamma, float* dbeta, void* stream) {
  MMT_ARG_CHECK(dh && ids && word && pos && type0 && mean && rstd && gamma && dgamma && dbeta, MMT_E_ARG, "mmt_txt_embed_ln_bwd: null pointer");
  CHECK_D(d); CHECK_P(p_drop);
  if (rows == 0) return 0;
  int grid = row_grid(rows);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  DISPATCH_VEC(d, (launch_pdl(txt_embed_ln_bwd_kernel<V>, dim3(grid), dim3(WARPS * 32), 0, (cudaStream_t)stream, dh, ids, word,
                              pos, type0, mean, rstd, gamma, rows, W, vocab, p_drop, seed, seed_ctr, site, dword, dpos,
                              dtype0, dgamma, dbeta)));
  MMT_LAUNCH_CHECK("txt_embed_ln_bwd");
  return 0;
}

static int txt_att_args(TAttArgs* a, const void* qkv16, const float* mask, int R, int H, int W, int dh, float scale,
                        float p_drop, uint64_t seed, const uint64_t* ctr, uint32_t site, int dtype, const char* who) {
  MMT_ARG_CHECK(qkv16 && mask, MMT_E_ARG, "%s: null pointer", who);
  MMT_ARG_CHECK(dh == TDH, MMT_E_SHAPE, "%s: head dim %d unsupported (only %d)", who, dh, TDH);
  MMT_ARG_CHECK(R > 0 && H > 0 && W > 0 && W <= TMAXW && R <= 65535, MMT_E_SHAPE, "%s: bad shape R=%d H=%d W=%d (W <= %d)", who, R, H, W, TMAXW);
  MMT_ARG_CHECK(p_drop >= 0.f && p_drop < 1.f, MMT_E_ARG, "%s: p_drop=%f", who, (double)p_drop);
  a->qkv16 = reinterpret_cast<const uint16_t*>(qkv16); a->mask = mask;
  a->R = R; a->H = H; a->W = W;
  a->scale = scale; a->p_drop = p_drop; a->inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  a->seed = seed; a->ctr = ctr; a->site = site; a->bf16 = dtype == MMT_DT_BF16 ? 1 : 0;
  return 0;
}

int mmt_txt_attention_fwd(const void* qkv16, const float* mask, int32_t R, int32_t H, int32_t W, int32_t dh, float scale,
                          float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site, void* ctx16, int32_t dtype,
                          void* stream) {
  TAttArgs a;
  int rc = txt_att_args(&a, qkv16, mask, R, H, W, dh, scale, p_drop, seed, seed_ctr, site, dtype, "mmt_txt_attention_fwd");
  if (rc) return rc;
  MMT_ARG_CHECK(ctx16 != nullptr, MMT_E_ARG, "mmt_txt_attention_fwd: null output");
  const size_t smem = sizeof(float) * (3 * (size_t)W * TPITCH + W);
  rc = ensure_dynamic_smem((const void*)txt_attention_fwd_kernel, 110 * 1024, "txt_attention_fwd smem attribute");
  if (rc) return rc;
  launch_pdl(txt_attention_fwd_kernel, dim3(H, R), dim3(128), smem, (cudaStream_t)stream, a, reinterpret_cast<uint16_t*>(ctx16));
  MMT_LAUNCH_CHECK("txt_attention_fwd");
  return 0;
}

int mmt_txt_attention_bwd(const void* qkv16, const void* dctx16, const float* mask, int32_t R, int32_t H, int32_t W,
                          int32_t dh, float scale, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site,
                          void* dqkv16, int32_t dtype, void* stream) {
  TAttArgs a;
  int rc = txt_att_args(&a, qkv16, mask, R, H, W, dh, scale, p_drop, seed, seed_ctr, site, dtype, "mmt_txt_attention_bwd");
  if (rc) return rc;
  MMT_ARG_CHECK(dctx16 && dqkv16, MMT_E_ARG, "mmt_txt_attention_bwd: null pointer");
  const size_t smem = sizeof(float) * (4 * (size_t)W * TPITCH + 2 * (size_t)W * (W + 1) + W);
  rc = ensure_dynamic_smem((const void*)txt_attention_bwd_kernel, 220 * 1024, "txt_attention_bwd smem attribute");
  if (rc) return rc;
  launch_pdl(txt_attention_bwd_kernel, dim3(H, R), dim3(128), smem, (cudaStream_t)stream, a,
             reinterpret_cast<const uint16_t*>(dctx16), reinterpret_cast<uint16_t*>(dqkv16));
  MMT_LAUNCH_CHECK("txt_attention_bwd");
  return 0;
}

int mmt_colsum16(const void* X16, int64_t rows, int32_t n, int64_t ld, float scale, float* out, int32_t dtype, void* stream) {
  MMT_ARG_CHECK(X16 && out, MMT_E_ARG, "mmt_colsum16: null pointer");
  MMT_ARG_CHECK(n > 0 && n % 4 == 0 && ld % 4 == 0 && ((uintptr_t)X16 % 8) == 0 && ((uintptr_t)out % 16) == 0, MMT_E_ALIGN,
                "mmt_colsum16: n=%d ld=%lld must be multiples of 4, aligned buffers", n, (long long)ld);
  if (rows == 0) return 0;
  const int gx = (n / 4 + 31) / 32;
  int64_t want = (rows + 63) / 64;
  const int64_t cap = 4 * num_sms() / gx + 1;
  const int gy = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  launch_pdl(colsum16_kernel, dim3(gx, gy), dim3(32, 8), 0, (cudaStream_t)stream, reinterpret_cast<const uint16_t*>(X16), rows, n,
             ld, scale, out, dtype == MMT_DT_BF16 ? 1 : 0);
  MMT_LAUNCH_CHECK("colsum16");
  return 0;
}

}  // extern "C"
