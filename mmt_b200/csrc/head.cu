// Retrieval head: text GatedEmbeddingUnit tail (BatchNorm + context gating + L2 norm), text
// mixture weights, weighted similarity combine, bi-directional max-margin loss, fused Adam.
// All HBM/latency-bound: coalesced float4 traffic, warp-shuffle reductions.
#include "cvt16.cuh"
#include "rowvec.cuh"

namespace mmt {
namespace {

// ------------------------------------------------------------------------------------------
// BatchNorm statistics over the R rows of G[:, c] for every column c of [R, C]  (C = M*d).
// blockDim = (32 columns, 8 row phases); two-pass mean / biased variance.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ G, int R, int C,
                                                       int training, float momentum, float eps,
                                                       float* __restrict__ run_mean,
                                                       float* __restrict__ run_var,
                                                       float* __restrict__ mean_o,
                                                       float* __restrict__ rstd_o) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  if (!training) {
    if (ty == 0 && c < C) {
      mean_o[c] = run_mean[c];
      rstd_o[c] = 1.0f / sqrtf(run_var[c] + eps);
    }
    return;
  }
  float s = 0.f;
  if (c < C)
    for (int r = ty; r < R; r += 8) s += G[(int64_t)r * C + c];
  red[ty][threadIdx.x] = s;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < 8; ++w) mean += red[w][threadIdx.x];
  mean /= R;
  __syncthreads();
  s = 0.f;
  if (c < C)
    for (int r = ty; r < R; r += 8) {
      float dlt = G[(int64_t)r * C + c] - mean;
      s += dlt * dlt;
    }
  red[ty][threadIdx.x] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float var = 0.f;
    for (int w = 0; w < 8; ++w) var += red[w][threadIdx.x];
    var /= R;                                            // biased variance normalises (BatchNorm1d)
    mean_o[c] = mean;
    rstd_o[c] = 1.0f / sqrtf(var + eps);
    // running statistics: momentum update with the UNBIASED variance (torch.nn.BatchNorm1d)
    const float unb = R > 1 ? var * R / (R - 1) : var;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * unb;
  }
}

// y = x * sigmoid(BN(g)) ; e = normalize(normalize(y))   one warp per (row, expert)
template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) geu_gate_fwd_kernel(
    const float* __restrict__ X, const float* __restrict__ G, const float* __restrict__ bn_w,
    const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
    const float* __restrict__ bn_rstd, int R, int M, float* __restrict__ E, float* __restrict__ Y,
    float* __restrict__ inv_n1, float* __restrict__ inv_n2) {
  pdl_trigger();
  pdl_wait();
  constexpr int d = 128 * VEC;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= (int64_t)R * M) return;
  const int m = (int)(row % M);
  const float* xp = X + row * d;                          // [R, M*d] == [R*M, d]
  float4 x[VEC], g[VEC], w[VEC], b[VEC], mu[VEC], rs[VEC];
  load_row<VEC>(xp, lane, x);
  load_row<VEC>(G + row * d, lane, g);
  load_row<VEC>(bn_w + (int64_t)m * d, lane, w);
  load_row<VEC>(bn_b + (int64_t)m * d, lane, b);
  load_row<VEC>(bn_mean + (int64_t)m * d, lane, mu);
  load_row<VEC>(bn_rstd + (int64_t)m * d, lane, rs);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float* px = reinterpret_cast<float*>(&x[i]);
    const float* pg = reinterpret_cast<const float*>(&g[i]);
    const float* pw = reinterpret_cast<const float*>(&w[i]);
    const float* pb = reinterpret_cast<const float*>(&b[i]);
    const float* pm = reinterpret_cast<const float*>(&mu[i]);
    const float* pr = reinterpret_cast<const float*>(&rs[i]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float gh = (pg[q] - pm[q]) * pr[q] * pw[q] + pb[q];
      px[q] = px[q] * (1.0f / (1.0f + expf(-gh)));        // F.glu(cat(x, x1)) = x * sigmoid(x1)
    }
  }
  store_row<VEC>(Y + row * d, lane, x);
  const float n1 = 1.0f / fmaxf(sqrtf(row_dot<VEC>(x, x)), 1e-12f);   // model.py:700-701
#pragma unroll
  for (int i = 0; i < VEC; ++i) F4_OP(x[i], x[i].x * n1, x[i].y * n1, x[i].z * n1, x[i].w * n1);
  const float n2 = 1.0f / fmaxf(sqrtf(row_dot<VEC>(x, x)), 1e-12f);   // model.py:624-625
#pragma unroll
  for (int i = 0; i < VEC; ++i) F4_OP(x[i], x[i].x * n2, x[i].y * n2, x[i].z * n2, x[i].w * n2);
  store_row<VEC>(E + row * d, lane, x);
  if (lane == 0) { inv_n1[row] = n1; inv_n2[row] = n2; }
}

// Backward stage 1 (row-wise): dE -> dY (two normalize backwards) -> dX_direct, dGhat.
// dGhat [R, M*d] is written into dG; the BatchNorm backward (column-wise) follows in stage 2.
template <int VEC>
__global__ void __launch_bounds__(WARPS * 32) geu_gate_bwd_rows_kernel(
    const float* __restrict__ dE, const float* __restrict__ X, const float* __restrict__ G,
    const float* __restrict__ Y, const float* __restrict__ E, const float* __restrict__ bn_w,
    const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
    const float* __restrict__ bn_rstd, const float* __restrict__ inv_n1,
    const float* __restrict__ inv_n2, int R, int M, float* __restrict__ dX,
    float* __restrict__ dGhat) {
  pdl_trigger();
  pdl_wait();
  constexpr int d = 128 * VEC;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= (int64_t)R * M) return;
  const int m = (int)(row % M);
  const float n1 = inv_n1[row], n2 = inv_n2[row];
  float4 g[VEC], e[VEC], y[VEC];
  load_row<VEC>(dE + row * d, lane, g);
  load_row<VEC>(E + row * d, lane, e);
  load_row<VEC>(Y + row * d, lane, y);
  // second normalize (input z = y*n1, output e): dz = n2 (g - e <e,g>)
  if (n2 < 1e12f) {
    const float dot = row_dot<VEC>(e, g);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      F4_OP(g[i], n2 * (g[i].x - e[i].x * dot), n2 * (g[i].y - e[i].y * dot),
            n2 * (g[i].z - e[i].z * dot), n2 * (g[i].w - e[i].w * dot));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) F4_OP(g[i], g[i].x * n2, g[i].y * n2, g[i].z * n2, g[i].w * n2);
  }
  // first normalize (input y, output z = y*n1): dy = n1 (dz - z <z,dz>)
  if (n1 < 1e12f) {
    float4 z[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) F4_OP(z[i], y[i].x * n1, y[i].y * n1, y[i].z * n1, y[i].w * n1);
    const float dot = row_dot<VEC>(z, g);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      F4_OP(g[i], n1 * (g[i].x - z[i].x * dot), n1 * (g[i].y - z[i].y * dot),
            n1 * (g[i].z - z[i].z * dot), n1 * (g[i].w - z[i].w * dot));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) F4_OP(g[i], g[i].x * n1, g[i].y * n1, g[i].z * n1, g[i].w * n1);
  }
  float4 x[VEC], gg[VEC], w[VEC], b[VEC], mu[VEC], rs[VEC];
  load_row<VEC>(X + row * d, lane, x);
  load_row<VEC>(G + row * d, lane, gg);
  load_row<VEC>(bn_w + (int64_t)m * d, lane, w);
  load_row<VEC>(bn_b + (int64_t)m * d, lane, b);
  load_row<VEC>(bn_mean + (int64_t)m * d, lane, mu);
  load_row<VEC>(bn_rstd + (int64_t)m * d, lane, rs);
  float4 dx[VEC], dgh[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float* pdy = reinterpret_cast<const float*>(&g[i]);
    const float* px = reinterpret_cast<const float*>(&x[i]);
    const float* pg = reinterpret_cast<const float*>(&gg[i]);
    const float* pw = reinterpret_cast<const float*>(&w[i]);
    const float* pb = reinterpret_cast<const float*>(&b[i]);
    const float* pm = reinterpret_cast<const float*>(&mu[i]);
    const float* pr = reinterpret_cast<const float*>(&rs[i]);
    float* pdx = reinterpret_cast<float*>(&dx[i]);
    float* pdg = reinterpret_cast<float*>(&dgh[i]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float gh = (pg[q] - pm[q]) * pr[q] * pw[q] + pb[q];
      const float sg = 1.0f / (1.0f + expf(-gh));
      pdx[q] = pdy[q] * sg;
      pdg[q] = pdy[q] * px[q] * sg * (1.0f - sg);         // gradient w.r.t. BN output
    }
  }
  store_row<VEC>(dX + row * d, lane, dx);
  store_row<VEC>(dGhat + row * d, lane, dgh);
}

// Backward stage 2 (column-wise BatchNorm backward), in place on dG (holds dGhat on entry).
__global__ void __launch_bounds__(256) bn_bwd_kernel(float* __restrict__ dG,
                                                     const float* __restrict__ G,
                                                     const float* __restrict__ bn_w,
                                                     const float* __restrict__ bn_mean,
                                                     const float* __restrict__ bn_rstd, int R, int C,
                                                     int training, float* __restrict__ dbn_w,
                                                     float* __restrict__ dbn_b) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red1[8][33], red2[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  const float mu = c < C ? bn_mean[c] : 0.f, rs = c < C ? bn_rstd[c] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  if (c < C)
    for (int r = ty; r < R; r += 8) {
      const float dg = dG[(int64_t)r * C + c];
      const float xh = (G[(int64_t)r * C + c] - mu) * rs;
      s1 += dg;
      s2 += dg * xh;
    }
  red1[ty][threadIdx.x] = s1;
  red2[ty][threadIdx.x] = s2;
  __syncthreads();
  s1 = 0.f; s2 = 0.f;
  for (int w = 0; w < 8; ++w) { s1 += red1[w][threadIdx.x]; s2 += red2[w][threadIdx.x]; }
  if (c >= C) return;
  if (ty == 0) { dbn_b[c] += s1; dbn_w[c] += s2; }
  const float w = bn_w[c];
  for (int r = ty; r < R; r += 8) {
    const float dg = dG[(int64_t)r * C + c];
    float o;
    if (training) {
      const float xh = (G[(int64_t)r * C + c] - mu) * rs;
      o = w * rs * (dg - s1 / R - xh * s2 / R);
    } else {
      o = w * rs * dg;                                    // eval: statistics are constants
    }
    dG[(int64_t)r * C + c] = o;
  }
}

// ------------------------------------------------------------------------------------------
// text mixture weights: w = L1norm(softmax(logits))   (model/model.py:280-281, 618)
// ------------------------------------------------------------------------------------------
__global__ void moe_softmax_fwd_kernel(const float* __restrict__ logits, int R, int M, int ld,
                                       float* __restrict__ w) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.y + threadIdx.y;
  if (r >= R) return;
  const int lane = threadIdx.x;
  const float x = lane < M ? logits[(int64_t)r * ld + lane] : -INFINITY;
  const float mx = warp_max(x);
  const float e = lane < M ? expf(x - mx) : 0.f;
  const float p = e / warp_sum(e);
  const float l1 = fmaxf(warp_sum(fabsf(p)), 1e-12f);     // F.normalize(p=1, eps=1e-12)
  if (lane < M) w[(int64_t)r * M + lane] = p / l1;
}

__global__ void moe_softmax_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ w,
                                       int R, int M, int ld, float* __restrict__ dlogits) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.y + threadIdx.y;
  if (r >= R) return;
  const int lane = threadIdx.x;
  // w = p / sum(p) with sum(p) == 1 up to rounding, so w == p; the L1 normalisation's Jacobian
  // (I - w 1^T) / l1 composed with softmax's (diag(p) - p p^T) is applied exactly with l1 = 1.
  const float p = lane < M ? w[(int64_t)r * M + lane] : 0.f;
  float g = lane < M ? dw[(int64_t)r * M + lane] : 0.f;
  g = g - warp_sum(g * p);                                // L1-normalise backward: dp = dw - <dw, w>
  const float dot = warp_sum(g * p);
  if (lane < ld) dlogits[(int64_t)r * ld + lane] = lane < M ? p * (g - dot) : 0.f;
}

// ------------------------------------------------------------------------------------------
// sims[i,j] = sum_m (vw[j,m] * tw[i,m] / norm[i,j]) * dots[m,i,j]   (model/model.py:803-836)
// one thread per output element, consecutive threads along j (coalesced dots / sims traffic)
// ------------------------------------------------------------------------------------------
constexpr int MAXM = 32;

__global__ void __launch_bounds__(256) sims_fwd_kernel(const float* __restrict__ dots,
                                                       const float* __restrict__ tw,
                                                       const float* __restrict__ vw, int Nq, int Nv,
                                                       int M, int caps, int merge_avg,
                                                       float* __restrict__ sims) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int io = blockIdx.y;                               // output row
  if (j >= Nv) return;
  const int reps = merge_avg ? caps : 1;
  float out = 0.f;
  for (int c = 0; c < reps; ++c) {
    const int i = merge_avg ? io * caps + c : io;
    float norm = 0.f;
    for (int m = 0; m < M; ++m) norm += vw[(int64_t)j * M + m] * tw[(int64_t)i * M + m];
    if (norm == 0.f) norm = 1e-5f;                         // model.py:816
    float s = 0.f;
    for (int m = 0; m < M; ++m) {
      const float wgt = (vw[(int64_t)j * M + m] * tw[(int64_t)i * M + m]) / norm;
      s += wgt * dots[((int64_t)m * Nq + i) * Nv + j];     // sims += moe[:, :, m] * matmul(...)
    }
    out += s;
  }
  if (merge_avg && caps > 1) out /= caps;                  // th.mean over captions (model.py:829)
  sims[(int64_t)io * Nv + j] = out;
}

// ddots[m,i,j] = dS[i,j] * w[i,j,m];  dtw[i,m] = sum_j dS[i,j] * d sims_ij / d tw_im.
// One block per query row i; threads stride over j; warp+block reduce for dtw.
__global__ void __launch_bounds__(256) sims_bwd_kernel(const float* __restrict__ dsims,
                                                       const float* __restrict__ dots,
                                                       const float* __restrict__ tw,
                                                       const float* __restrict__ vw, int Nq, int Nv,
                                                       int M, int caps, int merge_avg,
                                                       float* __restrict__ ddots,
                                                       float* __restrict__ dtw) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][MAXM];
  const int i = blockIdx.x;
  const int io = merge_avg ? i / caps : i;
  const float gscale = (merge_avg && caps > 1) ? 1.0f / caps : 1.0f;
  float twi[MAXM], acc[MAXM];
  for (int m = 0; m < M; ++m) { twi[m] = tw[(int64_t)i * M + m]; acc[m] = 0.f; }
  for (int j = threadIdx.x; j < Nv; j += blockDim.x) {
    const float ds = dsims[(int64_t)io * Nv + j] * gscale;
    float norm = 0.f;
    for (int m = 0; m < M; ++m) norm += vw[(int64_t)j * M + m] * twi[m];
    const bool zero = (norm == 0.f);
    if (zero) norm = 1e-5f;
    float s = 0.f;
    for (int m = 0; m < M; ++m) {
      const float wgt = (vw[(int64_t)j * M + m] * twi[m]) / norm;
      const float dt = dots[((int64_t)m * Nq + i) * Nv + j];
      s += wgt * dt;
      ddots[((int64_t)m * Nq + i) * Nv + j] = ds * wgt;
    }
    for (int m = 0; m < M; ++m) {
      const float v = vw[(int64_t)j * M + m];
      const float dt = dots[((int64_t)m * Nq + i) * Nv + j];
      // d/dtw_im [ sum_m' tw vw D / norm ] ; norm is a constant (1e-5) where it was replaced
      acc[m] += zero ? ds * v * dt / norm : ds * (v / norm) * (dt - s);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int m = 0; m < M; ++m) {
    const float v = warp_sum(acc[m]);
    if (lane == 0) red[warp][m] = v;
  }
  __syncthreads();
  if (threadIdx.x < M) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    dtw[(int64_t)i * M + threadIdx.x] = v;
  }
}

// ------------------------------------------------------------------------------------------
// MaxMarginRankingLoss forward + gradient (model/loss.py:38-65; closed form SURVEY App. A.8)
//   element (i,j), i != j contributes  relu(m - x_ii + x_ij) + relu(m - x_jj + x_ij)
//   dL/dx_ij = [1(a>0) + 1(b>0)] / cnt ;  dL/dx_ii = -(row_i count of a>0 + col_i count of b>0) / cnt
// Streaming: each block owns a strip of rows, threads walk j with float4 loads; row/col indicator
// counts go to a workspace with (few) atomics; a finishing kernel writes the diagonal and the mean.
// ------------------------------------------------------------------------------------------
// Forward-only streaming variant (no gradient): every element is treated alike --
//   sum_ij relu(m - x_ii + x_ij) + relu(m - x_jj + x_ij)
// -- and the diagonal's contribution (2 n relu(m)) is removed in the finishing kernel, so the
// inner loop is 6 instructions per element with 8 independent 16-byte loads in flight per thread.
__global__ void __launch_bounds__(256) max_margin_fwd_kernel(const float* __restrict__ x, int n,
                                                             int rows_per_block, float margin,
                                                             float* __restrict__ ws) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = 4 * (blockIdx.x * 256 + threadIdx.x);           // n % 4 == 0 guaranteed by the host
  float lsum = 0.f;
  if (j0 < n) {
    float mj[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) mj[q] = margin - __ldg(x + (int64_t)(j0 + q) * n + (j0 + q));
    const int i_begin = blockIdx.y * rows_per_block;
    const int i_end = min(n, i_begin + rows_per_block);
    constexpr int RU = 8;
    for (int ib = i_begin; ib < i_end; ib += RU) {
      float4 xr[RU];
      float mi[RU];
#pragma unroll
      for (int r = 0; r < RU; ++r) {
        const int i = min(ib + r, i_end - 1);                      // clamp: duplicates are masked below
        mi[r] = margin - __ldg(x + (int64_t)i * n + i);
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(xr[r].x), "=f"(xr[r].y), "=f"(xr[r].z), "=f"(xr[r].w)
                     : "l"(x + (int64_t)i * n + j0));
      }
#pragma unroll
      for (int r = 0; r < RU; ++r) {
        const float w = (ib + r < i_end) ? 1.f : 0.f;
        const float xv[4] = {xr[r].x, xr[r].y, xr[r].z, xr[r].w};
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) t += fmaxf(mi[r] + xv[q], 0.f) + fmaxf(mj[q] + xv[q], 0.f);
        lsum = fmaf(w, t, lsum);
      }
    }
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(ws, t);
  }
}

// grid = (column chunks of 1024, row strips); thread t owns columns 4*(chunk*256 + t) .. +3.
__global__ void __launch_bounds__(256) max_margin_kernel(const float* __restrict__ x, int n,
                                                         int rows_per_block, float margin,
                                                         int fix_norm, float inv_cnt,
                                                         float* __restrict__ dx,
                                                         float* __restrict__ ws /* [0]=loss sum, [2..2+n) diag counts */) {
  __shared__ float red[8];
  pdl_trigger();
  pdl_wait();                                    // before the first global access (the diagonal loads below)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = 4 * (blockIdx.x * 256 + threadIdx.x);
  const bool vec = (n % 4 == 0);
  float djj[4], colcnt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q) djj[q] = (j0 + q < n) ? __ldg(x + (int64_t)(j0 + q) * n + (j0 + q)) : 0.f;
  float lsum = 0.f;
  const int i_begin = blockIdx.y * rows_per_block;
  const int i_end = min(n, i_begin + rows_per_block);
  constexpr int RU = 4;                          // rows in flight per thread (memory-level parallelism)
  for (int ib = i_begin; ib < i_end; ib += RU) {
    float dii[RU];
    float4 xr[RU];
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      const int i = ib + r;
      dii[r] = (i < i_end) ? __ldg(x + (int64_t)i * n + i) : 0.f;
      xr[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < i_end && j0 < n) {
        if (vec) {
          asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(xr[r].x), "=f"(xr[r].y), "=f"(xr[r].z), "=f"(xr[r].w)
                       : "l"(x + (int64_t)i * n + j0));
        } else {
          float* pv = reinterpret_cast<float*>(&xr[r]);
#pragma unroll
          for (int q = 0; q < 4; ++q) if (j0 + q < n) pv[q] = x[(int64_t)i * n + j0 + q];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RU; ++r) {
      const int i = ib + r;
      if (i >= i_end) break;
      const float xv[4] = {xr[r].x, xr[r].y, xr[r].z, xr[r].w};
      float g[4];
      float rowcnt = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + q;
        const bool in = j < n;
        const float a = margin - dii[r] + xv[q], b = margin - djj[q] + xv[q];
        const bool off = in && (j != i);
        const bool use = in && (off || !fix_norm);
        const float ia = (use && a > 0.f) ? 1.f : 0.f, ib2 = (use && b > 0.f) ? 1.f : 0.f;
        lsum += ia * a + ib2 * b;
        g[q] = off ? (ia + ib2) * inv_cnt : 0.f;
        if (off) { rowcnt += ia; colcnt[q] += ib2; }
      }
      if (dx) {
        if (j0 < n) {
          if (vec) *reinterpret_cast<float4*>(dx + (int64_t)i * n + j0) = make_float4(g[0], g[1], g[2], g[3]);
          else {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (j0 + q < n) dx[(int64_t)i * n + j0 + q] = g[q];
          }
        }
        rowcnt = warp_sum(rowcnt);
        if (lane == 0 && rowcnt != 0.f) atomicAdd(ws + 2 + i, rowcnt);
      }
    }
  }
  if (dx) {
#pragma unroll
    for (int q = 0; q < 4; ++q) if (colcnt[q] != 0.f) atomicAdd(ws + 2 + j0 + q, colcnt[q]);
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(ws, t);
  }
}

__global__ void max_margin_finish_kernel(int n, float inv_cnt, int fix_norm, float diag_correction,
                                         const float* __restrict__ ws, float* __restrict__ loss,
                                         float* __restrict__ dx) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *loss = (n > 1 || !fix_norm) ? (ws[0] - diag_correction) * inv_cnt : __int_as_float(0x7fc00000);
  // fix_norm == 0: the diagonal terms relu(margin) are constants w.r.t. x (x_ii cancels) -> the
  // diagonal gradient is the same sum of off-diagonal indicators.
  if (dx && i < n) dx[(int64_t)i * n + i] = -ws[2 + i] * inv_cnt;
}

// ------------------------------------------------------------------------------------------
// fused Adam over a flat buffer (torch.optim.Adam semantics, no amsgrad)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   int64_t n4, int64_t n, float lr, float b1, float b2,
                                                   float eps, float wd, int step, const uint64_t* __restrict__ ctr,
                                                   float gscale, uint16_t* __restrict__ p16, int bf16) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const float t = (float)(step + (ctr ? (int)*ctr : 0));      // device-side step counter for graph replays
    s_bc[0] = 1.f - powf(b1, t);
    s_bc[1] = sqrtf(1.f - powf(b2, t));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = reinterpret_cast<float*>(&pp); float* G = reinterpret_cast<float*>(&gg);
    float* Mm = reinterpret_cast<float*>(&mm); float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float gr = G[q] * gscale + wd * P[q];
      Mm[q] = b1 * Mm[q] + (1.f - b1) * gr;
      V[q] = b2 * V[q] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(V[q]) / bc2_sqrt + eps;
      P[q] -= (lr / bc1) * (Mm[q] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    if (p16 != nullptr) reinterpret_cast<uint2*>(p16)[i] = pack4(pp, bf16 != 0);     // the GEMMs' 16-bit weight copy
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // scalar tail
  if (blockIdx.x == 0) {
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float gr = g[i] * gscale + wd * p[i];
      m[i] = b1 * m[i] + (1.f - b1) * gr;
      v[i] = b2 * v[i] + (1.f - b2) * gr * gr;
      p[i] -= (lr / bc1) * (m[i] / (sqrtf(v[i]) / bc2_sqrt + eps));
      if (p16 != nullptr) p16[i] = pack1(p[i], bf16 != 0);
    }
  }
}


// ------------------------------------------------------------------------------------------
// Retrieval ranks (model/metric.py:26-230): position of the ground truth in the row sorted by
// DESCENDING similarity, ties averaged = #(strictly better) + (#equal - 1) / 2.  Pure counting on
// the fp32 values (the reference compares the negated floats): exact, no sort.
// ------------------------------------------------------------------------------------------
// t2v: one warp per query (caption) i; ground-truth video = i / caps.
__global__ void __launch_bounds__(256) ranks_t2v_kernel(const float* __restrict__ sims, int Nq, int Nv, int caps,
                                                        float* __restrict__ ranks) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= Nq) return;
  const float* row = sims + (int64_t)i * Nv;
  const float g = row[i / caps];
  int better = 0, equal = 0;
  for (int j = lane; j < Nv; j += 32) {
    const float x = row[j];
    better += x > g;
    equal += x == g;
  }
  better = warp_sum_int(better);
  equal = warp_sum_int(equal);
  if (lane == 0) ranks[i] = (float)better + 0.5f * (float)(equal - 1);
}

// v2t: one block per video v; the candidates are ALL captions (column v of sims), captions whose mask is
// 0 count as infinitely far (metric.py:187-192); result = best rank over v's own unmasked captions
// (inf when it has none, as np.inf in the reference).
__global__ void __launch_bounds__(256) ranks_v2t_kernel(const float* __restrict__ sims, const int32_t* __restrict__ valid,
                                                        int Nq, int Nv, int caps, float* __restrict__ ranks) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_better[8], s_equal[8];
  const int v = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float best = INFINITY;
  for (int own = v * caps; own < (v + 1) * caps; ++own) {
    if (valid != nullptr && valid[own] == 0) continue;      // block-uniform
    const float g = sims[(int64_t)own * Nv + v];
    int better = 0, equal = 0;
    for (int c = threadIdx.x; c < Nq; c += blockDim.x) {
      if (valid != nullptr && valid[c] == 0) continue;      // at distance 1e8: never better, never equal
      const float x = sims[(int64_t)c * Nv + v];
      better += x > g;
      equal += x == g;
    }
    better = warp_sum_int(better);
    equal = warp_sum_int(equal);
    __syncthreads();
    if (lane == 0) { s_better[warp] = better; s_equal[warp] = equal; }
    __syncthreads();
    int tb = 0, te = 0;
    for (int w = 0; w < 8; ++w) { tb += s_better[w]; te += s_equal[w]; }
    best = fminf(best, (float)tb + 0.5f * (float)(te - 1));
  }
  if (threadIdx.x == 0) ranks[v] = best;
}

}  // namespace
}  // namespace mmt

using namespace mmt;

extern "C" {

int mmt_geu_gate_fwd(const float* X, const float* G, const float* bn_w, const float* bn_b,
                     float* run_mean, float* run_var, int32_t R, int32_t M, int32_t d,
                     int32_t training, float momentum, float bn_eps, float* E, float* Y,
                     float* bn_mean, float* bn_rstd, float* inv_n1, float* inv_n2, void* stream) {
  MMT_ARG_CHECK(X && G && bn_w && bn_b && run_mean && run_var && E && Y && bn_mean && bn_rstd &&
                inv_n1 && inv_n2, MMT_E_ARG, "mmt_geu_gate_fwd: null pointer");
  MMT_ARG_CHECK(R > 0 && M > 0, MMT_E_SHAPE, "mmt_geu_gate_fwd: bad shape R=%d M=%d", R, M);
  CHECK_D(d);
  const int C = M * d;
  launch_pdl(bn_stats_kernel, dim3((C + 31) / 32), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, 
      G, R, C, training, momentum, bn_eps, run_mean, run_var, bn_mean, bn_rstd);
  MMT_LAUNCH_CHECK("bn_stats");
  const int64_t rows = (int64_t)R * M;
  DISPATCH_VEC(d, (launch_pdl(geu_gate_fwd_kernel<V>, dim3((int)((rows + WARPS - 1) / WARPS)), dim3(WARPS * 32), 0, (cudaStream_t)stream, 
      X, G, bn_w, bn_b, bn_mean, bn_rstd, R, M, E, Y, inv_n1, inv_n2)));
  MMT_LAUNCH_CHECK("geu_gate_fwd");
  return 0;
}

int mmt_geu_gate_bwd(const float* dE, const float* X, const float* G, const float* Y,
                     const float* E, const float* bn_w, const float* bn_b, const float* bn_mean,
                     const float* bn_rstd, const float* inv_n1, const float* inv_n2, int32_t R,
                     int32_t M, int32_t d, int32_t training, float* dX, float* dG, float* dbn_w,
                     float* dbn_b, void* stream) {
  MMT_ARG_CHECK(dE && X && G && Y && E && bn_w && bn_b && bn_mean && bn_rstd && inv_n1 && inv_n2 &&
                dX && dG && dbn_w && dbn_b, MMT_E_ARG, "mmt_geu_gate_bwd: null pointer");
  CHECK_D(d);
  const int64_t rows = (int64_t)R * M;
  DISPATCH_VEC(d, (launch_pdl(geu_gate_bwd_rows_kernel<V>, dim3((int)((rows + WARPS - 1) / WARPS)), dim3(WARPS * 32), 0, (cudaStream_t)stream, 
      dE, X, G, Y, E, bn_w, bn_b, bn_mean, bn_rstd, inv_n1, inv_n2, R, M, dX, dG)));
  MMT_LAUNCH_CHECK("geu_gate_bwd_rows");
  const int C = M * d;
  launch_pdl(bn_bwd_kernel, dim3((C + 31) / 32), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, dG, G, bn_w, bn_mean, bn_rstd, R, C, training, dbn_w, dbn_b);
  MMT_LAUNCH_CHECK("bn_bwd");
  return 0;
}

int mmt_moe_softmax_fwd(const float* logits, int32_t R, int32_t M, int32_t ld, float* w, void* stream) {
  MMT_ARG_CHECK(logits && w, MMT_E_ARG, "mmt_moe_softmax_fwd: null pointer");
  MMT_ARG_CHECK(M >= 1 && M <= 32 && R > 0 && ld >= M && ld <= 32, MMT_E_SHAPE, "mmt_moe_softmax_fwd: M=%d ld=%d must satisfy 1 <= M <= ld <= 32", M, ld);
  launch_pdl(moe_softmax_fwd_kernel, dim3((R + 7) / 8), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, logits, R, M, ld, w);
  MMT_LAUNCH_CHECK("moe_softmax_fwd");
  return 0;
}

int mmt_moe_softmax_bwd(const float* dw, const float* w, int32_t R, int32_t M, int32_t ld,
                        float* dlogits, void* stream) {
  MMT_ARG_CHECK(dw && w && dlogits, MMT_E_ARG, "mmt_moe_softmax_bwd: null pointer");
  MMT_ARG_CHECK(M >= 1 && M <= 32 && R > 0 && ld >= M && ld <= 32, MMT_E_SHAPE, "mmt_moe_softmax_bwd: M=%d ld=%d must satisfy 1 <= M <= ld <= 32", M, ld);
  launch_pdl(moe_softmax_bwd_kernel, dim3((R + 7) / 8), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, dw, w, R, M, ld, dlogits);
  MMT_LAUNCH_CHECK("moe_softmax_bwd");
  return 0;
}

int mmt_sims_combine_fwd(const float* dots, const float* tw, const float* vw, int32_t Nq,
                         int32_t Nv, int32_t M, int32_t caps, int32_t merge_avg, float* sims,
                         void* stream) {
  MMT_ARG_CHECK(dots && tw && vw && sims, MMT_E_ARG, "mmt_sims_combine_fwd: null pointer");
  MMT_ARG_CHECK(M >= 1 && M <= MAXM && caps >= 1 && Nq == Nv * caps && Nv > 0, MMT_E_SHAPE,
                "mmt_sims_combine_fwd: bad shape Nq=%d Nv=%d M=%d caps=%d", Nq, Nv, M, caps);
  const int rows_out = merge_avg ? Nv : Nq;
  dim3 grid((Nv + 255) / 256, rows_out);
  MMT_ARG_CHECK(rows_out <= 65535, MMT_E_SHAPE, "mmt_sims_combine_fwd: %d output rows > 65535 (chunk the queries)", rows_out);
  launch_pdl(sims_fwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, dots, tw, vw, Nq, Nv, M, caps, merge_avg, sims);
  MMT_LAUNCH_CHECK("sims_fwd");
  return 0;
}

int mmt_sims_combine_bwd(const float* dsims, const float* dots, const float* tw, const float* vw,
                         int32_t Nq, int32_t Nv, int32_t M, int32_t caps, int32_t merge_avg,
                         float* ddots, float* dtw, void* stream) {
  MMT_ARG_CHECK(dsims && dots && tw && vw && ddots && dtw, MMT_E_ARG, "mmt_sims_combine_bwd: null pointer");
  MMT_ARG_CHECK(M >= 1 && M <= MAXM && caps >= 1 && Nq == Nv * caps && Nv > 0, MMT_E_SHAPE,
                "mmt_sims_combine_bwd: bad shape Nq=%d Nv=%d M=%d caps=%d", Nq, Nv, M, caps);
  launch_pdl(sims_bwd_kernel, dim3(Nq), dim3(256), 0, (cudaStream_t)stream, dsims, dots, tw, vw, Nq, Nv, M, caps, merge_avg, ddots, dtw);
  MMT_LAUNCH_CHECK("sims_bwd");
  return 0;
}

int mmt_max_margin_fwd_bwd(const float* x, int32_t n, float margin, int32_t fix_norm, float* loss,
                           float* dx, float* workspace, void* stream) {
  MMT_ARG_CHECK(x && loss && workspace, MMT_E_ARG, "mmt_max_margin_fwd_bwd: null pointer");
  MMT_ARG_CHECK(n >= 1, MMT_E_SHAPE, "mmt_max_margin_fwd_bwd: n=%d", n);
  cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)(n + 2), (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_status(e, "max_margin memset");
  const double cnt = fix_norm ? 2.0 * n * (double)(n - 1) : 2.0 * n * (double)n;
  const float inv_cnt = cnt > 0 ? (float)(1.0 / cnt) : 0.f;
  const int chunks = (n + 1023) / 1024;
  int strips = (num_sms() * 8 + chunks - 1) / chunks;
  if (strips > n) strips = n;
  if (strips > 65535) strips = 65535;
  const int rows_per_block = (n + strips - 1) / strips;
  strips = (n + rows_per_block - 1) / rows_per_block;
  float diag_correction = 0.f;
  if (dx == nullptr && n % 4 == 0 && ((uintptr_t)x % 16) == 0) {
    // forward only: uniform streaming loop; the n diagonal elements each contributed 2 relu(margin)
    launch_pdl(max_margin_fwd_kernel, dim3(dim3(chunks, strips)), dim3(256), 0, (cudaStream_t)stream, x, n, rows_per_block, margin, workspace);
    if (fix_norm) diag_correction = 2.f * n * (margin > 0.f ? margin : 0.f);
  } else {
    launch_pdl(max_margin_kernel, dim3(dim3(chunks, strips)), dim3(256), 0, (cudaStream_t)stream, x, n, rows_per_block, margin, fix_norm, inv_cnt, dx, workspace);
  }
  MMT_LAUNCH_CHECK("max_margin");
  launch_pdl(max_margin_finish_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, n, inv_cnt, fix_norm, diag_correction, workspace, loss, dx);
  MMT_LAUNCH_CHECK("max_margin_finish");
  return 0;
}

int mmt_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                  void* stream) {
  MMT_ARG_CHECK(p && g && m && v, MMT_E_ARG, "mmt_adam_step: null pointer");
  MMT_ARG_CHECK(step >= 1, MMT_E_ARG, "mmt_adam_step: step=%d must be >= 1", step);
  MMT_ARG_CHECK(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                ((uintptr_t)v % 16 == 0), MMT_E_ALIGN, "mmt_adam_step: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(adam_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n4, n, lr, beta1, beta2, eps, weight_decay, step, g_step_ctr, grad_scale, (uint16_t*)nullptr, 0);
  MMT_LAUNCH_CHECK("adam");
  return 0;
}

int mmt_adam16_step(float* p, const float* g, float* m, float* v, void* p16, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, const uint64_t* step_ctr,
                    float grad_scale, int32_t dtype, void* stream) {
  MMT_ARG_CHECK(p && g && m && v, MMT_E_ARG, "mmt_adam16_step: null pointer");
  MMT_ARG_CHECK(step >= 1 || step_ctr != nullptr, MMT_E_ARG, "mmt_adam16_step: step=%d must be >= 1", step);
  MMT_ARG_CHECK(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                ((uintptr_t)v % 16 == 0) && ((uintptr_t)p16 % 8 == 0), MMT_E_ALIGN, "mmt_adam16_step: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(adam_kernel, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n4, n, lr, beta1, beta2,
             eps, weight_decay, step, step_ctr, grad_scale, reinterpret_cast<uint16_t*>(p16), dtype == MMT_DT_BF16 ? 1 : 0);
  MMT_LAUNCH_CHECK("adam16");
  return 0;
}

int mmt_retrieval_ranks(const float* sims, const int32_t* valid, int32_t Nq, int32_t Nv, int32_t v2t,
                        float* ranks, void* stream) {
  MMT_ARG_CHECK(sims && ranks, MMT_E_ARG, "mmt_retrieval_ranks: null pointer");
  MMT_ARG_CHECK(Nv >= 1 && Nq >= Nv && Nq % Nv == 0, MMT_E_SHAPE, "mmt_retrieval_ranks: Nq=%d must be a multiple of Nv=%d", Nq, Nv);
  const int caps = Nq / Nv;
  if (v2t) {
    launch_pdl(ranks_v2t_kernel, dim3(Nv), dim3(256), 0, (cudaStream_t)stream, sims, valid, Nq, Nv, caps, ranks);
  } else {
    launch_pdl(ranks_t2v_kernel, dim3((Nq + 7) / 8), dim3(256), 0, (cudaStream_t)stream, sims, Nq, Nv, caps, ranks);
  }
  MMT_LAUNCH_CHECK("retrieval_ranks");
  return 0;
}

}  // extern "C"
