// fp32 CUDA-core GEMM with generic operand strides (MMT_PREC_FP32).
//
// Used for (a) the exact-fp32 mode of every linear layer (parity to ~1e-6 against the oracle),
// (b) permanently for the small / oddly-shaped products of the text head and the similarity
// matrix, where "ranking indices bit-exact at the similarity boundary" needs true fp32 FMAs
// (SURVEY.md §7 hard parts), and (c) as the on-GPU cross-check of the tcgen05 path in tests.
//
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile per thread, register-staged double
// buffering.  Operands are fetched with element strides, so no transposes are ever materialised.
#include "common.cuh"

namespace mmt {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PAD = 4;

struct GemmArgs {
  mmt_gemm_desc d;
  int split_k;     // CTAs along K per output tile (atomic reduction into a zeroed C when > 1)
  int k_chunk;     // K elements per split, multiple of BK
};

__device__ __forceinline__ int64_t a_off(const mmt_gemm_desc& d, int m, int k) {
  if (d.a_kb > 0) return (int64_t)m * d.a_ms + (int64_t)(k / d.a_kb) * d.a_kbs + (int64_t)(k % d.a_kb) * d.a_ks;
  return (int64_t)m * d.a_ms + (int64_t)k * d.a_ks;
}
__device__ __forceinline__ int64_t c_off(const mmt_gemm_desc& d, int m) {
  if (d.c_mb > 0) return (int64_t)(m / d.c_mb) * d.c_mbs + (int64_t)(m % d.c_mb) * d.c_ms;
  return (int64_t)m * d.c_ms;
}

// A_KFAST: consecutive threads walk k (operand contiguous along k); else they walk m / n.
template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(NT) sgemm_kernel(const GemmArgs args) {
  pdl_trigger();
  pdl_wait();
  const mmt_gemm_desc& d = args.d;
  __shared__ float As[2][BK][BM + PAD];
  __shared__ float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int z = blockIdx.z / args.split_k;
  const int ksplit = blockIdx.z % args.split_k;
  const int k_begin = ksplit * args.k_chunk;
  const int k_end = min(d.K, k_begin + args.k_chunk);
  const int z0 = z / d.batch_inner, z1 = z % d.batch_inner;
  const float* __restrict__ A = d.A + z0 * d.a_bs0 + z1 * d.a_bs1;
  const float* __restrict__ B = d.B + z0 * d.b_bs0 + z1 * d.b_bs1;
  const int64_t c_base = z0 * d.c_bs0 + z1 * d.c_bs1;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  // loader mapping: 128x16 elements per operand per tile = 2048 / 256 threads = 8 each
  float ra[8], rb[8];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int e = tid + i * NT;
      int mm, kk;
      if (A_KFAST) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int m = m0 + mm, k = k0 + kk;
      ra[i] = (m < d.M && k < k_end) ? __ldg(A + a_off(d, m, k)) : 0.f;
      int nn;
      if (B_KFAST) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int n = n0 + nn;
      k = k0 + kk;
      rb[i] = (n < d.N && k < k_end) ? __ldg(B + (int64_t)n * d.b_ns + (int64_t)k * d.b_ks) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int e = tid + i * NT;
      int mm, kk, nn;
      if (A_KFAST) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      As[buf][kk][mm] = ra[i];
      if (B_KFAST) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      Bs[buf][kk][nn] = rb[i];
    }
  };

  const int tx = tid % 16, ty = tid / 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (k_end - k_begin + BK - 1) / BK;
  if (nk <= 0) return;
  load_tile(k_begin);
  store_tile(0);
  __syncthreads();
  for (int t = 0; t < nk; ++t) {
    const int buf = t & 1;
    if (t + 1 < nk) load_tile(k_begin + (t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= d.M) continue;
    const int64_t row = c_base + c_off(d, m);
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int nb = n0 + jh * 64 + tx * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nb + j;
        if (n >= d.N) continue;
        float v = acc[i][jh * 4 + j] * d.alpha;
        if (args.split_k > 1) {
          if (ksplit == 0) {
            if (d.bias) v += __ldg(d.bias + z * d.bias_bs + n);
            if (d.add) v += d.add[row + n];
          }
          atomicAdd(d.C + row + n, v);
          continue;
        }
        if (d.bias) v += __ldg(d.bias + z * d.bias_bs + n);
        if (d.add) v += d.add[row + n];
        if (d.epilogue == MMT_EPI_GELU) {
          d.aux[row + n] = v;
          v = gelu_erf(v);
        } else if (d.epilogue == MMT_EPI_DGELU) {
          v *= dgelu_erf(d.aux[row + n]);
        }
        d.C[row + n] = v;
      }
    }
  }
}

// 32 x 32 x 16 tiles for small outputs (the per-expert similarity dot products at training batch
// sizes: M = N = 64, K = 512, 7 batches): 16x more CTAs than the 128-tile kernel, which for these
// shapes is pure latency.  One thread = 2 x 2 outputs; fp32 FMAs in k order.
constexpr int SB = 32;
__global__ void __launch_bounds__(NT) sgemm_small_kernel(const GemmArgs args) {
  pdl_trigger();
  pdl_wait();
  const mmt_gemm_desc& d = args.d;
  __shared__ float As[BK][SB + 1];
  __shared__ float Bs[BK][SB + 1];
  const int tid = threadIdx.x;
  const int z = blockIdx.z;
  const int z0 = z / d.batch_inner, z1 = z % d.batch_inner;
  const float* __restrict__ A = d.A + z0 * d.a_bs0 + z1 * d.a_bs1;
  const float* __restrict__ B = d.B + z0 * d.b_bs0 + z1 * d.b_bs1;
  const int64_t c_base = z0 * d.c_bs0 + z1 * d.c_bs1;
  const int m0 = blockIdx.y * SB, n0 = blockIdx.x * SB;
  const int tx = tid % 16, ty = tid / 16;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < d.K; k0 += BK) {
    // 32 x 16 elements per operand = 512 / 256 threads = 2 each; k fastest when contiguous along k
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = tid + i * NT;
      int mm, kk;
      if (d.a_ks == 1) { kk = e % BK; mm = e / BK; } else { mm = e % SB; kk = e / SB; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < d.M && k < d.K) ? __ldg(A + a_off(d, m, k)) : 0.f;
      int nn;
      if (d.b_ks == 1) { kk = e % BK; nn = e / BK; } else { nn = e % SB; kk = e / SB; }
      const int n = n0 + nn, k2 = k0 + kk;
      Bs[kk][nn] = (n < d.N && k2 < d.K) ? __ldg(B + (int64_t)n * d.b_ns + (int64_t)k2 * d.b_ks) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float a0 = As[k][ty], a1 = As[k][ty + 16], b0 = Bs[k][tx], b1 = Bs[k][tx + 16];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= d.M) continue;
    const int64_t row = c_base + c_off(d, m);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= d.N) continue;
      float v = acc[i][j] * d.alpha;
      if (d.bias) v += __ldg(d.bias + z * d.bias_bs + n);
      if (d.add) v += d.add[row + n];
      if (d.epilogue == MMT_EPI_GELU) { d.aux[row + n] = v; v = gelu_erf(v); }
      else if (d.epilogue == MMT_EPI_DGELU) v *= dgelu_erf(d.aux[row + n]);
      d.C[row + n] = v;
    }
  }
}

// Small outputs whose operands are both contiguous along k (the similarity dot products: rows of
// normalised embeddings): one WARP per 8 x 8 output block.  Each lane walks k with float4 loads
// (lane, lane + 32, ... float4s of a row: fully coalesced), keeps 64 partial sums in registers and
// the warp reduces them at the end.  28 CTAs of the tiled kernel spent 45 us on 29 MFLOP (latency);
// here M*N/64 warps share the work and every row read is a 128-byte-per-lane-group stream.
constexpr int DB = 8;
__global__ void __launch_bounds__(128) sdot_block_kernel(const mmt_gemm_desc d, int blocks_n, int blocks_per_z) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);          // global warp = (z, block_m, block_n)
  const int z = gw / blocks_per_z;
  if (z >= d.batch) return;
  const int rem = gw - z * blocks_per_z;
  const int m0 = (rem / blocks_n) * DB, n0 = (rem % blocks_n) * DB;
  const int z0 = z / d.batch_inner, z1 = z % d.batch_inner;
  const float* __restrict__ A = d.A + z0 * d.a_bs0 + z1 * d.a_bs1;
  const float* __restrict__ B = d.B + z0 * d.b_bs0 + z1 * d.b_bs1;
  float acc[DB][DB];
#pragma unroll
  for (int i = 0; i < DB; ++i)
#pragma unroll
    for (int j = 0; j < DB; ++j) acc[i][j] = 0.f;
  for (int k = lane * 4; k < d.K; k += 128) {
    float4 b[DB];
#pragma unroll
    for (int j = 0; j < DB; ++j)
      b[j] = (n0 + j < d.N) ? __ldg(reinterpret_cast<const float4*>(B + (int64_t)(n0 + j) * d.b_ns + k))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < DB; ++i) {
      const float4 a = (m0 + i < d.M) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)(m0 + i) * d.a_ms + k))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < DB; ++j)
        acc[i][j] = fmaf(a.w, b[j].w, fmaf(a.z, b[j].z, fmaf(a.y, b[j].y, fmaf(a.x, b[j].x, acc[i][j]))));
    }
  }
  // Halving reduction: after the exchange over distance 16 each lane keeps half of the values, ...;
  // 5 rounds leave lane l with the totals of outputs 2l and 2l+1 (62 shuffles instead of 320).
  float v[DB * DB];
#pragma unroll
  for (int i = 0; i < DB; ++i)
#pragma unroll
    for (int j = 0; j < DB; ++j) v[i * DB + j] = acc[i][j];
  int n = DB * DB;
#pragma unroll
  for (int sh = 16; sh >= 1; sh >>= 1) {
    n >>= 1;
    const bool upper = (lane & sh) != 0;
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      if (t < n) {
        const float send = upper ? v[t] : v[t + n];              // give away the half this lane drops
        const float keep = upper ? v[t + n] : v[t];
        v[t] = keep + __shfl_xor_sync(0xffffffffu, send, sh);
      }
    }
  }
  // lane l now holds outputs whose index has bit pattern: bit5..1 = (l&16 ? 1:0),(l&8),(l&4),(l&2),(l&1) as the
  // successive halves chosen -> index = hi-bits from the lane, low bit t in {0,1}
  const int base = ((lane >> 4) & 1) * 32 + ((lane >> 3) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 +
                   (lane & 1) * 2;
  const int64_t c_base = z0 * d.c_bs0 + z1 * d.c_bs1;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int idx = base + t, i = idx / DB, j = idx % DB;
    const int m = m0 + i, nn = n0 + j;
    if (m < d.M && nn < d.N) {
      float o = v[t] * d.alpha;
      if (d.bias) o += __ldg(d.bias + z * d.bias_bs + nn);
      d.C[c_base + c_off(d, m) + nn] = o;
    }
  }
}

}  // namespace

int gemm_simt(const mmt_gemm_desc& d, cudaStream_t stream) {
  // up to 256 x 256 outputs (the data-parallel head at global batch 128 / 256): the 128-tile kernel
  // would run 7-28 CTAs there
  if (d.M <= 256 && d.N <= 256 && d.a_ks == 1 && d.b_ks == 1 && d.a_kb == 0 && d.K >= 128 && (d.K & 3) == 0 &&
      ((d.a_ms | d.b_ns | d.a_bs0 | d.a_bs1 | d.b_bs0 | d.b_bs1) & 3) == 0 &&
      ((((uintptr_t)d.A | (uintptr_t)d.B)) & 15) == 0 && d.epilogue == MMT_EPI_NONE && d.add == nullptr) {
    const int bm = (d.M + DB - 1) / DB, bn = (d.N + DB - 1) / DB;   // one warp per 8 x 8 outputs
    const int warps = bm * bn * d.batch;
    launch_pdl(sdot_block_kernel, dim3((warps + 3) / 4), dim3(128), 0, stream, d, bn, bm * bn);
    MMT_LAUNCH_CHECK("sdot_block_kernel");
    return 0;
  }
  if (d.M <= 96 && d.N <= 96) {                     // small outputs: 32 x 32 tiles
    GemmArgs small{d, 1, 0};
    dim3 g((d.N + SB - 1) / SB, (d.M + SB - 1) / SB, d.batch);
    launch_pdl(sgemm_small_kernel, dim3(g), dim3(NT), 0, stream, small);
    MMT_LAUNCH_CHECK("sgemm_small_kernel");
    return 0;
  }
  GemmArgs args{d, 1, ((d.K + BK - 1) / BK) * BK};
  const int tiles = ((d.N + BN - 1) / BN) * ((d.M + BM - 1) / BM);
  // weight-gradient shape: few output tiles, long K, dense un-batched C that is not also `add`
  if ((d.flags & MMT_GEMM_SPLIT_K) && d.batch == 1 && d.c_mb == 0 && d.c_ms == d.N && d.epilogue == MMT_EPI_NONE &&
      d.add != d.C && tiles * 2 <= num_sms() && d.K >= 32 * BK) {
    int split = (2 * num_sms() + tiles - 1) / tiles;
    const int max_split = d.K / (8 * BK);
    if (split > max_split) split = max_split;
    if (split > 1) {
      args.k_chunk = (((d.K + split - 1) / split + BK - 1) / BK) * BK;
      args.split_k = (d.K + args.k_chunk - 1) / args.k_chunk;
      cudaError_t e = cudaMemsetAsync(d.C, 0, sizeof(float) * (size_t)d.M * d.N, stream);
      if (e != cudaSuccess) return cuda_status(e, "sgemm split-K memset");
    }
  }
  dim3 grid((d.N + BN - 1) / BN, (d.M + BM - 1) / BM, d.batch * args.split_k);
  const bool a_kfast = (d.a_ks == 1);
  const bool b_kfast = (d.b_ks == 1);
  if (a_kfast && b_kfast) launch_pdl(sgemm_kernel<true, true>, dim3(grid), dim3(NT), 0, stream, args);
  else if (a_kfast) launch_pdl(sgemm_kernel<true, false>, dim3(grid), dim3(NT), 0, stream, args);
  else if (b_kfast) launch_pdl(sgemm_kernel<false, true>, dim3(grid), dim3(NT), 0, stream, args);
  else launch_pdl(sgemm_kernel<false, false>, dim3(grid), dim3(NT), 0, stream, args);
  MMT_LAUNCH_CHECK("sgemm_kernel");
  return 0;
}

}  // namespace mmt
