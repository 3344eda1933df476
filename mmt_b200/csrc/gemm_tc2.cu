// Persistent tcgen05 TF32 GEMM on single CTAs: one CTA per SM walks a static list of 128 x 256 (or
// 128 x 128) output tiles, also across (batch_outer, batch_inner) work items addressed through rank-4
// tensor maps -- the batched attention products and the problems too small for the CTA-pair kernel.
//
//   warp 0    TMA producer     4-stage ring of {A 128x32, B 256x32} fp32 tiles (48 KB / stage);
//                              6 stages of 32 KB for the 128-wide tile
//   warp 1    MMA issuer       tcgen05.mma kind::tf32 M=128 N=256 K=8; the 512 TMEM columns hold TWO
//                              accumulators so tile i+1's main loop overlaps tile i's epilogue
//   warps 2-5 epilogue         tcgen05.ld -> per-warp shared-memory transpose -> fused
//                              bias / residual / erf-GELU / GELU' and fully coalesced 128-bit
//                              global loads/stores (each store instruction covers 4 complete
//                              128-byte row segments); plain epilogues (alpha, bias, column sums)
//                              instead write the lane's row to a 128-byte-swizzled 32 x 32 tile and
//                              leave through ONE TMA store per tile (see the epilogue)
// Tiles are ordered m-fastest so the CTAs that run together share the same B (weight) tile in L2.
// Weight-gradient shapes (few tiles, K = B*S) are split along K into (tile, k-range) work items
// whose epilogue reduces with red.global.add.v4.f32 into a zeroed C.
#include <cstdlib>

#include "tc_ptx.cuh"

namespace mmt {
namespace {
using namespace tc;

constexpr int BM = 128, BK = 32, UMMA_K = 8;
constexpr int NUM_THREADS = 192;
constexpr uint32_t A_BYTES = BM * BK * 4;
constexpr int STG_PITCH = 36;                         // floats; 16 B aligned rows, conflict-free v4 phases
constexpr uint32_t STG_BYTES_PER_WARP = 5120;         // 32 x 36 floats (4608 B), rounded up so every warp's tile is 1024 B aligned

struct Tc2Args {
  mmt_gemm_desc d;
  int c_tma;                                            // plain epilogue: C tiles leave through TMA stores
  int num_m_tiles, num_n_tiles;
  int split_k, kb_per_split, num_kb;
};

// BN = 256: 4-stage ring of 48 KB; BN = 128 (narrow / batched attention problems): 6 stages of 32 KB.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_b,
                                                                  const __grid_constant__ CUtensorMap map_c,
                                                                  const Tc2Args args) {
  constexpr int STAGES = BN == 256 ? 4 : 6;
  constexpr uint32_t B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* staging = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 4 * STG_BYTES_PER_WARP);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;       // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const mmt_gemm_desc& d = args.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = args.num_m_tiles * args.num_n_tiles;
  const int num_work = num_tiles * args.split_k * d.batch;      // split-K and batching are exclusive

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                           // prologue done: now wait for the producer grid

  // work item -> (batch z, m0, n0, k-block range)
  auto decode = [&](int w, int& z, int& m0, int& n0, int& kb0, int& nkb) {
    z = w / (num_tiles * args.split_k);
    w -= z * num_tiles * args.split_k;
    const int tile = w / args.split_k, ks = w % args.split_k;
    n0 = (tile / args.num_m_tiles) * BN;
    m0 = (tile % args.num_m_tiles) * BM;
    kb0 = ks * args.kb_per_split;
    nkb = min(args.num_kb, kb0 + args.kb_per_split) - kb0;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t g = 0;                                  // global k-block counter (ring position)
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int z, m0, n0, kb0, nkb;
        decode(w, z, m0, n0, kb0, nkb);
        const int z0 = z / d.batch_inner, z1 = z % d.batch_inner;
        for (int i = 0; i < nkb; ++i, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = (kb0 + i) * BK;
          if (!A_MN) {
            tma_load_4d(sa, &map_a, &full_bar[s], k0, m0, z1, z0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_4d(sa + j * (BK * 128), &map_a, &full_bar[s], m0 + 32 * j, k0, z1, z0);
          }
          if (!B_MN) {
            tma_load_4d(sb, &map_b, &full_bar[s], k0, n0, z1, z0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_4d(sb + j * (BK * 128), &map_b, &full_bar[s], n0 + 32 * j, k0, z1, z0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(BM, BN, A_MN, B_MN);
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, A_SBO = A_MN ? 512 : 1024, A_STEP = A_MN ? 1024 : UMMA_K * 4;
      constexpr uint32_t B_LBO = B_MN ? BK * 128 : 16, B_SBO = B_MN ? 512 : 1024, B_STEP = B_MN ? 1024 : UMMA_K * 4;
      constexpr uint32_t A_LT = A_MN ? 1 : 2, B_LT = B_MN ? 1 : 2;
      uint32_t g = 0;
      int it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        int z, m0, n0, kb0, nkb;
        decode(w, z, m0, n0, kb0, nkb);
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);        // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
        for (int i = 0; i < nkb; ++i, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(sa + k * A_STEP, A_LBO, A_SBO, A_LT);
            const uint64_t db = make_smem_desc(sb + k * B_STEP, B_LBO, B_SBO, B_LT);
            umma_tf32(acc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                               // TMEM lane quarter == output rows 32q..32q+31
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(staging) + q * STG_BYTES_PER_WARP);
    const bool vec_ok = ((d.c_ms & 3) == 0) && (((d.c_bs0 | d.c_bs1 | d.bias_bs) & 3) == 0) &&
                        ((((uintptr_t)d.C | (uintptr_t)d.bias | (uintptr_t)d.add | (uintptr_t)d.aux) & 15) == 0);
    const int sub_r = lane >> 3;                          // store phase: row within a group of 4
    const int sub_c = (lane & 7) * 4;                     // store phase: first of this lane's 4 columns
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      int z, m0, n0, kb0, nkb;
      decode(w, z, m0, n0, kb0, nkb);
      const int buf = it & 1;
      const bool lead = (kb0 == 0);
      const int64_t zoff = (int64_t)(z / d.batch_inner) * d.c_bs0 + (int64_t)(z % d.batch_inner) * d.c_bs1;
      const float* bias = d.bias ? d.bias + (int64_t)z * d.bias_bs : nullptr;
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nb = n0 + c * 32;
        if (nb >= d.N) break;                             // warp-uniform
        float v[32];
        tmem_ld32(acc + (uint32_t)(c * 32), v);
        if (args.c_tma) {
          // Plain epilogue (C = alpha * acc [+ bias], optional column sums): the lane's row goes to the
          // warp's [32 x 32] staging tile in the 128-byte swizzle the tensor map expects and ONE TMA
          // store writes the tile (rows >= M and columns >= N are clipped by the map).  No transposed
          // read-back, no per-row address arithmetic: the batched attention products (K = 128 / 218)
          // were bound by exactly that epilogue work on four warps.
          const uint32_t stg_u = smem_u32(stg);
          if (bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], d.alpha, (nb + j < d.N) ? __ldg(bias + nb + j) : 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= d.alpha;
          }
          if (lane == 0) bulk_wait_read0();                   // the previous store has read the tile
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(stg) + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&map_c, stg_u, nb, m0 + q * 32, z % d.batch_inner, z / d.batch_inner);
            bulk_commit();
          }
          if (d.colsum != nullptr && nb + lane < d.N) {        // lane = column: sum this warp's valid rows
            const int rows_ok = min(32, d.M - (m0 + q * 32));
            float cs = 0.f;
            for (int r = 0; r < rows_ok; ++r)
              cs += *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(stg) + r * 128 +
                                                    (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
            atomicAdd(d.colsum + (int64_t)(z % d.batch_inner) * d.colsum_bs + nb + lane, cs);
          }
          continue;
        }
        // phase 1: this lane's row -> warp-private staging (row pitch 36 floats)
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + lane * STG_PITCH + j) =
              make_float4(v[j] * d.alpha, v[j + 1] * d.alpha, v[j + 2] * d.alpha, v[j + 3] * d.alpha);
        __syncwarp();
        // phase 2: coalesced: lanes 8r..8r+7 cover one 128-byte row segment
        const int col = nb + sub_c;
        const bool full = vec_ok && (col + 4 <= d.N);
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias && (lead || args.split_k == 1)) {
#pragma unroll
          for (int t = 0; t < 4; ++t) if (col + t < d.N) bv[t] = bias[col + t];
        }
        // (a) gather this lane's 8 row-segments (registers), (b) one batch of independent loads,
        // (c) math on 32 independent values (ILP hides the ALU/MUFU latency that four epilogue
        // warps per SM cannot hide with thread-level parallelism), (d) stores.
        float4 o[8];
        bool ok[8];
        int64_t off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + sub_r;
          const int m = m0 + q * 32 + rl;
          ok[i] = (m < d.M) && (col < d.N);
          off[i] = zoff + (int64_t)m * d.c_ms + col;
          o[i] = *reinterpret_cast<const float4*>(stg + rl * STG_PITCH + sub_c);
          o[i].x += bv[0]; o[i].y += bv[1]; o[i].z += bv[2]; o[i].w += bv[3];
        }
        if (full) {
          if (d.add && (lead || args.split_k == 1)) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ok[i] ? *reinterpret_cast<const float4*>(d.add + off[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { o[i].x += a[i].x; o[i].y += a[i].y; o[i].z += a[i].z; o[i].w += a[i].w; }
          }
          if (args.split_k > 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (ok[i]) atomicAdd(reinterpret_cast<float4*>(d.C + off[i]), o[i]);
          } else {
            if (d.epilogue == MMT_EPI_GELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) if (ok[i]) *reinterpret_cast<float4*>(d.aux + off[i]) = o[i];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = make_float4(gelu_fast(o[i].x), gelu_fast(o[i].y), gelu_fast(o[i].z), gelu_fast(o[i].w));
            } else if (d.epilogue == MMT_EPI_DGELU) {
              float4 u[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) u[i] = ok[i] ? *reinterpret_cast<const float4*>(d.aux + off[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                o[i].x *= dgelu_fast(u[i].x); o[i].y *= dgelu_fast(u[i].y);
                o[i].z *= dgelu_fast(u[i].z); o[i].w *= dgelu_fast(u[i].w);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) if (ok[i]) *reinterpret_cast<float4*>(d.C + off[i]) = o[i];
            if (d.colsum != nullptr) {                        // fused bias gradient: column sums of the output
              float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (ok[i]) { cs.x += o[i].x; cs.y += o[i].y; cs.z += o[i].z; cs.w += o[i].w; }
#pragma unroll
              for (int sh = 8; sh <= 16; sh <<= 1) {          // lanes sharing sub_c differ in bits 3,4
                cs.x += __shfl_xor_sync(0xffffffffu, cs.x, sh); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, sh);
                cs.z += __shfl_xor_sync(0xffffffffu, cs.z, sh); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, sh);
              }
              if (sub_r == 0 && col < d.N)
                atomicAdd(reinterpret_cast<float4*>(d.colsum + (int64_t)(z % d.batch_inner) * d.colsum_bs + col), cs);
            }
          }
        } else {
          // ragged right edge / unaligned C: predicated scalars
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!ok[i]) continue;
            float ov[4] = {o[i].x, o[i].y, o[i].z, o[i].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (col + t >= d.N) continue;
              float val = ov[t];
              if (d.add && (lead || args.split_k == 1)) val += d.add[off[i] + t];
              if (args.split_k > 1) { atomicAdd(d.C + off[i] + t, val); continue; }
              if (d.epilogue == MMT_EPI_GELU) { d.aux[off[i] + t] = val; val = gelu_fast(val); }
              else if (d.epilogue == MMT_EPI_DGELU) val *= dgelu_fast(d.aux[off[i] + t]);
              d.C[off[i] + t] = val;
              if (d.colsum != nullptr) atomicAdd(d.colsum + (int64_t)(z % d.batch_inner) * d.colsum_bs + col + t, val);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);         // this warp no longer reads the accumulator
    }
    if (args.c_tma && lane == 0) bulk_wait0();             // outstanding C stores complete before the CTA retires
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, bool A_MN, bool B_MN>
int launch2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const Tc2Args& args,
            cudaStream_t stream) {
  constexpr int STAGES = BN == 256 ? 4 : 6;
  constexpr size_t smem = STAGES * (A_BYTES + BN * BK * 4) + 4 * STG_BYTES_PER_WARP + 1024 + 256;
  static bool configured = false;
  auto kern = gemm_tc2_kernel<BN, A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "gemm_tc2 smem attribute");
    configured = true;
  }
  const int work = args.num_m_tiles * args.num_n_tiles * args.split_k * args.d.batch;
  const int grid = work < num_sms() ? work : num_sms();
  launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), smem, stream, ma, mb, mc, args);
  MMT_LAUNCH_CHECK("gemm_tc2_kernel");
  return 0;
}

}  // namespace

int make_tf32_map(CUtensorMap* map, const float* base, int rows, int K, int64_t rs, int64_t ks, bool mn_major,
                  int tile_rows, int batch_outer, int batch_inner, int64_t bs0, int64_t bs1, const char* what);

namespace {
template <int BN>
int dispatch2(const mmt_gemm_desc& d, Tc2Args& args, bool a_mn, bool b_mn, cudaStream_t stream) {
  CUtensorMap ma, mb;
  const int bo = d.batch / d.batch_inner;
  int rc = make_tf32_map(&ma, d.A, d.M, d.K, d.a_ms, d.a_ks, a_mn, BM, bo, d.batch_inner, d.a_bs0, d.a_bs1, "A");
  if (rc) return rc;
  rc = make_tf32_map(&mb, d.B, d.N, d.K, d.b_ns, d.b_ks, b_mn, BN, bo, d.batch_inner, d.b_bs0, d.b_bs1, "B");
  if (rc) return rc;
  args.num_n_tiles = (d.N + BN - 1) / BN;
  // plain epilogues leave through TMA stores (32 x 32 tiles, 128-byte swizzle, clipped at M / N)
  CUtensorMap mc = ma;
  static const bool tma_epi = [] { const char* e = getenv("MMT_TMA_EPILOGUE"); return !(e && e[0] == '0'); }();
  args.c_tma = tma_epi && d.epilogue == MMT_EPI_NONE && d.add == nullptr && args.split_k == 1 && d.c_mb == 0 &&
               (d.c_ms & 3) == 0 && ((d.c_bs0 | d.c_bs1) & 3) == 0 && ((uintptr_t)d.C & 15) == 0;
  if (args.c_tma) {
    rc = make_tf32_map(&mc, d.C, d.M, d.N, d.c_ms, 1, false, 32, bo, d.batch_inner, d.c_bs0, d.c_bs1, "C");
    if (rc) return rc;
  }
  if (!a_mn && !b_mn) return launch2<BN, false, false>(ma, mb, mc, args, stream);
  if (!a_mn && b_mn) return launch2<BN, false, true>(ma, mb, mc, args, stream);
  if (a_mn && !b_mn) return launch2<BN, true, false>(ma, mb, mc, args, stream);
  return launch2<BN, true, true>(ma, mb, mc, args, stream);
}
}  // namespace

// Sets *taken when this kernel handled the problem (else the caller uses the tiled kernel).
int gemm_tc_persistent(const mmt_gemm_desc& d, cudaStream_t stream, bool* taken) {
  *taken = false;
  if (d.c_mb != 0 || d.a_kb != 0 || d.batch % d.batch_inner != 0) return 0;
  const int64_t work128 = (int64_t)((d.M + BM - 1) / BM) * ((d.N + 127) / 128) * d.batch;
  const bool splittable = (d.flags & MMT_GEMM_SPLIT_K) && d.batch == 1 && d.K >= 1024;
  if (d.M < 128 || (work128 < 96 && !splittable)) return 0;  // tiny problems: tiled kernel
  const bool a_mn = (d.a_ks != 1), b_mn = (d.b_ks != 1);
  const bool wide = d.N > 160;                               // 128 x 256 tiles unless N is narrow
  Tc2Args args;
  args.d = d;
  args.d.alpha = d.alpha * kTf32TruncComp;
  args.num_m_tiles = (d.M + BM - 1) / BM;
  args.num_n_tiles = (d.N + (wide ? 255 : 127)) / (wide ? 256 : 128);
  args.num_kb = (d.K + BK - 1) / BK;
  args.split_k = 1;
  args.kb_per_split = args.num_kb;
  const int tiles = args.num_m_tiles * args.num_n_tiles;
  if ((d.flags & MMT_GEMM_SPLIT_K) && d.batch == 1 && d.c_ms == d.N && d.epilogue == MMT_EPI_NONE &&
      d.add != d.C && tiles * 2 <= num_sms() && args.num_kb >= 32) {
    int split = num_sms() / tiles;
    if (split > args.num_kb / 8) split = args.num_kb / 8;
    if (split > 1) {
      args.kb_per_split = (args.num_kb + split - 1) / split;
      args.split_k = (args.num_kb + args.kb_per_split - 1) / args.kb_per_split;
      cudaError_t e = cudaMemsetAsync(d.C, 0, sizeof(float) * (size_t)d.M * d.N, stream);
      if (e != cudaSuccess) return cuda_status(e, "gemm_tc2 split-K memset");
    }
  }
  *taken = true;
  return wide ? dispatch2<256>(d, args, a_mn, b_mn, stream) : dispatch2<128>(d, args, a_mn, b_mn, stream);
}

}  // namespace mmt
