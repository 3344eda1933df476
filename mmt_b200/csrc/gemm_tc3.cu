// Persistent tcgen05 TF32 GEMM on CTA PAIRS (cta_group::2): two CTAs of a cluster (one per SM of a TPC)
// own one 256 x 256 output tile.  Each CTA loads its 128 rows of A and only HALF of the B tile
// (128 of the 256 n-rows); the leader CTA issues tcgen05.mma.cta_group::2 (M=256, N=256, K=8), the
// tensor cores of both SMs read both halves.  Per SM and k-step this ingests 256 operand rows for a
// 128 x 256 output slice instead of the 384 rows of the 1-CTA 128x256 kernel (gemm_tc2.cu); operand
// ingest is what bounds the fp32-operand main loop on B200.
//
//   warp 0    TMA producer     5-stage ring of {A 128x32, B 128x32} fp32 tiles per CTA (32 KB / stage);
//                              cp.async.bulk.tensor...cta_group::2 credits both CTAs' bytes to the
//                              leader's mbarrier
//   warp 1    MMA issuer       (leader CTA) tcgen05.mma.cta_group::2 kind::tf32 M=256 N=256 K=8; the 512
//                              TMEM columns hold TWO accumulators so tile i+1's main loop overlaps tile
//                              i's epilogue; tcgen05.commit...multicast::cluster frees a stage in both CTAs
//   warps 2-9 epilogue         two warps per TMEM lane quarter (half of the columns each):
//                              tcgen05.ld -> per-warp shared-memory transpose -> fused
//                              bias / residual / erf-GELU / GELU' / column sums and fully coalesced
//                              128-bit global loads/stores (each store instruction covers 4 complete
//                              128-byte row segments)
// Tiles are rasterised in groups of 2 m-tiles x all n-tiles, so the 74 pairs that run together read
// few distinct A (activation) row blocks as well as few B (weight) tiles: the fp32-operand main loop
// is bound by L2 -> SM traffic (~9 TB/s aggregate measured), and concurrent readers of the same lines
// are cheaper than distinct streams (measured 3-8 % faster than m-fastest on the encoder's shapes).
// Weight-gradient shapes (few tiles, K = B*S) are split along K into (tile, k-range) work items
// whose epilogue reduces with red.global.add.v4.f32 into a zeroed C.
#include <cstdlib>

#include "pair_ptx.cuh"

namespace mmt {
namespace {
using namespace tc;

constexpr int BM = 128, BK = 32, UMMA_K = 8;
constexpr int EPI_WARPS = 8;                           // two warps per TMEM lane quarter, half the columns each
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t A_BYTES = BM * BK * 4;
constexpr int STG_PITCH = 36;                         // floats; 16 B aligned rows, conflict-free v4 phases
constexpr uint32_t STG_BYTES_PER_WARP = 32 * STG_PITCH * 4;

struct Tc3Args {
  mmt_gemm_desc d;
  int num_m_tiles, num_n_tiles;
  int split_k, kb_per_split, num_kb;
  int group_m;                                          // rasterisation: m-tiles per group (n advances inside a group)
};

// Pair tile 256 (m) x 256 (n); per CTA and stage: A 128x32 + B 128x32 fp32 = 32 KB, 6 stages.
// BF16 (experimental 16-bit operand mode): the same 128-byte rows hold 64 bf16 instead of 32 fp32, one
// k-block is 64 k, one MMA (kind::f16) covers K = 16; staging, barriers, TMEM use and the fp32 epilogue
// are unchanged.
template <bool A_MN, bool B_MN, bool BF16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const Tc3Args args) {
  constexpr int BN = 256, BNH = 128, STAGES = 5;
  constexpr uint32_t B_BYTES = BNH * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* staging = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP);
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
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * EPI_WARPS); }  // every epilogue warp of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // peer barriers initialised, TMEM allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                           // prologue done: now wait for the producer grid

  // work item -> (batch z, m0, n0, k-block range)
  auto decode = [&](int w, int& z, int& m0, int& n0, int& kb0, int& nkb) {
    z = w / (num_tiles * args.split_k);
    w -= z * num_tiles * args.split_k;
    const int tile = w / args.split_k, ks = w % args.split_k;
    // grouped rasterisation: groups of group_m m-tiles; inside a group m runs fastest, then n
    const int per_group = args.group_m * args.num_n_tiles;
    const int grp = tile / per_group, in_grp = tile - grp * per_group;
    const int gm = min(args.group_m, args.num_m_tiles - grp * args.group_m);   // last group may be short
    n0 = (in_grp / gm) * BN;
    m0 = (grp * args.group_m + in_grp % gm) * (2 * BM);   // pair tile: 256 rows
    kb0 = ks * args.kb_per_split;
    nkb = min(args.num_kb, kb0 + args.kb_per_split) - kb0;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t g = 0;                                  // global k-block counter (ring position)
      for (int w = pair_id; w < num_work; w += num_pairs) {
        int z, m0, n0, kb0, nkb;
        decode(w, z, m0, n0, kb0, nkb);
        const int z0 = z / d.batch_inner, z1 = z % d.batch_inner;
        for (int i = 0; i < nkb; ++i, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);                 // this CTA's slot is free (both CTAs get the commit)
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = (kb0 + i) * (BF16 ? 2 * BK : BK);      // elements: 128-byte rows either way
          const int ma = m0 + (int)rank * BM, nb_ = n0 + (int)rank * BNH;
          // MN-major operands arrive as boxes of one 128-byte row of MNW elements x KR k-rows
          constexpr int MNW = BF16 ? 64 : 32, KR = BF16 ? 64 : 32;
          if (!A_MN) {
            tma_load_4d_2sm(sa, &map_a, &full_bar[s], k0, ma, z1, z0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / MNW; ++j) tma_load_4d_2sm(sa + j * (KR * 128), &map_a, &full_bar[s], ma + MNW * j, k0, z1, z0);
          }
          if (!B_MN) {
            tma_load_4d_2sm(sb, &map_b, &full_bar[s], k0, nb_, z1, z0);
          } else {
#pragma unroll
            for (int j = 0; j < BNH / MNW; ++j) tma_load_4d_2sm(sb + j * (KR * 128), &map_b, &full_bar[s], nb_ + MNW * j, k0, z1, z0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      const uint32_t idesc = BF16 ? make_idesc_bf16(2 * BM, BN, A_MN, B_MN)
                                  : make_idesc_tf32(2 * BM, BN, A_MN, B_MN);     // M = 256 across the pair
      // MN-major operands.  tf32: 32-byte-atom swizzle (layout 1): 4-k-row groups 512 B apart (SBO), MN chunks
      // of 32 one box (32 k-rows x 128 B) apart (LBO), 8 k per MMA = 1024 B.  bf16: plain 128-byte swizzle
      // (layout 2): 8-k-row atoms 1024 B apart (SBO), MN chunks of 64 one box (64 k-rows x 128 B) apart (LBO),
      // 16 k per MMA = 2048 B.
      constexpr uint32_t MN_LBO = BF16 ? 64 * 128 : BK * 128, MN_SBO = BF16 ? 1024 : 512, MN_STEP = BF16 ? 2048 : 1024;
      constexpr uint32_t MN_LT = BF16 ? 2 : 1;
      constexpr uint32_t A_LBO = A_MN ? MN_LBO : 16, A_SBO = A_MN ? MN_SBO : 1024, A_STEP = A_MN ? MN_STEP : UMMA_K * 4;
      constexpr uint32_t B_LBO = B_MN ? MN_LBO : 16, B_SBO = B_MN ? MN_SBO : 1024, B_STEP = B_MN ? MN_STEP : UMMA_K * 4;
      constexpr uint32_t A_LT = A_MN ? MN_LT : 2, B_LT = B_MN ? MN_LT : 2;
      uint32_t g = 0;
      int it = 0;
      for (int w = pair_id; w < num_work; w += num_pairs, ++it) {
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
            if (BF16) umma_bf16_2sm(acc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
            else umma_tf32_2sm(acc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[s]);                    // frees the slot in both CTAs
        }
        umma_commit_2sm(&tfull_bar[buf]);                    // accumulator ready in both CTAs
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                               // TMEM lane quarter == output rows 32q..32q+31
    const int chalf = (warp - 2) >> 2;                    // which half of the 256 columns this warp drains
    const uint32_t stg = smem_u32(staging) + (uint32_t)(warp - 2) * STG_BYTES_PER_WARP;
    const bool vec_ok = ((d.c_ms & 3) == 0) && (((d.c_bs0 | d.c_bs1 | d.bias_bs) & 3) == 0) &&
                        ((((uintptr_t)d.C | (uintptr_t)d.bias | (uintptr_t)d.add | (uintptr_t)d.aux) & 15) == 0);
    const int sub_r = lane >> 3;                          // store phase: row within a group of 4
    const int sub_c = (lane & 7) * 4;                     // store phase: first of this lane's 4 columns
    int it = 0;
    for (int w = pair_id; w < num_work; w += num_pairs, ++it) {
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
      for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
        const int nb = n0 + c * 32;
        if (nb >= d.N) break;                             // warp-uniform
        float v[32];
        tmem_ld32(acc + (uint32_t)(c * 32), v);
        // phase 1: this lane's row -> warp-private staging (row pitch 36 floats)
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg + (uint32_t)(lane * STG_PITCH + j) * 4,
                 make_float4(v[j] * d.alpha, v[j + 1] * d.alpha, v[j + 2] * d.alpha, v[j + 3] * d.alpha));
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
          const int m = m0 + (int)rank * BM + q * 32 + rl;
          ok[i] = (m < d.M) && (col < d.N);
          off[i] = zoff + (int64_t)m * d.c_ms + col;
          o[i] = lds128(stg + (uint32_t)(rl * STG_PITCH + sub_c) * 4);
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
      if (lane == 0) {                                       // this warp no longer reads the accumulator
        if (leader) mbar_arrive_relaxed(&tempty_bar[buf]); else mbar_arrive_on_leader(&tempty_bar[buf]);
      }
    }
  }
  __syncthreads();
  cluster_sync_all();                                     // peer may still be reading this CTA's smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

template <bool A_MN, bool B_MN, bool BF16 = false>
int launch3(const CUtensorMap& ma, const CUtensorMap& mb, const Tc3Args& args, cudaStream_t stream) {
  constexpr size_t smem = 5 * (A_BYTES + 128 * BK * 4) + EPI_WARPS * STG_BYTES_PER_WARP + 1024 + 256;
  static bool configured = false;
  auto kern = gemm_tc3_kernel<A_MN, B_MN, BF16>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "gemm_tc3 smem attribute");
    configured = true;
  }
  const int work = args.num_m_tiles * args.num_n_tiles * args.split_k;
  const int max_pairs = num_sms() / 2;
  const int pairs = work < max_pairs ? work : max_pairs;
  launch_pdl(kern, dim3(2 * pairs), dim3(NUM_THREADS), smem, stream, ma, mb, args);          // __cluster_dims__(2,1,1)
  MMT_LAUNCH_CHECK("gemm_tc3_kernel");
  return 0;
}

}  // namespace

int make_tf32_map(CUtensorMap* map, const float* base, int rows, int K, int64_t rs, int64_t ks, bool mn_major,
                  int tile_rows, int batch_outer, int batch_inner, int64_t bs0, int64_t bs1, const char* what);

// CTA-pair kernel for large un-batched problems.  Sets *taken when it handled the problem.
int gemm_tc_pair(const mmt_gemm_desc& d, cudaStream_t stream, bool* taken) {
  *taken = false;
  if (d.batch != 1 || d.c_mb != 0 || d.a_kb != 0 || d.M < 512 || d.N < 192) return 0;
  const bool a_mn = (d.a_ks != 1), b_mn = (d.b_ks != 1);
  Tc3Args args;
  args.d = d;
  args.d.alpha = d.alpha * kTf32TruncComp;
  args.num_m_tiles = (d.M + 2 * BM - 1) / (2 * BM);
  args.num_n_tiles = (d.N + 255) / 256;
  args.num_kb = (d.K + BK - 1) / BK;
  args.split_k = 1;
  args.kb_per_split = args.num_kb;
  static const int group_m_env = [] { const char* e = getenv("MMT_PAIR_GROUP_M"); return e ? atoi(e) : 0; }();  // tuning switch
  args.group_m = group_m_env > 0 ? group_m_env : 2;
  if (args.group_m > args.num_m_tiles) args.group_m = args.num_m_tiles;
  const int tiles = args.num_m_tiles * args.num_n_tiles;
  const int max_pairs = num_sms() / 2;
  const bool can_split = (d.flags & MMT_GEMM_SPLIT_K) && d.c_ms == d.N && d.epilogue == MMT_EPI_NONE &&
                         d.add != d.C && d.colsum == nullptr;
  if (tiles * 2 <= max_pairs) {
    if (!can_split || args.num_kb < 32) return 0;            // too few tiles for 74 pairs: 1-CTA kernels
    int split = max_pairs / tiles;
    if (split > args.num_kb / 8) split = args.num_kb / 8;
    if (split > 1) {
      args.kb_per_split = (args.num_kb + split - 1) / split;
      args.split_k = (args.num_kb + args.kb_per_split - 1) / args.kb_per_split;
      cudaError_t e = cudaMemsetAsync(d.C, 0, sizeof(float) * (size_t)d.M * d.N, stream);
      if (e != cudaSuccess) return cuda_status(e, "gemm_tc3 split-K memset");
    }
  }
  CUtensorMap ma, mb;
  int rc = make_tf32_map(&ma, d.A, d.M, d.K, d.a_ms, d.a_ks, a_mn, BM, 1, 1, 0, 0, "A");
  if (rc) return rc;
  rc = make_tf32_map(&mb, d.B, d.N, d.K, d.b_ns, d.b_ks, b_mn, 128, 1, 1, 0, 0, "B");
  if (rc) return rc;
  *taken = true;
  if (!a_mn && !b_mn) return launch3<false, false>(ma, mb, args, stream);
  if (!a_mn && b_mn) return launch3<false, true>(ma, mb, args, stream);
  if (a_mn && !b_mn) return launch3<true, false>(ma, mb, args, stream);
  return launch3<true, true>(ma, mb, args, stream);
}

// Experimental: C = epilogue(alpha * A B^T ...) with bf16 operands (d.A / d.B point to bf16 data, strides in
// elements, each contiguous along k or along m / n), fp32 C and epilogue operands.  CTA-pair kernel only.
int make_bf16_map(CUtensorMap* map, const void* base, int rows, int K, int64_t ld, bool mn_major, int tile_rows,
                  const char* what);

int gemm_tc_pair_bf16(const mmt_gemm_desc& d, cudaStream_t stream) {
  MMT_ARG_CHECK(d.batch == 1 && d.c_mb == 0 && d.a_kb == 0, MMT_E_UNSUPPORTED, "mmt_gemm(bf16): un-batched operands only");
  const bool a_mn = (d.a_ks != 1), b_mn = (d.b_ks != 1);
  MMT_ARG_CHECK((!a_mn || d.a_ms == 1) && (!b_mn || d.b_ns == 1), MMT_E_UNSUPPORTED,
                "mmt_gemm(bf16): operands must be contiguous along k or along m / n");
  MMT_ARG_CHECK(d.M >= 256 && d.N >= 128 && d.K >= 64, MMT_E_UNSUPPORTED, "mmt_gemm(bf16): problem too small (M=%d N=%d K=%d)", d.M, d.N, d.K);
  MMT_ARG_CHECK(!(d.flags & MMT_GEMM_SPLIT_K), MMT_E_UNSUPPORTED, "mmt_gemm(bf16): split-K is not wired up");
  Tc3Args args;
  args.d = d;                                             // bf16 x bf16 products are exact in fp32: no compensation
  args.num_m_tiles = (d.M + 2 * BM - 1) / (2 * BM);
  args.num_n_tiles = (d.N + 255) / 256;
  args.num_kb = (d.K + 63) / 64;
  args.split_k = 1;
  args.kb_per_split = args.num_kb;
  args.group_m = 2 < args.num_m_tiles ? 2 : args.num_m_tiles;
  CUtensorMap ma, mb;
  int rc = make_bf16_map(&ma, d.A, d.M, d.K, a_mn ? d.a_ks : d.a_ms, a_mn, BM, "A");
  if (rc) return rc;
  rc = make_bf16_map(&mb, d.B, d.N, d.K, b_mn ? d.b_ks : d.b_ns, b_mn, 128, "B");
  if (rc) return rc;
  if (!a_mn && !b_mn) return launch3<false, false, true>(ma, mb, args, stream);
  if (!a_mn && b_mn) return launch3<false, true, true>(ma, mb, args, stream);
  if (a_mn && !b_mn) return launch3<true, false, true>(ma, mb, args, stream);
  return launch3<true, true, true>(ma, mb, args, stream);
}

}  // namespace mmt
