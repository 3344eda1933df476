// 16-bit-operand tcgen05 GEMM of the train step (mmt_gemm16): fp16 or bf16 A / B tiles, fp32 accumulation
// in TMEM, fp32 and / or 16-bit outputs.  One persistent CTA-PAIR kernel (cta_group::2, 256 x 256 pair
// tile, see gemm_tc3.cu for the tf32 ancestor) serves every dense product of the step:
//
//   warp 0    TMA producer     5-stage ring of {A 128 x 64, B 128 x 64} 16-bit tiles per CTA (32 KB / stage):
//                              K-major operands as one 128-byte-swizzled box, MN-major ones (dgrad: W read
//                              "transposed", wgrad: dY^T and X^T) as 64 x 64 boxes -- no transposes in HBM
//   warp 1    MMA issuer       (leader CTA) tcgen05.mma.cta_group::2.kind::f16, M=256 N=256 K=16; two TMEM
//                              accumulators, so tile i+1's main loop overlaps tile i's epilogue
//   warps 2-9 epilogue         tcgen05.ld -> per-warp smem transpose -> fused epilogue on coalesced accesses:
//                                v = alpha * acc + bias
//                                GELU : aux16 <- v (pre-activation for the backward), v = gelu_erf(v)
//                                DGELU: v *= gelu_erf'(aux16)
//                                dropout (Philox mask of (seed, site, row, col/4) -- the mask mmt_ln16_bwd
//                                regenerates), then + add (fp32 residual)
//                                C32 <- v, C16 <- rn16(v * out16_scale), colsum[n] += colsum_scale * v
// Work items are (batch z, tile) or (tile, k-range) for split-K weight gradients (fp32 red.add into a
// zeroed C32).  Rasterisation in groups of 2 m-tiles x all n-tiles (co-running pairs share A row blocks).
//
// Operand precision: products of two 11-bit significands are exact in fp32, so the only rounding is the
// producers' round-to-nearest conversion of the operands (no truncation, no compensation factor).
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "cvt16.cuh"
#include "pair_ptx.cuh"

namespace mmt {
namespace {
using namespace tc;

constexpr int BM = 128, BK = 64, UMMA_K = 16, STAGES = 4;
constexpr int BN_MAX = 256;                           // pair tile 256 x BN, BN = 256 or 128 (template parameter)
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t A_BYTES = BM * BK * 2, STAGE_BYTES = A_BYTES + (BN_MAX / 2) * BK * 2;   // ring slots sized for BN = 256
constexpr int STG_PITCH = 36;                         // 32-column chunks (fp32 outputs)
constexpr int STG_PITCH16 = 68;                       // 64-column chunks (16-bit-only outputs)
constexpr uint32_t STG_BYTES_PER_WARP = 32 * STG_PITCH16 * 4;
constexpr size_t SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP + 1024 + 256 + EPI_WARPS * 512;

struct G16Args {
  mmt_gemm16_desc d;
  int num_m_tiles, num_n_tiles;
  int split_k, kb_per_split, num_kb;
  int group_m;
};

__device__ __forceinline__ uint2 ldg_u2(const void* p) { return *reinterpret_cast<const uint2*>(p); }

// OUT16: the GEMM's only output is 16-bit (C16, optionally the GELU pre-activation aux16).  These are the
// output-heavy, short-K products (QKV, FFN-up, GELU' dgrad: up to 2 x 86 MB written per launch); their epilogue
// computes in the TMEM layout and moves data with TMA only (see the epilogue).  With per-lane global stores they
// ran epilogue-bound at 1/3 - 1/2 of the tensor-core rate.
// BN: 256, or 128 when the 256-wide tiling would leave the last wave of CTA pairs half empty (N = 512 outputs of
// 13952 rows: 110 tiles on 74 pairs = 2 rounds at 74 % -> 220 tiles = 3 rounds at 99 %).
template <bool OUT16, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_x, const G16Args args) {
  constexpr int BNH = BN / 2;
  constexpr uint32_t TX_BYTES = A_BYTES + BNH * BK * 2;     // bytes one CTA receives per k-block
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
  uint64_t* aux_bar = tempty_bar + 2;             // [EPI_WARPS] GELU' operand tiles (OUT16 epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES_PER_WARP + 256);   // [EPI_WARPS][128]

  const mmt_gemm16_desc& d = args.d;
  const bool bf16 = d.dtype == MMT_DT_BF16;
  const bool a_mn = d.a_mn != 0, b_mn = d.b_mn != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = args.num_m_tiles * args.num_n_tiles;
  const int num_work = num_tiles * args.split_k * d.batch;      // split-K and batching are exclusive

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * EPI_WARPS); }
    for (int b = 0; b < EPI_WARPS; ++b) mbar_init(&aux_bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  auto decode = [&](int w, int& z, int& m0, int& n0, int& kb0, int& nkb) {
    z = w / (num_tiles * args.split_k);
    w -= z * num_tiles * args.split_k;
    const int tile = w / args.split_k, ks = w % args.split_k;
    const int per_group = args.group_m * args.num_n_tiles;
    const int grp = tile / per_group, in_grp = tile - grp * per_group;
    const int gm = min(args.group_m, args.num_m_tiles - grp * args.group_m);
    n0 = (in_grp / gm) * BN;
    m0 = (grp * args.group_m + in_grp % gm) * (2 * BM);
    kb0 = ks * args.kb_per_split;
    nkb = min(args.num_kb, kb0 + args.kb_per_split) - kb0;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t g = 0;
      for (int w = pair_id; w < num_work; w += num_pairs) {
        int z, m0, n0, kb0, nkb;
        decode(w, z, m0, n0, kb0, nkb);
        const int z0 = z / d.batch_inner, z1 = z % d.batch_inner;
        const int ma = m0 + (int)rank * BM, nb_ = n0 + (int)rank * BNH;
        for (int i = 0; i < nkb; ++i, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * TX_BYTES);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = (kb0 + i) * BK;
          if (!a_mn) {
            tma_load_4d_2sm(sa, &map_a, &full_bar[s], k0, ma, z1, z0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_4d_2sm(sa + j * (BK * 128), &map_a, &full_bar[s], ma + 64 * j, k0, z1, z0);
          }
          if (!b_mn) {
            tma_load_4d_2sm(sb, &map_b, &full_bar[s], k0, nb_, z1, z0);
          } else {
#pragma unroll
            for (int j = 0; j < BNH / 64; ++j) tma_load_4d_2sm(sb + j * (BK * 128), &map_b, &full_bar[s], nb_ + 64 * j, k0, z1, z0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      const uint32_t idesc = make_idesc_16(2 * BM, BN, a_mn, b_mn, bf16);
      // K-major: 128-byte rows of 64 k, 8-row groups 1024 B apart (SBO), 16 k per MMA = +32 B.
      // MN-major: rows of 64 m/n per k, 8-k-row atoms 1024 B apart (SBO), 64-wide m/n chunks one box
      // (64 k-rows x 128 B) apart (LBO), 16 k per MMA = +2048 B.  128-byte swizzle either way.
      const uint32_t a_lbo = a_mn ? BK * 128 : 16, a_step = a_mn ? 2048 : UMMA_K * 2;
      const uint32_t b_lbo = b_mn ? BK * 128 : 16, b_step = b_mn ? 2048 : UMMA_K * 2;
      uint32_t g = 0;
      int it = 0;
      for (int w = pair_id; w < num_work; w += num_pairs, ++it) {
        int z, m0, n0, kb0, nkb;
        decode(w, z, m0, n0, kb0, nkb);
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
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
            const uint64_t da = make_smem_desc(sa + k * a_step, a_lbo, 1024, 2);
            const uint64_t db = make_smem_desc(sb + k * b_step, b_lbo, 1024, 2);
            umma_bf16_2sm(acc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);     // kind::f16 (format bits in idesc)
          }
          umma_commit_2sm(&empty_bar[s]);
        }
        umma_commit_2sm(&tfull_bar[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                               // TMEM lane quarter == output rows 32q..32q+31
    const int chalf = (warp - 2) >> 2;                    // which half of the 256 columns this warp drains
    const uint32_t stg = smem_u32(staging) + (uint32_t)(warp - 2) * STG_BYTES_PER_WARP;
    const int sub_r = lane >> 3;                          // access phase: row within a group of 4
    uint16_t* C16 = reinterpret_cast<uint16_t*>(d.C16);
    uint16_t* X16 = reinterpret_cast<uint16_t*>(d.aux16);
    int it = 0;
    if constexpr (OUT16) {
      // ---------- 16-bit-only outputs: all math in the TMEM layout (thread = output row, 64 columns per chunk),
      // results leave through TMA stores from a 128-byte-swizzled [32 rows x 64 columns] tile per warp; the GELU'
      // operand arrives the same way (TMA load).  No fp32 staging, no per-lane global accesses. ----------
      uint8_t* tile_c = reinterpret_cast<uint8_t*>(staging) + (uint32_t)(warp - 2) * 8192;   // 2 x 4 KB tiles, 1024-aligned (swizzle phase)
      uint8_t* tile_x = tile_c + 4096;
      float* sbias = bias_s + (warp - 2) * 128;               // this warp's 128 bias values of the current tile
      uint64_t* xbar = &aux_bar[warp - 2];
      const uint32_t tc_u = smem_u32(tile_c), tx_u = smem_u32(tile_x);
      const uint32_t sw = (uint32_t)(lane & 7);
      uint32_t xphase = 0;
      for (int w = pair_id; w < num_work; w += num_pairs, ++it) {
        int z, m0, n0, kb0, nkb;
        decode(w, z, m0, n0, kb0, nkb);
        const int buf = it & 1;
        const int z0 = z / d.batch_inner, z1 = z % d.batch_inner;
        const float* bias = d.bias ? d.bias + (int64_t)z * d.bias_bs : nullptr;
        const int row0 = m0 + (int)rank * BM + q * 32;
        const bool row_ok = row0 + lane < d.M;
        if (bias) {
          const int bc = n0 + chalf * BNH + lane * 4;
          const float4 b = (lane * 4 < BNH && bc < d.N) ? __ldg(reinterpret_cast<const float4*>(bias + bc)) : make_float4(0.f, 0.f, 0.f, 0.f);
          __syncwarp();
          *reinterpret_cast<float4*>(sbias + lane * 4) = b;
          __syncwarp();
        }
        mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c = chalf * (BN / 128); c < (chalf + 1) * (BN / 128); ++c) {
          const int col0 = n0 + c * 64;
          if (col0 >= d.N || (d.flags & 512)) break;         // warp-uniform (512: timing experiment, no epilogue)
          if (d.epilogue == MMT_EPI_DGELU && lane == 0) {    // GELU' operand tile: asynchronous, consumed below
            mbar_arrive_expect_tx(xbar, 4096);
            tma_load_4d(tile_x, &map_x, xbar, col0, row0, z1, z0);
          }
          float v[64];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float t[32];
            tmem_ld32(acc + (uint32_t)(c * 64 + hh * 32), t);
            if (d.alpha != 1.0f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[hh * 32 + j] = t[j] * d.alpha;
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[hh * 32 + j] = t[j];
            }
          }
          if (bias) {                                        // same 64 values for every lane: broadcast reads
            const float* sb = sbias + (c - chalf * (BN / 128)) * 64;
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(sb + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (d.flags & 256) continue;                       // timing experiment: TMEM drain only
          if (lane == 0) bulk_wait_read0();                  // the previous chunk's TMA stores have read both tiles
          __syncwarp();
          if (d.epilogue == MMT_EPI_GELU) {
            // activation and its derivative from ONE erf / exp evaluation; the derivative is what the backward
            // GEMM's epilogue multiplies by (aux16), so GELU' costs the backward two instructions per element
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float g[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) gelu_both(v[8 * j + t], v[8 * j + t], g[t]);
              sts128u(tx_u + (uint32_t)lane * 128 + (((uint32_t)j ^ sw) << 4), pack2(g[0], g[1], bf16), pack2(g[2], g[3], bf16),
                      pack2(g[4], g[5], bf16), pack2(g[6], g[7], bf16));
            }
          } else if (d.epilogue == MMT_EPI_DGELU) {
            mbar_wait(xbar, xphase);
            xphase ^= 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 u;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                           : "r"(tx_u + (uint32_t)lane * 128 + (((uint32_t)j ^ sw) << 4)) : "memory");
              const float2 u0 = unpack2(u.x, bf16), u1 = unpack2(u.y, bf16), u2 = unpack2(u.z, bf16), u3 = unpack2(u.w, bf16);
              v[8 * j] *= u0.x; v[8 * j + 1] *= u0.y; v[8 * j + 2] *= u1.x; v[8 * j + 3] *= u1.y;
              v[8 * j + 4] *= u2.x; v[8 * j + 5] *= u2.y; v[8 * j + 6] *= u3.x; v[8 * j + 7] *= u3.y;
            }
          }
          if (d.out16_scale != 1.0f) {
            const float s16 = d.out16_scale;
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] *= s16;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128u(tc_u + (uint32_t)lane * 128 + (((uint32_t)j ^ sw) << 4), pack2(v[8 * j], v[8 * j + 1], bf16),
                    pack2(v[8 * j + 2], v[8 * j + 3], bf16), pack2(v[8 * j + 4], v[8 * j + 5], bf16),
                    pack2(v[8 * j + 6], v[8 * j + 7], bf16));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {                                   // clipped at the tensor's edges by the TMA unit
            tma_store_4d(&map_c, tc_u, col0, row0, z1, z0);
            if (d.epilogue == MMT_EPI_GELU) tma_store_4d(&map_x, tx_u, col0, row0, z1, z0);
            bulk_commit();
          }
          if (d.colsum != nullptr) {
            // fused bias gradient: column sums over this warp's 32 rows by a halving butterfly (62 shuffles for 64
            // columns); afterwards lane L holds columns 2 * bitrev5(L) and 2 * bitrev5(L) + 1 ... see `cid`
            if (!row_ok) {
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float send = (lane & 16) ? v[j] : v[j + 32], keep = (lane & 16) ? v[j + 32] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float send = (lane & 8) ? v[j] : v[j + 16], keep = (lane & 8) ? v[j + 16] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float send = (lane & 4) ? v[j] : v[j + 8], keep = (lane & 4) ? v[j + 8] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float send = (lane & 2) ? v[j] : v[j + 4], keep = (lane & 2) ? v[j + 4] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float send = (lane & 1) ? v[j] : v[j + 2], keep = (lane & 1) ? v[j + 2] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
            // lane L now holds columns cid and cid + 1 with cid = 32*b4 + 16*b3 + 8*b2 + 4*b1 + 2*b0
            const int cid = ((lane >> 4) & 1) * 32 + ((lane >> 3) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 +
                            (lane & 1) * 2;
            if (col0 + cid < d.N) {
              float* p = d.colsum + (int64_t)z1 * d.colsum_bs + col0 + cid;
              const float k = d.colsum_scale / d.out16_scale;   // v carries out16_scale already
              atomicAdd(p, v[0] * k);
              atomicAdd(p + 1, v[1] * k);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive_relaxed(&tempty_bar[buf]); else mbar_arrive_on_leader(&tempty_bar[buf]);
        }
      }
      if (lane == 0) bulk_wait0();                           // every TMA store has completed before the CTA retires
    } else {
    // ---------- general epilogue: 32-column chunks, 4 columns per lane (fp32 lines are complete at 16 B / lane) ----------
    const bool vec32 = (d.C32 == nullptr || (((d.c32_ld | d.c_bs0 | d.c_bs1) & 3) == 0 && ((uintptr_t)d.C32 & 15) == 0)) &&
                       (d.add == nullptr || ((d.add_ld & 3) == 0 && ((uintptr_t)d.add & 15) == 0)) &&
                       (d.bias == nullptr || ((d.bias_bs & 3) == 0 && ((uintptr_t)d.bias & 15) == 0)) &&
                       (d.colsum == nullptr || ((d.colsum_bs & 3) == 0 && ((uintptr_t)d.colsum & 15) == 0));
    const bool vec16 = (d.C16 == nullptr || (((d.c16_ld | d.c_bs0 | d.c_bs1) & 3) == 0 && ((uintptr_t)d.C16 & 7) == 0)) &&
                       (d.aux16 == nullptr || (((d.aux_ld | d.c_bs0 | d.c_bs1) & 3) == 0 && ((uintptr_t)d.aux16 & 7) == 0));
    const bool vec_ok = vec32 && vec16;
    const int sub_c = (lane & 7) * 4;                     // store phase: first of this lane's 4 columns
    const float inv_keep = d.p_drop > 0.f ? 1.f / (1.f - d.p_drop) : 1.f;
    const uint64_t seed = d.seed + (d.seed_ctr ? *d.seed_ctr : 0);
    const uint32_t key32 = drop_key(seed, d.site), thr16 = (uint32_t)(d.p_drop * 65536.0f);
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
        if (nb >= d.N || (d.flags & 512)) break;           // warp-uniform (512: timing experiment, no epilogue)
        const int col = nb + sub_c;
        const bool full = vec_ok && (col + 4 <= d.N);
        bool ok[8];
        int mrow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          mrow[i] = m0 + (int)rank * BM + q * 32 + 4 * i + sub_r;
          ok[i] = (mrow[i] < d.M) && (col < d.N);
        }
        // residual rows first: their global-load latency runs under the TMEM drain and the staging round trip
        float4 a[8];
        const bool pre_add = full && d.add != nullptr && args.split_k == 1;
        if (pre_add) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            a[i] = ok[i] ? *reinterpret_cast<const float4*>(d.add + zoff + (int64_t)mrow[i] * d.add_ld + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float v[32];
        tmem_ld32(acc + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg + (uint32_t)(lane * STG_PITCH + j) * 4,
                 make_float4(v[j] * d.alpha, v[j + 1] * d.alpha, v[j + 2] * d.alpha, v[j + 3] * d.alpha));
        __syncwarp();
        if (d.flags & 256) { __syncwarp(); continue; }     // timing experiment: TMEM drain + staging only
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias && lead) {
#pragma unroll
          for (int t = 0; t < 4; ++t) if (col + t < d.N) bv[t] = bias[col + t];
        }
        float4 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = lds128(stg + (uint32_t)((4 * i + sub_r) * STG_PITCH + sub_c) * 4);
          o[i].x += bv[0]; o[i].y += bv[1]; o[i].z += bv[2]; o[i].w += bv[3];
        }
        if (args.split_k > 1) {
          // weight-gradient work item: fp32 reduction into the zeroed C32
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!ok[i]) continue;
            float* p = d.C32 + zoff + (int64_t)mrow[i] * d.c32_ld + col;
            if (full) atomicAdd(reinterpret_cast<float4*>(p), o[i]);
            else {
              const float ov[4] = {o[i].x, o[i].y, o[i].z, o[i].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) if (col + t < d.N) atomicAdd(p + t, ov[t]);
            }
          }
        } else if (full) {
          if (d.epilogue == MMT_EPI_GELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 g;
              gelu_both(o[i].x, o[i].x, g.x); gelu_both(o[i].y, o[i].y, g.y);
              gelu_both(o[i].z, o[i].z, g.z); gelu_both(o[i].w, o[i].w, g.w);
              if (ok[i]) *reinterpret_cast<uint2*>(X16 + zoff + (int64_t)mrow[i] * d.aux_ld + col) = pack4(g, bf16);
            }
          } else if (d.epilogue == MMT_EPI_DGELU) {
            uint2 u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = ok[i] ? ldg_u2(X16 + zoff + (int64_t)mrow[i] * d.aux_ld + col) : make_uint2(0u, 0u);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 uu = unpack4(u[i], bf16);
              o[i].x *= uu.x; o[i].y *= uu.y; o[i].z *= uu.z; o[i].w *= uu.w;
            }
          }
          if (d.p_drop > 0.f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 sc = dropout_scale4_fast(key32, (uint32_t)mrow[i], (uint32_t)(col >> 2), thr16, inv_keep);
              o[i].x *= sc.x; o[i].y *= sc.y; o[i].z *= sc.z; o[i].w *= sc.w;
            }
          }
          if (pre_add) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { o[i].x += a[i].x; o[i].y += a[i].y; o[i].z += a[i].z; o[i].w += a[i].w; }
          }
          if (d.C32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (ok[i]) *reinterpret_cast<float4*>(d.C32 + zoff + (int64_t)mrow[i] * d.c32_ld + col) = o[i];
          }
          if (C16) {
            const float s16 = d.out16_scale;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (ok[i]) *reinterpret_cast<uint2*>(C16 + zoff + (int64_t)mrow[i] * d.c16_ld + col) =
                  pack4(make_float4(o[i].x * s16, o[i].y * s16, o[i].z * s16, o[i].w * s16), bf16);
          }
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
            const float k = d.colsum_scale;
            if (sub_r == 0)
              atomicAdd(reinterpret_cast<float4*>(d.colsum + (int64_t)(z % d.batch_inner) * d.colsum_bs + col),
                        make_float4(cs.x * k, cs.y * k, cs.z * k, cs.w * k));
          }
        } else {
          // ragged right edge / unaligned operands: predicated scalars
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!ok[i]) continue;
            const float ov[4] = {o[i].x, o[i].y, o[i].z, o[i].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (col + t >= d.N) continue;
              float val = ov[t];
              if (d.epilogue == MMT_EPI_GELU) {
                float g;
                gelu_both(val, val, g);
                X16[zoff + (int64_t)mrow[i] * d.aux_ld + col + t] = pack1(g, bf16);
              } else if (d.epilogue == MMT_EPI_DGELU) val *= unpack1(X16[zoff + (int64_t)mrow[i] * d.aux_ld + col + t], bf16);
              if (d.p_drop > 0.f) {
                const float4 sc = dropout_scale4_fast(key32, (uint32_t)mrow[i], (uint32_t)((col + t) >> 2), thr16, inv_keep);
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
                val *= scv[(col + t) & 3];
              }
              if (d.add) val += d.add[zoff + (int64_t)mrow[i] * d.add_ld + col + t];
              if (d.C32) d.C32[zoff + (int64_t)mrow[i] * d.c32_ld + col + t] = val;
              if (C16) C16[zoff + (int64_t)mrow[i] * d.c16_ld + col + t] = pack1(val * d.out16_scale, bf16);
              if (d.colsum != nullptr) atomicAdd(d.colsum + (int64_t)(z % d.batch_inner) * d.colsum_bs + col + t, val * d.colsum_scale);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive_relaxed(&tempty_bar[buf]); else mbar_arrive_on_leader(&tempty_bar[buf]);
      }
    }
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

// ---- tensor maps (cached: the encode call costs ~1.5 us and a train step needs ~200) -----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode16() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}

struct MapKey {
  uint64_t base, rows, K, ld, bs0, bs1;
  uint32_t mn, tile_rows, bo, bi, dtype, rank;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && K == o.K && ld == o.ld && bs0 == o.bs0 && bs1 == o.bs1 && mn == o.mn &&
           tile_rows == o.tile_rows && bo == o.bo && bi == o.bi && dtype == o.dtype && rank == o.rank;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    const uint64_t w[10] = {k.base, k.rows, k.K, k.ld, k.bs0, k.bs1, ((uint64_t)k.mn << 32) | k.tile_rows,
                            ((uint64_t)k.bo << 32) | k.bi, k.dtype, k.rank};
    for (uint64_t x : w) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

}  // namespace

// 16-bit operand of `rows` x K (element (r, k) at r*ld + k, or k*ld + r when mn_major), optional two batch
// dimensions; rank-4 map {inner, outer, batch_inner, batch_outer}, 128-byte swizzle.
// K-major: box {64 k, tile_rows};  MN-major: box {64 rows, 64 k}.
int make_map16(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, bool mn_major, int tile_rows,
               int batch_outer, int batch_inner, int64_t bs0, int64_t bs1, int dtype, const char* what, bool output = false) {
  EncodeTiledFn enc = get_encode16();
  MMT_ARG_CHECK(enc != nullptr, MMT_E_UNSUPPORTED, "mmt_gemm16: cuTensorMapEncodeTiled unavailable");
  MMT_ARG_CHECK(((uintptr_t)base % 16) == 0 && (ld * 2) % 16 == 0 && ld >= 1, MMT_E_ALIGN,
                "mmt_gemm16: operand %s needs a 16-byte aligned base and pitch (ld=%lld)", what, (long long)ld);
  MapKey key{(uint64_t)(uintptr_t)base, (uint64_t)rows, (uint64_t)K, (uint64_t)ld, (uint64_t)bs0, (uint64_t)bs1,
             mn_major ? 1u : 0u, (uint32_t)tile_rows, (uint32_t)batch_outer, (uint32_t)batch_inner, (uint32_t)dtype,
             output ? 5u : 4u};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *map = it->second; return 0; }
  }
  cuuint64_t dims[4] = {(cuuint64_t)(mn_major ? rows : K), (cuuint64_t)(mn_major ? K : rows), (cuuint64_t)batch_inner,
                        (cuuint64_t)batch_outer};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)((batch_inner > 1 ? bs1 : ld) * 2),
                           (cuuint64_t)((batch_outer > 1 ? bs0 : ld) * 2)};
  MMT_ARG_CHECK(strides[1] % 16 == 0 && strides[2] % 16 == 0, MMT_E_ALIGN,
                "mmt_gemm16: operand %s batch strides must be multiples of 8 elements", what);
  cuuint32_t box[4] = {64, (cuuint32_t)(mn_major ? 64 : tile_rows), 1, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, dtype == MMT_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, output ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMT_ARG_CHECK(r == CUDA_SUCCESS, MMT_E_UNSUPPORTED, "mmt_gemm16: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_maps.size() > 8192) g_maps.clear();
    g_maps.emplace(key, *map);
  }
  return 0;
}

}  // namespace mmt

using namespace mmt;

namespace mmt {
namespace {
// CTA pairs of the persistent grid: one per two SMs.  MMT_GEMM16_PAIRS caps it (experiment: leaving SMs to the gradient
// all-reduce that runs underneath the backward GEMMs did NOT help at N = 2 -- 3.55 ms uncapped, 3.61 / 3.57 / 3.66 ms
// with 70 / 66 / 60 pairs -- so nothing sets it).
int gemm16_pairs_limit() {
  static const int env_cap = [] { const char* e = getenv("MMT_GEMM16_PAIRS"); return e ? atoi(e) : 0; }();
  int lim = num_sms() / 2;
  if (env_cap > 0 && env_cap < lim) lim = env_cap;
  return lim < 1 ? 1 : lim;
}
}  // namespace
}  // namespace mmt

extern "C" int mmt_gemm16(const mmt_gemm16_desc* dp, void* stream_) {
  MMT_ARG_CHECK(dp != nullptr, MMT_E_ARG, "mmt_gemm16: null descriptor");
  const mmt_gemm16_desc& d = *dp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MMT_ARG_CHECK(d.A && d.B && (d.C32 || d.C16), MMT_E_ARG, "mmt_gemm16: null operand / no output");
  MMT_ARG_CHECK(d.M >= 0 && d.N >= 0 && d.K >= 1 && d.batch >= 1 && d.batch_inner >= 1 && d.batch % d.batch_inner == 0,
                MMT_E_SHAPE, "mmt_gemm16: bad shape M=%d N=%d K=%d batch=%d/%d", d.M, d.N, d.K, d.batch, d.batch_inner);
  MMT_ARG_CHECK(d.dtype == MMT_DT_F16 || d.dtype == MMT_DT_BF16, MMT_E_ARG, "mmt_gemm16: bad dtype %d", d.dtype);
  MMT_ARG_CHECK(d.epilogue >= MMT_EPI_NONE && d.epilogue <= MMT_EPI_DGELU, MMT_E_ARG, "mmt_gemm16: bad epilogue %d", d.epilogue);
  MMT_ARG_CHECK(d.epilogue == MMT_EPI_NONE || d.aux16 != nullptr, MMT_E_ARG, "mmt_gemm16: epilogue %d needs aux16", d.epilogue);
  MMT_ARG_CHECK(d.p_drop >= 0.f && d.p_drop < 1.f, MMT_E_ARG, "mmt_gemm16: p_drop=%f", (double)d.p_drop);
  MMT_ARG_CHECK(d.p_drop == 0.f || d.N <= 4096, MMT_E_UNSUPPORTED, "mmt_gemm16: epilogue dropout needs N <= 4096 (N=%d)", d.N);
  if (d.M == 0 || d.N == 0) return 0;
  G16Args args;
  args.d = d;
  args.num_m_tiles = (d.M + 2 * BM - 1) / (2 * BM);
  // tile width: 256 unless the narrower tiling fills the CTA-pair waves markedly better (see the kernel comment)
  int BN = 256;
  const int max_pairs = gemm16_pairs_limit();
  {
    const int pairs_max = max_pairs;
    const auto eff = [&](int bn) {
      const long t = (long)args.num_m_tiles * ((d.N + bn - 1) / bn) * d.batch;
      const long rounds = (t + pairs_max - 1) / pairs_max;
      return (double)t / (double)(rounds * pairs_max);
    };
    // ... and only for short K: at K = 3072 the narrower tile's extra A-operand traffic costs more than the wave gains
    // (measured: O-proj K=512 28.7 -> 23.8 us, FFN-down K=3072 53.5 -> 58.8 us)
    if (!(d.flags & MMT_GEMM_SPLIT_K) && d.N > 128 && d.K <= 1024 && eff(128) > 1.15 * eff(256)) BN = 128;
    static const int force = [] { const char* e = getenv("MMT_GEMM16_BN"); return e ? atoi(e) : 0; }();   // A/B switch
    if (force == 128 || force == 256) BN = force;
  }
  args.num_n_tiles = (d.N + BN - 1) / BN;
  args.num_kb = (d.K + BK - 1) / BK;
  args.split_k = 1;
  args.kb_per_split = args.num_kb;
  args.group_m = 2 < args.num_m_tiles ? 2 : args.num_m_tiles;
  const int tiles = args.num_m_tiles * args.num_n_tiles;
  if (d.flags & MMT_GEMM_SPLIT_K) {
    MMT_ARG_CHECK(d.batch == 1 && d.C32 && !d.C16 && d.epilogue == MMT_EPI_NONE && !d.add && !d.bias && !d.colsum &&
                  d.p_drop == 0.f, MMT_E_UNSUPPORTED, "mmt_gemm16: split-K needs a plain un-batched fp32 output");
    if (tiles * 2 <= max_pairs && args.num_kb >= 8) {
      int split = max_pairs / tiles;
      if (split > args.num_kb / 4) split = args.num_kb / 4;
      if (split > 1) {
        args.kb_per_split = (args.num_kb + split - 1) / split;
        args.split_k = (args.num_kb + args.kb_per_split - 1) / args.kb_per_split;
      }
    }
    if (args.split_k > 1) {
      // rows of C32 may be strided (c32_ld >= N): zero row by row through a 2-D memset
      cudaError_t e = cudaMemset2DAsync(d.C32, sizeof(float) * (size_t)d.c32_ld, 0, sizeof(float) * (size_t)d.N, (size_t)d.M, stream);
      if (e != cudaSuccess) return cuda_status(e, "mmt_gemm16 split-K memset");
    }
  }
  CUtensorMap ma, mb;
  const int bo = d.batch / d.batch_inner;
  int rc = make_map16(&ma, d.A, d.M, d.K, d.a_ld, d.a_mn != 0, BM, bo, d.batch_inner, d.a_bs0, d.a_bs1, d.dtype, "A");
  if (rc) return rc;
  rc = make_map16(&mb, d.B, d.N, d.K, d.b_ld, d.b_mn != 0, BN / 2, bo, d.batch_inner, d.b_bs0, d.b_bs1, d.dtype, "B");
  if (rc) return rc;
  // 16-bit-only outputs take the 64-column epilogue (16-byte accesses, complete 128-byte lines)
  const bool out16 = d.C16 && !d.C32 && !d.add && d.p_drop == 0.f && args.split_k == 1 && (d.N % 8) == 0 &&
                     ((d.c16_ld | d.c_bs0 | d.c_bs1 | d.bias_bs | d.colsum_bs) % 8) == 0 && ((uintptr_t)d.C16 % 16) == 0 &&
                     (!d.aux16 || ((d.aux_ld % 8) == 0 && ((uintptr_t)d.aux16 % 16) == 0)) &&
                     (!d.bias || ((uintptr_t)d.bias % 16) == 0) && (!d.colsum || ((uintptr_t)d.colsum % 16) == 0);
  {
    const void* fn = out16 ? (BN == 256 ? (const void*)gemm16_kernel<true, 256> : (const void*)gemm16_kernel<true, 128>)
                           : (BN == 256 ? (const void*)gemm16_kernel<false, 256> : (const void*)gemm16_kernel<false, 128>);
    rc = ensure_dynamic_smem(fn, SMEM_BYTES, "mmt_gemm16 smem attribute");
    if (rc) return rc;
  }
  const int work = tiles * args.split_k * d.batch;
  const int pairs = work < max_pairs ? work : max_pairs;
  CUtensorMap mc = ma, mx = ma;                              // placeholders unless the TMA epilogue runs
  if (out16) {
    rc = make_map16(&mc, d.C16, d.M, d.N, d.c16_ld, false, 32, bo, d.batch_inner, d.c_bs0, d.c_bs1, d.dtype, "C16", true);
    if (rc) return rc;
    if (d.aux16) {
      rc = make_map16(&mx, d.aux16, d.M, d.N, d.aux_ld, false, 32, bo, d.batch_inner, d.c_bs0, d.c_bs1, d.dtype, "aux16", true);
      if (rc) return rc;
    }
  }
  if (out16 && BN == 256) launch_pdl(gemm16_kernel<true, 256>, dim3(2 * pairs), dim3(NUM_THREADS), SMEM_BYTES, stream, ma, mb, mc, mx, args);
  else if (out16) launch_pdl(gemm16_kernel<true, 128>, dim3(2 * pairs), dim3(NUM_THREADS), SMEM_BYTES, stream, ma, mb, mc, mx, args);
  else if (BN == 256) launch_pdl(gemm16_kernel<false, 256>, dim3(2 * pairs), dim3(NUM_THREADS), SMEM_BYTES, stream, ma, mb, mc, mx, args);
  else launch_pdl(gemm16_kernel<false, 128>, dim3(2 * pairs), dim3(NUM_THREADS), SMEM_BYTES, stream, ma, mb, mc, mx, args);
  MMT_LAUNCH_CHECK("gemm16_kernel");
  return 0;
}
