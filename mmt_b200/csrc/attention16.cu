// Fused self-attention on 16-bit operands (model/bert.py:136-172 and its autograd), tcgen05 + TMA.
// Nothing of size S x S ever reaches HBM: the forward keeps scores / probabilities in TMEM and saves only
// the per-row log-sum-exp; the backward recomputes the probabilities from (Q, K, lse) and regenerates the
// dropout decisions from (seed, site).
//
// FORWARD  one CTA per (b, h, 128-query tile), two CTAs per SM (88 KB smem, 256 TMEM columns each), so one
//          CTA's loads and MMAs run under the other's softmax:
//   warp 0    TMA: Q tile [128 x dh] and K block [224 x dh] (K-major, 128-byte swizzle); once the score MMAs
//             have drained K, V [224 x dh] into the SAME buffer (the identical bytes are an MN-major B operand)
//   warp 1    MMA: S = Q K^T (kind::f16, M=128, N=224, fp32 in TMEM), then O += P V with P read from TMEM
//   warps 2-5 softmax, one thread per query row: scale + additive mask, max, exp2, row sum, dropout, P
//             written IN PLACE over S as packed 16-bit pairs (unnormalised, in (0,1]); finally
//             O * inv_keep / l -> ctx16, lse.  S > 224 takes several key blocks with the online-softmax
//             rescale of the TMEM accumulator (then 512 TMEM columns, one CTA per SM).
// BACKWARD one CTA per (b, h, 128-key tile), keys on the TMEM lanes; loop over 128-query tiles i:
//   S^T = K_j Q_i^T and dPd^T = V_j dO_i^T                          (TMEM, 2 x 128 columns)
//   8 softmax warps (two threads per key row): P^T = exp2(S^T*c + mask_k - lse_q), Pd^T = keep * P^T / (1-p),
//   dS^T = P^T * (keep * dPd^T / (1-p) - delta_q) * scale  ->  Pd^T and dS^T as 16-bit tiles in shared memory
//   (row = key, 128-byte swizzle: the same bytes serve as K-major A [keys x queries] and MN-major A [queries x keys])
//   dV_j += Pd^T dO_i,  dK_j += dS^T Q_i  (TMEM accumulators across i),  dQ_i = dS K_j -> fp32 red.global.add
//   into dq32 (each element receives one addend per key tile; two for S <= 256: order-independent).
// delta_q = dO_q . O_q comes from a small pre-kernel; a post-kernel turns dq32 into the 16-bit dQ block of
// dqkv16, takes its bias-gradient column sums and leaves dq32 zeroed for the next layer.
//
// Attention-probability dropout uses one 32-bit integer hash per PAIR of adjacent keys (16 random bits per
// decision, P(drop) = floor(p * 65536) / 65536): with Philox the RNG alone was more than half of the softmax
// instructions, and the softmax -- not the tensor core -- bounds this kernel at dh = 128.
#include "cvt16.cuh"
#include "rowvec.cuh"
#include "tc_ptx.cuh"
#include "pair_ptx.cuh"

namespace mmt {
namespace {
using namespace tc;

constexpr int DH = 128;
constexpr float LOG2E = 1.44269504088896340736f;

// ---- PTX helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint32_t idesc16(int m, int n, bool a_mn, bool b_mn, bool bf16) {
  const uint32_t f = bf16 ? 1u : 0u;
  return (1u << 4) | (f << 7) | (f << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16u(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld16u(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
        "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// ---- dropout decisions: one hash per pair of adjacent keys ------------------------------------------
// random word of the key pair (key >> 1) of probability row `prow` (= (b*H + h)*S + q); low half -> even key
__device__ __forceinline__ uint32_t drop_word(uint32_t key32, uint32_t prow, uint32_t half_pitch, uint32_t kpair) {
  return hash32((prow * half_pitch + kpair) ^ key32);
}

struct AttArgs {
  const float* mask;      // [B, S] 1 = attend
  uint16_t* ctx16;        // [B*S, H*DH]
  float* lse;             // [B, H, S]
  int B, H, S;
  float scale_log2;       // (1/sqrt(dh)) * log2(e)
  float p_drop, inv_keep;
  uint64_t seed;
  const uint64_t* ctr;
  uint32_t site;
  int bf16;
};

// ======================================== forward ====================================================
namespace fwd {
constexpr int QM = 128, KB = 224;
constexpr int THREADS = 192;                         // TMA warp, MMA warp, 4 softmax warps
constexpr uint32_t Q_BYTES = QM * DH * 2;            // 32 KB: 2 sub-tiles [128 rows x 128 B]
constexpr uint32_t QSUB = QM * 128;
constexpr uint32_t KV_BYTES = KB * DH * 2;           // 56 KB: 2 sub-tiles [224 rows x 128 B]
constexpr uint32_t KSUB = KB * 128;
constexpr size_t SMEM = Q_BYTES + KV_BYTES + 1024 /*align*/ + 128 /*barriers*/ + KB * 4;

__global__ void __launch_bounds__(THREADS, 2) attention16_fwd_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                     const __grid_constant__ CUtensorMap map_k,
                                                                     const __grid_constant__ CUtensorMap map_o,
                                                                     const AttArgs args) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = smem;
  uint8_t* skv = smem + Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_BYTES + KV_BYTES);
  uint64_t* qk_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* k_free = bars + 2;              // score MMAs have finished reading K
  uint64_t* v_free = bars + 3;              // P V MMAs have finished reading V (and P)
  uint64_t* s_full = bars + 4;
  uint64_t* p_full = bars + 5;              // 128 arrivals
  uint64_t* o_full = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* smask = reinterpret_cast<float*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, S = args.S;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * QM;
  const int nblk = (S + KB - 1) / KB;
  const int d_model = H * DH;
  const bool bf16 = args.bf16 != 0;
  const uint32_t tm_cols = nblk == 1 ? 256u : 512u;   // single key block: O reuses the upper score columns
  const uint32_t TM_S = 0, TM_O = nblk == 1 ? 128u : 256u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], i == 5 ? 128 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tm_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int row_q = b * S + q0;
      for (int j = 0; j < nblk; ++j) {
        const int row_k = b * S + j * KB;
        mbar_wait(v_free, (j & 1) ^ 1);
        mbar_arrive_expect_tx(qk_full, (j == 0 ? Q_BYTES : 0) + KV_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j == 0) tma_load_2d(sq + t * QSUB, &map_q, qk_full, h * DH + 64 * t, row_q);
          tma_load_2d(skv + t * KSUB, &map_k, qk_full, d_model + h * DH + 64 * t, row_k);
        }
        mbar_wait(k_free, j & 1);
        mbar_arrive_expect_tx(v_full, KV_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t) tma_load_2d(skv + t * KSUB, &map_k, v_full, 2 * d_model + h * DH + 64 * t, row_k);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = idesc16(QM, KB, false, false, bf16);      // S[128 x 224] = Q K^T
      const uint32_t idesc_o = idesc16(QM, DH, false, true, bf16);       // O[128 x dh] += P V (V MN-major)
      for (int j = 0; j < nblk; ++j) {
        if (j > 0) mbar_wait(o_full, (j - 1) & 1);            // P (aliasing S) no longer read
        mbar_wait(qk_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          const uint64_t da = make_smem_desc(smem_u32(sq) + (ks >> 2) * QSUB + (ks & 3) * 32, 16, 1024, 2);
          const uint64_t db = make_smem_desc(smem_u32(skv) + (ks >> 2) * KSUB + (ks & 3) * 32, 16, 1024, 2);
          umma_f16_ss(tmem + TM_S, da, db, idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(k_free);
        umma_commit(s_full);
        mbar_wait(v_full, j & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < KB / 16; ++ks) {
          // V as MN-major B: 64-wide dh chunks KSUB apart (LBO), 8-key atoms 1024 B apart (SBO), 16 keys = 2048 B
          const uint64_t db = make_smem_desc(smem_u32(skv) + ks * 2048, KSUB, 1024, 2);
          umma_f16_ts(tmem + TM_O, tmem + TM_S + ks * 8, db, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(v_free);
        umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / epilogue (warps 2..5), thread = query row =====================
    const int q = warp & 3;                                    // TMEM lane quarter of this warp
    const int r = q * 32 + lane;
    const int qi = q0 + r;
    const bool row_ok = qi < S;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float* mrow = args.mask + (int64_t)b * S;
    const uint32_t prow = (uint32_t)(((int64_t)b * H + h) * S + qi);
    const uint32_t half_pitch = (uint32_t)((S + 1) >> 1);
    const uint64_t seed = args.seed + (args.ctr ? *args.ctr : 0);
    const uint32_t key32 = drop_key(seed, args.site);
    const uint32_t thr = (uint32_t)(args.p_drop * 65536.0f);
    const bool drop = args.p_drop > 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int key0 = j * KB;
      for (int t = threadIdx.x - 64; t < KB; t += 128) {
        const int key = key0 + t;
        smask[t] = key < S ? (1.0f - __ldg(mrow + key)) * (-10000.0f * LOG2E) : -INFINITY;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max of the masked, scaled scores (log2 domain)
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < KB / 32; ++c) {
        float v[32];
        tmem_ld32(tmem + TM_S + lane_addr + c * 32, v);
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
          const float4 mk = *reinterpret_cast<const float4*>(smask + c * 32 + t);
          m_blk = fmaxf(fmaxf(m_blk, fmaf(v[t], args.scale_log2, mk.x)), fmaf(v[t + 1], args.scale_log2, mk.y));
          m_blk = fmaxf(fmaxf(m_blk, fmaf(v[t + 2], args.scale_log2, mk.z)), fmaf(v[t + 3], args.scale_log2, mk.w));
        }
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = (j == 0) ? 0.f : fast_ex2(m_run - m_new);
      // pass 2: p = 2^(x - m), row sum of the un-dropped p, dropout, packed 16-bit P over S: columns 16c..16c+15
      // receive the 32 probabilities of score columns 32c..32c+31, all of which this thread has already read
      float l_blk = 0.f;
#pragma unroll 1
      for (int c = 0; c < KB / 32; ++c) {
        float v[32];
        tmem_ld32(tmem + TM_S + lane_addr + c * 32, v);
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
          const float2 mk = *reinterpret_cast<const float2*>(smask + c * 32 + t);
          float p0 = fast_ex2(fmaf(v[t], args.scale_log2, mk.x) - m_new);
          float p1 = fast_ex2(fmaf(v[t + 1], args.scale_log2, mk.y) - m_new);
          l_blk += p0 + p1;
          if (drop) {
            const uint32_t w = drop_word(key32, prow, half_pitch, (uint32_t)((key0 + c * 32 + t) >> 1));
            if ((w & 0xffffu) < thr) p0 = 0.f;
            if ((w >> 16) < thr) p1 = 0.f;
          }
          pk[t >> 1] = pack2(p0, p1, bf16);
        }
        tmem_st16u(tmem + TM_S + lane_addr + c * 16, pk);
      }
      if (j > 0) {
        // online softmax: rescale the running accumulator (previous P V has landed: o_full(j-1))
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < DH / 16; ++c) {
          float v[16];
          tmem_ld16f(tmem + TM_O + lane_addr + c * 16, v);
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] *= alpha;
          tmem_st16f(tmem + TM_O + lane_addr + c * 16, v);
        }
      }
      tmem_st_wait();
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      tc_fence_before();
      mbar_arrive(p_full);
      asm volatile("bar.sync 1, 128;" ::: "memory");          // smask is rewritten by the next block
    }
    // epilogue: O * inv_keep / l -> ctx16, log-sum-exp for the backward pass
    mbar_wait(o_full, (nblk - 1) & 1);
    tc_fence_after();
    const float inv_l = args.inv_keep / l_run;
    // context tile -> the Q buffer (dead: every score MMA has completed), 128-byte-swizzled, -> two TMA stores clipped at
    // S by the [B, S, d] map.  (Per-lane 16-byte stores to rows 1 KB apart take 32 LSU wavefronts per instruction.)
    const uint32_t sq_u = smem_u32(sq);
    const uint32_t rrow = (uint32_t)r;                          // query row within the tile == TMEM lane
#pragma unroll 1
    for (int c = 0; c < DH / 16; ++c) {
      float v[16];
      tmem_ld16f(tmem + TM_O + lane_addr + c * 16, v);
      const uint32_t rb = sq_u + (uint32_t)(c >> 2) * QSUB + rrow * 128;
      const uint32_t ch0 = (uint32_t)((c & 3) * 2), sw = rrow & 7;
      sts128u(rb + ((ch0 ^ sw) << 4), pack2(v[0] * inv_l, v[1] * inv_l, bf16), pack2(v[2] * inv_l, v[3] * inv_l, bf16),
              pack2(v[4] * inv_l, v[5] * inv_l, bf16), pack2(v[6] * inv_l, v[7] * inv_l, bf16));
      sts128u(rb + (((ch0 + 1) ^ sw) << 4), pack2(v[8] * inv_l, v[9] * inv_l, bf16), pack2(v[10] * inv_l, v[11] * inv_l, bf16),
              pack2(v[12] * inv_l, v[13] * inv_l, bf16), pack2(v[14] * inv_l, v[15] * inv_l, bf16));
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 64) {
      tma_store_3d(&map_o, sq_u, h * DH, q0, b);
      tma_store_3d(&map_o, sq_u + QSUB, h * DH + 64, q0, b);
      bulk_commit();
      bulk_wait_read0();
    }
    if (row_ok && args.lse) args.lse[((int64_t)b * H + h) * S + qi] = (m_run + log2f(l_run)) * 0.69314718055994530942f;
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, tm_cols);
  }
}
}  // namespace fwd

// ======================================== backward ===================================================
struct AttBwdArgs {
  const float* mask;      // [B, S]
  const float* lse;       // [B, H, S] natural log
  const float* delta;     // [B, H, S] dO . O in the scale16 domain
  uint16_t* dqkv16;       // [B*S, 3*H*DH]; this kernel writes the K and V blocks
  float* dq32;            // [B*S, H*DH] fp32, accumulated with red.add
  float* dbias;           // [3*H*DH] fp32 bias gradient (K and V blocks accumulated here)
  int B, H, S;
  float scale, scale_log2;
  float p_drop, inv_keep;
  float inv_scale16;
  uint64_t seed;
  const uint64_t* ctr;
  uint32_t site;
  int bf16;
  int dq_mode;            // 0: fp32 atomics into dq32 + attn_dq_finish_kernel (any S); 1: one key tile, dQ final in-kernel;
                          // 2: two key tiles = a 2-CTA cluster, partial handed over through dq32 (see the kernel)
  int debug;              // timing experiments (MMT_ATT_BWD_DEBUG): 1 no dQ atomics, 2 no softmax math, 4 no dV/dK stores
};

namespace bwd {
constexpr int KT = 128, QT = 128;
// TMA warp, MMA warp, NSW softmax warps (8 or 16: PARTS = NSW / 4 warps share a TMEM lane quarter and split the columns)
constexpr uint32_t TILE = 128 * DH * 2;              // 32 KB: 2 sub-tiles [128 rows x 128 B]
constexpr uint32_t SUB = 128 * 128;                  // 16 KB
constexpr size_t SMEM = 6 * TILE + 1024 /*align*/ + 128 /*barriers*/ + 2 * QT * 4;
constexpr uint32_t TM_ST = 0, TM_DP = 128, TM_DV = 256, TM_DK = 384, TM_DQ = 0;

// column sums of a [32 rows (lanes) x 16 columns] register block by a halving butterfly (16 + 8 + 4 + 2 + 1 shuffles);
// lane L (even) ends up with column bitrev4(L >> 1) and adds it to dst
__device__ __forceinline__ void warp_colsum16(float (&v)[16], int lane, float* dst, float scale) {
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float send = (lane & 16) ? v[t] : v[t + 8], keep = (lane & 16) ? v[t + 8] : v[t];
    v[t] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float send = (lane & 8) ? v[t] : v[t + 4], keep = (lane & 8) ? v[t + 4] : v[t];
    v[t] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const float send = (lane & 4) ? v[t] : v[t + 2], keep = (lane & 4) ? v[t + 2] : v[t];
    v[t] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const float send = (lane & 2) ? v[0] : v[1], keep = (lane & 2) ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  if ((lane & 1) == 0) {
    const int colid = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    atomicAdd(dst + colid, v[0] * scale);
  }
}

template <int NSW>
__global__ void __launch_bounds__(64 + NSW * 32, 1) attention16_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv,
                                                                     const __grid_constant__ CUtensorMap map_do,
                                                                     const __grid_constant__ CUtensorMap map_out,
                                                                     const __grid_constant__ CUtensorMap map_dq,
                                                                     const AttBwdArgs args) {
  constexpr int SMT = NSW * 32;                       // softmax threads
  constexpr int PARTS = NSW / 4;                      // column parts per lane quarter
  constexpr int PCOLS = 128 / PARTS;                  // columns of a 128-wide accumulator per part
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sk = smem;
  uint8_t* sv = smem + TILE;
  uint8_t* sq = smem + 2 * TILE;
  uint8_t* sdo = smem + 3 * TILE;
  uint8_t* sp = smem + 4 * TILE;                     // Pd^T  [keys x queries]
  uint8_t* sds = smem + 5 * TILE;                    // dS^T  [keys x queries]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;
  uint64_t* st_full = bars + 2;              // S^T and dPd^T in TMEM
  uint64_t* pds_full = bars + 3;             // Pd^T / dS^T tiles written (SMT arrivals)
  uint64_t* mma2_done = bars + 4;            // dV / dK / dQ MMAs of this query tile complete
  uint64_t* dq_drained = bars + 5;           // dQ read out of TMEM (SMT arrivals)
  uint64_t* xch_bar = bars + 6;              // dq_mode 2: the peer CTA's dQ partial of a query tile is in dq32 (one remote arrival)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_lse2 = reinterpret_cast<float*>(bars + 16);       // [QT] log2-domain lse of the query tile (+inf: no such query)
  float* s_delta = s_lse2 + QT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, S = args.S;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * KT;
  const int nqt = (S + QT - 1) / QT;
  const int d_model = H * DH;
  const bool bf16 = args.bf16 != 0;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1); mbar_init(qdo_full, 1); mbar_init(st_full, 1);
    mbar_init(pds_full, SMT); mbar_init(mma2_done, 1); mbar_init(dq_drained, SMT);
    mbar_init(xch_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (args.dq_mode == 2) cluster_sync_all();          // the peer's xch_bar is initialised before anyone arrives on it
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int row_k = b * S + k0;
      mbar_arrive_expect_tx(kv_full, 2 * TILE);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        tma_load_2d(sk + t * SUB, &map_qkv, kv_full, d_model + h * DH + 64 * t, row_k);
        tma_load_2d(sv + t * SUB, &map_qkv, kv_full, 2 * d_model + h * DH + 64 * t, row_k);
      }
      for (int i = 0; i < nqt; ++i) {
        if (i > 0) mbar_wait(mma2_done, (i - 1) & 1);          // Q_i / dO_i of the previous tile no longer read
        const int row_q = b * S + i * QT;
        mbar_arrive_expect_tx(qdo_full, 2 * TILE);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          tma_load_2d(sq + t * SUB, &map_qkv, qdo_full, h * DH + 64 * t, row_q);
          tma_load_2d(sdo + t * SUB, &map_do, qdo_full, h * DH + 64 * t, row_q);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_kk = idesc16(128, 128, false, false, bf16);   // A K-major, B K-major
      const uint32_t id_km = idesc16(128, 128, false, true, bf16);    // A K-major, B MN-major
      const uint32_t id_mm = idesc16(128, 128, true, true, bf16);     // A MN-major, B MN-major
      mbar_wait(kv_full, 0);
      for (int i = 0; i < nqt; ++i) {
        mbar_wait(qdo_full, i & 1);
        if (i > 0) mbar_wait(dq_drained, (i - 1) & 1);         // dQ_{i-1} has left the S^T columns
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {                  // S^T = K_j Q_i^T, dPd^T = V_j dO_i^T
          const uint32_t off = (ks >> 2) * SUB + (ks & 3) * 32;
          umma_f16_ss(tmem + TM_ST, make_smem_desc(smem_u32(sk) + off, 16, 1024, 2),
                      make_smem_desc(smem_u32(sq) + off, 16, 1024, 2), id_kk, ks > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          const uint32_t off = (ks >> 2) * SUB + (ks & 3) * 32;
          umma_f16_ss(tmem + TM_DP, make_smem_desc(smem_u32(sv) + off, 16, 1024, 2),
                      make_smem_desc(smem_u32(sdo) + off, 16, 1024, 2), id_kk, ks > 0 ? 1u : 0u);
        }
        umma_commit(st_full);
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < QT / 16; ++ks) {                  // dV_j += Pd^T dO_i ; dK_j += dS^T Q_i  (k = queries)
          const uint32_t aoff = (ks >> 2) * SUB + (ks & 3) * 32;
          umma_f16_ss(tmem + TM_DV, make_smem_desc(smem_u32(sp) + aoff, 16, 1024, 2),
                      make_smem_desc(smem_u32(sdo) + ks * 2048, SUB, 1024, 2), id_km, (i > 0 || ks > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < QT / 16; ++ks) {
          const uint32_t aoff = (ks >> 2) * SUB + (ks & 3) * 32;
          umma_f16_ss(tmem + TM_DK, make_smem_desc(smem_u32(sds) + aoff, 16, 1024, 2),
                      make_smem_desc(smem_u32(sq) + ks * 2048, SUB, 1024, 2), id_km, (i > 0 || ks > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks)                    // dQ_i = dS K_j  (k = keys; both operands MN-major)
          umma_f16_ss(tmem + TM_DQ, make_smem_desc(smem_u32(sds) + ks * 2048, SUB, 1024, 2),
                      make_smem_desc(smem_u32(sk) + ks * 2048, SUB, 1024, 2), id_mm, ks > 0 ? 1u : 0u);
        umma_commit(mma2_done);
      }
    }
  } else {
    // ===================== softmax backward + epilogues (warps 2 .. 2+NSW) =====================
    const int q4 = warp & 3;                                   // TMEM lane quarter
    const int part = (warp - 2) >> 2;                          // which PCOLS of the 128 columns
    const int r = q4 * 32 + lane;                              // key row within the tile == TMEM lane
    const int key = k0 + r;
    const bool key_ok = key < S;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const float mterm = key_ok ? (1.0f - __ldg(args.mask + (int64_t)b * S + key)) * (-10000.0f * LOG2E) : -INFINITY;
    const uint64_t seed = args.seed + (args.ctr ? *args.ctr : 0);
    const uint32_t key32 = drop_key(seed, args.site);
    const uint32_t thr = (uint32_t)(args.p_drop * 65536.0f);
    const bool drop = args.p_drop > 0.f;
    const uint32_t half_pitch = (uint32_t)((S + 1) >> 1);
    const uint32_t kpair = (uint32_t)key >> 1;
    const uint32_t kshift = (key & 1) ? 16u : 0u;
    const uint32_t prow0 = (uint32_t)(((int64_t)b * H + h) * S);
    // Pd^T / dS^T tiles: sub-tile = 64 queries, row = key, 16-byte chunk c of a row stored at c ^ (row & 7)
    const uint32_t sp_u = smem_u32(sp), sds_u = smem_u32(sds);
    uint32_t xphase = 0;
    for (int i = 0; i < nqt; ++i) {
      const int qbase = i * QT;
      // per-query vectors of this tile (the previous tile's readers are past pds_full of i-1 ... and its dQ drain)
      for (int t = threadIdx.x - 64; t < QT; t += SMT) {
        const int qq = qbase + t;
        const bool ok = qq < S;
        s_lse2[t] = ok ? __ldg(args.lse + ((int64_t)b * H + h) * S + qq) * LOG2E : INFINITY;
        s_delta[t] = ok ? __ldg(args.delta + ((int64_t)b * H + h) * S + qq) : 0.f;
      }
      if (threadIdx.x == 64 && i > 0) bulk_wait_read0();      // dQ_{i-1}'s TMA reduction has read the Pd^T / dS^T buffers
      asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");
      mbar_wait(st_full, i & 1);
      tc_fence_after();
      // 16 query columns at a time; the column parts take interleaved 16-column chunks so that a ragged last
      // query tile (S = 218: 90 of 128 columns) leaves them with (nearly) the same amount of work
#pragma unroll 1
      for (int cc = part; cc < 8; cc += PARTS) {
        const int col = cc * 16;
        const uint32_t rb = (uint32_t)(cc >> 2) * SUB + (uint32_t)r * 128;      // sub-tile (64 queries), this key's row
        const uint32_t ch0 = (uint32_t)((cc & 3) * 2), sw = (uint32_t)(r & 7);
        if (qbase + col >= S || k0 + q4 * 32 >= S || (args.debug & 2)) {   // no such queries / no such keys in this warp: zeros
          sts128u(sp_u + rb + ((ch0 ^ sw) << 4), 0u, 0u, 0u, 0u);
          sts128u(sp_u + rb + (((ch0 + 1) ^ sw) << 4), 0u, 0u, 0u, 0u);
          sts128u(sds_u + rb + ((ch0 ^ sw) << 4), 0u, 0u, 0u, 0u);
          sts128u(sds_u + rb + (((ch0 + 1) ^ sw) << 4), 0u, 0u, 0u, 0u);
          continue;
        }
        float sv_[16], dp[16];
        {
          uint32_t r0[16], r1[16];
          tmem_ld16u(tmem + TM_ST + lane_addr + col, r0);
          tmem_ld16u(tmem + TM_DP + lane_addr + col, r1);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 16; ++t) { sv_[t] = __uint_as_float(r0[t]); dp[t] = __uint_as_float(r1[t]); }
        }
        uint32_t pp[8], dd[8];
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int qc = col + t + u;
            const float p = fast_ex2(fmaf(sv_[t + u], args.scale_log2, mterm) - s_lse2[qc]);
            float keep = args.inv_keep;
            if (drop) {
              const uint32_t w = drop_word(key32, prow0 + (uint32_t)(qbase + qc), half_pitch, kpair);
              if (((w >> kshift) & 0xffffu) < thr) keep = 0.f;
            }
            pv[u] = p * keep;
            dv[u] = p * fmaf(dp[t + u], keep, -s_delta[qc]) * args.scale;
          }
          pp[t >> 1] = pack2(pv[0], pv[1], bf16);
          dd[t >> 1] = pack2(dv[0], dv[1], bf16);
        }
        sts128u(sp_u + rb + ((ch0 ^ sw) << 4), pp[0], pp[1], pp[2], pp[3]);
        sts128u(sp_u + rb + (((ch0 + 1) ^ sw) << 4), pp[4], pp[5], pp[6], pp[7]);
        sts128u(sds_u + rb + ((ch0 ^ sw) << 4), dd[0], dd[1], dd[2], dd[3]);
        sts128u(sds_u + rb + (((ch0 + 1) ^ sw) << 4), dd[4], dd[5], dd[6], dd[7]);
      }
      fence_proxy_async_smem();                                 // generic-proxy smem writes -> visible to the MMAs
      tc_fence_before();
      mbar_arrive(pds_full);
      // dQ_i out of TMEM: lane = query row, this thread's PCOLS of the dh columns.
      //   dq_mode 0 (more than two key tiles): fp32 atomics into the zeroed dq32; attn_dq_finish_kernel rounds it.
      //   dq_mode 2 (two key tiles = the two CTAs of a cluster): the CTAs take turns -- for query tile i, CTA (i+1)&1
      //     stores its partial to dq32 (plain stores, L2) and arrives on the other's xch_bar; CTA i&1 adds that to its
      //     own TMEM partial and writes the final 16-bit dQ and the Q-bias column sums.  No zeroing, no atomics, no
      //     extra kernel (the finish kernel cost 20 us per layer: 71 MB of traffic for 14 MB of result).
      //   dq_mode 1 (one key tile): final directly.
      mbar_wait(mma2_done, i & 1);
      tc_fence_after();
      {
        const int qq = qbase + r;
        const bool q_ok = qq < S;
        float* drow = args.dq32 + ((int64_t)b * S + qq) * d_model + h * DH + part * PCOLS;
        const bool finalize = args.dq_mode == 1 || (args.dq_mode == 2 && (i & 1) == kt);
        if (!finalize && args.dq_mode == 0) {
          // fp32 [128 queries x 128] tile staged in the Pd^T / dS^T buffers (free until the next softmax; four 128-byte-
          // swizzled boxes of 32 columns) and ADDED to dq32 by TMA (cp.reduce.async.bulk: whole lines, clipped at S).
          // Per-lane float4 atomics on rows 2 KB apart cost 32 LSU wavefronts per instruction: ~2 us per tile.
#pragma unroll
          for (int c = 0; c < PCOLS / 16; ++c) {
            float v[16];
            tmem_ld16f(tmem + TM_DQ + lane_addr + part * PCOLS + c * 16, v);
            const int col = part * PCOLS + c * 16;
            const uint32_t rb = sp_u + (uint32_t)(col >> 5) * (128 * 128) + (uint32_t)r * 128;
            const uint32_t ch0 = (uint32_t)((col & 31) >> 2), sw = (uint32_t)(r & 7);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              sts128(rb + (((ch0 + t) ^ sw) << 4), make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]));
          }
          tc_fence_before();
          mbar_arrive(dq_drained);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");
          if (threadIdx.x == 64 && !(args.debug & 1)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_reduce_add_3d(&map_dq, sp_u + j * (128 * 128), h * DH + 32 * j, qbase, b);
            bulk_commit();
          }
        } else if (!finalize) {
#pragma unroll 1
          for (int c = 0; c < PCOLS / 16; ++c) {
            float v[16];
            tmem_ld16f(tmem + TM_DQ + lane_addr + part * PCOLS + c * 16, v);
            if (q_ok && !(args.debug & 1)) {
#pragma unroll
              for (int t = 0; t < 16; t += 4)
                __stcg(reinterpret_cast<float4*>(drow + c * 16 + t), make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]));
            }
          }
          tc_fence_before();
          mbar_arrive(dq_drained);
          if (args.dq_mode == 2) {                              // one cumulative release for the SMT writers
            asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");
            if (threadIdx.x == 64) mbar_arrive_remote_release(xch_bar, (uint32_t)(kt ^ 1));
          }
        } else {
          uint16_t* qrow = args.dqkv16 + ((int64_t)b * S + qq) * (3 * d_model) + h * DH + part * PCOLS;
          float* bsum = args.dbias + h * DH + part * PCOLS;
          if (args.dq_mode == 2) {
            mbar_wait_acquire_cluster(xch_bar, xphase);
            xphase ^= 1;
          }
          // the peer's partial first (all 16-byte loads in flight at once: one L2 latency, not one per chunk)
          float4 pz[PCOLS / 4];
          if (args.dq_mode == 2 && q_ok) {
#pragma unroll
            for (int t = 0; t < PCOLS / 4; ++t) pz[t] = __ldcg(reinterpret_cast<const float4*>(drow + 4 * t));
          } else {
#pragma unroll
            for (int t = 0; t < PCOLS / 4; ++t) pz[t] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int c = 0; c < PCOLS / 16; ++c) {
            float v[16];
            tmem_ld16f(tmem + TM_DQ + lane_addr + part * PCOLS + c * 16, v);
            if (q_ok) {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                v[4 * t] += pz[4 * c + t].x; v[4 * t + 1] += pz[4 * c + t].y;
                v[4 * t + 2] += pz[4 * c + t].z; v[4 * t + 3] += pz[4 * c + t].w;
              }
              uint4 o0, o1;
              o0.x = pack2(v[0], v[1], bf16);   o0.y = pack2(v[2], v[3], bf16);
              o0.z = pack2(v[4], v[5], bf16);   o0.w = pack2(v[6], v[7], bf16);
              o1.x = pack2(v[8], v[9], bf16);   o1.y = pack2(v[10], v[11], bf16);
              o1.z = pack2(v[12], v[13], bf16); o1.w = pack2(v[14], v[15], bf16);
              *reinterpret_cast<uint4*>(qrow + c * 16) = o0;
              *reinterpret_cast<uint4*>(qrow + c * 16 + 8) = o1;
            } else {
#pragma unroll
              for (int t = 0; t < 16; ++t) v[t] = 0.f;
            }
            warp_colsum16(v, lane, bsum + c * 16, args.inv_scale16);
          }
          tc_fence_before();
          mbar_arrive(dq_drained);
        }
      }
    }
    // epilogue: dV_j, dK_j (lane = key row) -> 128-byte-swizzled [128 keys x 64 columns] smem tiles in the (now free)
    // Q / dO buffers -> TMA stores into the V and K blocks of dqkv16, clipped at S by the [B, S, 3d] map.  Row-per-lane
    // global stores (3 KB apart) cost ~20 us per launch here.  Bias-gradient column sums from the same registers.
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t tm = which == 0 ? TM_DV : TM_DK;
      const int blk = which == 0 ? 2 : 1;                       // column block of dqkv16: Q | K | V
      const uint32_t tile_u = smem_u32(which == 0 ? sq : sdo);
      float* bsum = args.dbias + blk * d_model + h * DH + part * PCOLS;
#pragma unroll 1
      for (int c = 0; c < PCOLS / 16; ++c) {
        float v[16];
        tmem_ld16f(tmem + tm + lane_addr + part * PCOLS + c * 16, v);
        if (!(args.debug & 4)) {
          const int col = part * PCOLS + c * 16;
          const uint32_t rb = tile_u + (uint32_t)(col >> 6) * SUB + (uint32_t)r * 128;
          const uint32_t ch0 = (uint32_t)((col & 63) >> 3), sw = (uint32_t)(r & 7);
          sts128u(rb + ((ch0 ^ sw) << 4), pack2(v[0], v[1], bf16), pack2(v[2], v[3], bf16), pack2(v[4], v[5], bf16),
                  pack2(v[6], v[7], bf16));
          sts128u(rb + (((ch0 + 1) ^ sw) << 4), pack2(v[8], v[9], bf16), pack2(v[10], v[11], bf16),
                  pack2(v[12], v[13], bf16), pack2(v[14], v[15], bf16));
        }
        if (args.debug & 8) continue;
        if (!key_ok) {
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = 0.f;
        }
        warp_colsum16(v, lane, bsum + c * 16, args.inv_scale16);
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");
    if (threadIdx.x == 64) {
      if (!(args.debug & 4)) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          tma_store_3d(&map_out, smem_u32(sq) + t * SUB, 2 * d_model + h * DH + 64 * t, k0, b);
          tma_store_3d(&map_out, smem_u32(sdo) + t * SUB, d_model + h * DH + 64 * t, k0, b);
        }
        bulk_commit();
      }
      bulk_wait_read0();                                      // smem may go once every bulk store / reduction has read it
    }
    tc_fence_before();
  }
  __syncthreads();
  if (args.dq_mode == 2) cluster_sync_all();          // no CTA retires while its peer may still arrive on its xch_bar
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// delta[b,h,q] = sum_c dO[b,q,h*dh+c] * O[b,q,h*dh+c]   (one warp per token row, dh = 128: 8 elements per lane, two heads per pass)
__global__ void __launch_bounds__(256) attn_delta16_kernel(const uint16_t* __restrict__ dctx16, const uint16_t* __restrict__ ctx16,
                                                           int64_t rows, int S, int H, float* __restrict__ delta, int bf16) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int d_model = H * DH;
  // 16 lanes x 16 bytes per head: a warp takes two heads of a row per load instruction
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
    const int64_t b = r / S, q = r % S;
#pragma unroll 2
    for (int h0 = 0; h0 < H; h0 += 2) {
      const int h = h0 + (lane >> 4);
      const bool ok = h < H;
      const int64_t off = r * d_model + (ok ? h : 0) * DH + (lane & 15) * 8;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(dctx16 + off));
      const uint4 o = __ldg(reinterpret_cast<const uint4*>(ctx16 + off));
      const float2 a0 = unpack2(a.x, bf16 != 0), a1 = unpack2(a.y, bf16 != 0), a2 = unpack2(a.z, bf16 != 0), a3 = unpack2(a.w, bf16 != 0);
      const float2 o0 = unpack2(o.x, bf16 != 0), o1 = unpack2(o.y, bf16 != 0), o2 = unpack2(o.z, bf16 != 0), o3 = unpack2(o.w, bf16 != 0);
      float s = ((a0.x * o0.x + a0.y * o0.y) + (a1.x * o1.x + a1.y * o1.y)) + ((a2.x * o2.x + a2.y * o2.y) + (a3.x * o3.x + a3.y * o3.y));
#pragma unroll
      for (int sh = 8; sh >= 1; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
      if ((lane & 15) == 0 && ok) delta[(b * H + h) * S + q] = s;
    }
  }
}

// dq32 [rows, d] -> the Q block of dqkv16 (16-bit, already in the scale16 domain), its bias-gradient column sums
// (divided by scale16), and dq32 zeroed again for the next layer's accumulation.
// One CTA owns FIN_ROWS consecutive rows and ALL d columns: thread (x, y) walks rows y, y+2, ... of its float4 column
// x with four rows in flight, keeps the column sum in registers and the CTA issues ONE atomic per column -- the
// first version flushed per warp-row-group and put ~1200 atomics on each of 128 hot addresses.
constexpr int FIN_ROWS = 32;
__global__ void __launch_bounds__(256) attn_dq_finish_kernel(float4* __restrict__ dq32, int64_t rows, int d4,
                                                            uint16_t* __restrict__ dqkv16, float* __restrict__ dbias,
                                                            float inv_scale16, int bf16) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 red[256];
  const int x = threadIdx.x, y = threadIdx.y;                // blockDim = (d4 <= 128 ? d4 : 128, 256 / that)
  const int ny = blockDim.y;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c4 = x; c4 < d4; c4 += blockDim.x) {
    float4 acc = zero;
    const int64_t r0 = (int64_t)blockIdx.x * FIN_ROWS;
    for (int k = y; k < FIN_ROWS; k += 8 * ny) {               // 8 independent 16-byte loads in flight per thread
      float4 g[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t r = r0 + k + u * ny;
        g[u] = (k + u * ny < FIN_ROWS && r < rows) ? __ldcs(dq32 + r * d4 + c4) : zero;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t r = r0 + k + u * ny;
        if (k + u * ny < FIN_ROWS && r < rows) {
          dq32[r * d4 + c4] = zero;
          *reinterpret_cast<uint2*>(dqkv16 + r * 3 * (4 * (int64_t)d4) + 4 * c4) = pack4(g[u], bf16 != 0);
          acc.x += g[u].x; acc.y += g[u].y; acc.z += g[u].z; acc.w += g[u].w;
        }
      }
    }
    red[y * blockDim.x + x] = acc;
    __syncthreads();
    if (y == 0) {
      for (int w = 1; w < ny; ++w) {
        const float4 t = red[w * blockDim.x + x];
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      atomic_add4(dbias + 4 * c4, make_float4(acc.x * inv_scale16, acc.y * inv_scale16, acc.z * inv_scale16, acc.w * inv_scale16));
    }
    __syncthreads();
  }
}
}  // namespace bwd

// N consecutive TMEM columns of this thread's lane (N = 16 or 32)
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[16]) { tmem_ld16f(taddr, v); }

// ======================================== backward, pipelined =========================================
// Same math and operand layouts as bwd::attention16_bwd_kernel, restructured so that the softmax warps (the busiest
// resource: ~24 M warp-instructions per launch) never wait for a load or an MMA.  The query axis is walked in tiles of
// 64 with double-buffered Q / dO / Pd^T / dS^T tiles in shared memory and two S^T accumulators in TMEM:
//   TMEM (512 columns): dV 0..127 | dK 128..255 | S^T stage 0 / 1: 256..319 / 320..383 | dPd^T 384..447 | dQ^T 448..511
//   MMA warp     S^T(i+1) = K Q_{i+1}^T is issued BEFORE softmax(i) has finished (other TMEM stage); once softmax(i) has
//                arrived: dPd^T(i+1) = V dO_{i+1}^T, dV += Pd^T dO_i, dK += dS^T Q_i, dQ_i^T = K^T dS (M = dh, N = 64 queries)
//   softmax      p(i) from S^T while dPd^T(i) is still being computed, then dS(i); then dQ^T(i-1) (own TMEM columns, so the next
//                S^T never waits for this drain) -> fp32 staging -> TMA reduce-add into dq32.  (Direct red.global.add.f32 --
//                coalesced here, lane = dh -- was measured 14 us per launch slower: 32-bit atomics at the L2.)
//   TMA warp     Q_{i+2} / dO_{i+2} as soon as the MMAs of tile i have completed
// The previous kernel ran load -> MMA -> softmax -> MMA -> drain serially with one CTA per SM: of its 83 us, 45 us remained
// with ALL softmax math, dQ traffic and dV/dK stores switched off (MMT_ATT_BWD_DEBUG=7).
namespace bwd2 {
constexpr int KT = 128, QT = 64;
// NSW softmax warps (8 or 16): PARTS = NSW / 4 warps share a TMEM lane quarter and split a tile's 64 query columns
constexpr uint32_t KTILE = KT * DH * 2;              // 32 KB: K or V, 2 sub-tiles [128 keys x 128 B]
constexpr uint32_t KSUB = KT * 128;                  // 16 KB
constexpr uint32_t QTILE = QT * DH * 2;              // 16 KB: Q or dO tile, 2 sub-tiles [64 queries x 128 B]
constexpr uint32_t QSUB = QT * 128;                  // 8 KB
constexpr uint32_t PTILE = KT * QT * 2;              // 16 KB: Pd^T or dS^T [128 keys x 64 queries]
constexpr int QST = 2;                               // Q / dO stages
constexpr uint32_t STG = QT * DH * 4;                // 32 KB: fp32 dQ staging, 4 boxes [64 queries x 32 dh]
constexpr uint32_t OFF_K = 0, OFF_V = KTILE, OFF_Q = 2 * KTILE, OFF_DO = OFF_Q + QST * QTILE, OFF_P = OFF_DO + QST * QTILE,
                   OFF_DS = OFF_P + 2 * PTILE, OFF_STG = OFF_DS + 2 * PTILE, OFF_BAR = OFF_STG + STG;
// no alignment slack: the __align__(1024) extern declaration aligns the dynamic window (it shows up as 1 KB of static
// shared memory, which counts against the 227 KB limit); the kernel traps if that ever fails to hold
constexpr size_t SMEM = OFF_BAR + 128 /*barriers*/ + 4 * QT * 4 /*lse, delta x 2 stages*/;
static_assert(SMEM + 1024 <= 232448, "attention16 bwd2: shared memory");
constexpr uint32_t TM_DV = 0, TM_DK = 128, TM_ST = 256, TM_DP = 384, TM_DQ = 448;

template <int NSW>
__global__ void __launch_bounds__(64 + NSW * 32, 1) attention16_bwd2_kernel(const __grid_constant__ CUtensorMap map_kv,
                                                                      const __grid_constant__ CUtensorMap map_q,
                                                                      const __grid_constant__ CUtensorMap map_do,
                                                                      const __grid_constant__ CUtensorMap map_out,
                                                                      const __grid_constant__ CUtensorMap map_dq,
                                                                      const AttBwdArgs args) {
  constexpr int SMT = NSW * 32;                       // softmax threads
  constexpr int PARTS = NSW / 4;
  constexpr int CW = QT / PARTS;                      // query columns of a tile per thread (32 or 16)
  constexpr int EW = DH / PARTS;                      // dh columns of dV / dK per thread (64 or 32)
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;             // [3]
  uint64_t* st_full = bars + 4;              // [2] S^T stage written
  uint64_t* dp_full = bars + 6;              // dPd^T written
  uint64_t* pds_full = bars + 7;             // [2] Pd^T / dS^T tiles written, S^T stage and dPd^T consumed (SMT arrivals)
  uint64_t* mma2_done = bars + 9;            // [2] dV / dK / dQ^T MMAs of the tile complete (its Pd^T / dS^T stage is free)
  uint64_t* dq_drained = bars + 11;          // dQ^T read out of TMEM (SMT arrivals)
  uint64_t* qdo_free = bars + 12;            // [3] the tile's MMAs no longer read its Q / dO stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  float* s_lse2 = reinterpret_cast<float*>(bars + 16);       // [2][QT] log2-domain lse (+inf: no such query)
  float* s_delta = s_lse2 + 2 * QT;                          // [2][QT]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, S = args.S;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * KT;
  const int nq = (S + QT - 1) / QT;
  const int d_model = H * DH;
  const bool bf16 = args.bf16 != 0;
  const uint32_t smem_u = smem_u32(smem);

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(&st_full[t], 1); mbar_init(&pds_full[t], SMT); mbar_init(&mma2_done[t], 1); }
    for (int t = 0; t < QST; ++t) { mbar_init(&qdo_full[t], 1); mbar_init(&qdo_free[t], 1); }
    mbar_init(dp_full, 1); mbar_init(dq_drained, SMT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int row_k = b * S + k0;
      mbar_arrive_expect_tx(kv_full, 2 * KTILE);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        tma_load_2d(smem + OFF_K + t * KSUB, &map_kv, kv_full, d_model + h * DH + 64 * t, row_k);
        tma_load_2d(smem + OFF_V + t * KSUB, &map_kv, kv_full, 2 * d_model + h * DH + 64 * t, row_k);
      }
      for (int i = 0; i < nq; ++i) {
        const int s = i % QST;
        if (i >= QST) mbar_wait(&qdo_free[s], ((i / QST) - 1) & 1);  // tile i-QST (same stage) no longer read
        const int row_q = b * S + i * QT;
        mbar_arrive_expect_tx(&qdo_full[s], 2 * QTILE);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          tma_load_2d(smem + OFF_Q + s * QTILE + t * QSUB, &map_q, &qdo_full[s], h * DH + 64 * t, row_q);
          tma_load_2d(smem + OFF_DO + s * QTILE + t * QSUB, &map_do, &qdo_full[s], h * DH + 64 * t, row_q);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_s = idesc16(128, QT, false, false, bf16);    // S^T, dPd^T: A (K / V) and B (Q / dO) K-major
      const uint32_t id_v = idesc16(128, DH, false, true, bf16);     // dV, dK: A (Pd^T / dS^T) K-major over queries, B MN-major
      const uint32_t id_q = idesc16(128, QT, true, true, bf16);      // dQ^T: A (K) and B (dS^T) MN-major over keys
      const uint32_t sk = smem_u + OFF_K, sv = smem_u + OFF_V;
      auto mma_st = [&](int j) {                                     // S^T(j) = K Q_j^T   (k = dh)
        const uint32_t sq = smem_u + OFF_Q + (j % QST) * QTILE, acc = tmem + TM_ST + (uint32_t)(j & 1) * QT;
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks)
          umma_f16_ss(acc, make_smem_desc(sk + (ks >> 2) * KSUB + (ks & 3) * 32, 16, 1024, 2),
                      make_smem_desc(sq + (ks >> 2) * QSUB + (ks & 3) * 32, 16, 1024, 2), id_s, ks > 0 ? 1u : 0u);
      };
      auto mma_dp = [&](int j) {                                     // dPd^T(j) = V dO_j^T
        const uint32_t sdo = smem_u + OFF_DO + (j % QST) * QTILE;
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks)
          umma_f16_ss(tmem + TM_DP, make_smem_desc(sv + (ks >> 2) * KSUB + (ks & 3) * 32, 16, 1024, 2),
                      make_smem_desc(sdo + (ks >> 2) * QSUB + (ks & 3) * 32, 16, 1024, 2), id_s, ks > 0 ? 1u : 0u);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      tc_fence_after();
      mma_st(0);
      umma_commit(&st_full[0]);
      mma_dp(0);
      umma_commit(dp_full);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        // next S^T while the softmax warps work on tile i.  (Polling qdo_full and pds_full alternately, so that this tile's
        // MMAs never queue behind a late load, was measured no faster: 94.7 vs 94.4 us per launch.)
        if (i + 1 < nq) {
          mbar_wait(&qdo_full[(i + 1) % QST], ((i + 1) / QST) & 1);
          tc_fence_after();
          mma_st(i + 1);
          umma_commit(&st_full[s ^ 1]);
        }
        mbar_wait(&pds_full[s], (i >> 1) & 1);
        tc_fence_after();
        if (i + 1 < nq) {
          mma_dp(i + 1);
          umma_commit(dp_full);
        }
        const uint32_t sq = smem_u + OFF_Q + (i % QST) * QTILE, sdo = smem_u + OFF_DO + (i % QST) * QTILE;
        const uint32_t sp = smem_u + OFF_P + s * PTILE, sds = smem_u + OFF_DS + s * PTILE;
#pragma unroll
        for (int ks = 0; ks < QT / 16; ++ks)                         // dV += Pd^T dO_i   (k = queries)
          umma_f16_ss(tmem + TM_DV, make_smem_desc(sp + ks * 32, 16, 1024, 2), make_smem_desc(sdo + ks * 2048, QSUB, 1024, 2),
                      id_v, (i > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < QT / 16; ++ks)                         // dK += dS^T Q_i
          umma_f16_ss(tmem + TM_DK, make_smem_desc(sds + ks * 32, 16, 1024, 2), make_smem_desc(sq + ks * 2048, QSUB, 1024, 2),
                      id_v, (i > 0 || ks > 0) ? 1u : 0u);
        if (i >= 1) {
          mbar_wait(dq_drained, (i - 1) & 1);                        // dQ^T(i-1) has left its TMEM columns
          tc_fence_after();
        }
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks)                         // dQ_i^T = K^T dS   (k = keys)
          umma_f16_ss(tmem + TM_DQ, make_smem_desc(sk + ks * 2048, KSUB, 1024, 2), make_smem_desc(sds + ks * 2048, PTILE, 1024, 2),
                      id_q, ks > 0 ? 1u : 0u);
        umma_commit(&mma2_done[s]);
        umma_commit(&qdo_free[i % QST]);
      }
    }
  } else {
    // ===================== softmax backward, dQ drain, epilogue (warps 2..9) =====================
    const int q4 = warp & 3;                                   // TMEM lane quarter
    const int part = (warp - 2) >> 2;                          // which CW of a tile's 64 query columns
    const int r = q4 * 32 + lane;                              // key row (S^T, dV, dK) or dh row (dQ^T) == TMEM lane
    const int key = k0 + r;
    const bool key_ok = key < S;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const float mterm = key_ok ? (1.0f - __ldg(args.mask + (int64_t)b * S + key)) * (-10000.0f * LOG2E) : -INFINITY;
    const uint64_t seed = args.seed + (args.ctr ? *args.ctr : 0);
    const uint32_t key32 = drop_key(seed, args.site);
    const uint32_t thr = (uint32_t)(args.p_drop * 65536.0f);
    const bool drop = args.p_drop > 0.f;
    const uint32_t half_pitch = (uint32_t)((S + 1) >> 1);
    const uint32_t kpair = (uint32_t)key >> 1;
    const uint32_t kshift = (key & 1) ? 16u : 0u;
    const uint32_t prow0 = (uint32_t)(((int64_t)b * H + h) * S);
    const int cbase = part * CW;
    const uint32_t sw = (uint32_t)(r & 7);
    const int tq = (int)threadIdx.x - 64;                      // 0 .. SMT-1: the first QT threads fetch the per-query vectors
    const float* lse_g = args.lse + ((int64_t)b * H + h) * S;
    const float* delta_g = args.delta + ((int64_t)b * H + h) * S;
    if (tq < QT) {
      const bool ok = tq < S;
      s_lse2[tq] = ok ? __ldg(lse_g + tq) * LOG2E : INFINITY;
      s_delta[tq] = ok ? __ldg(delta_g + tq) : 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");

    // dQ^T(j) [dh lanes x 64 query columns] -> fp32 staging [query rows x dh], four 128-byte-swizzled boxes of 32 dh
    // columns (one per lane quarter) -> TMA reduce-add into dq32 (clipped at S).  Ends with a barrier of the SMT threads.
    const uint32_t stg_u = smem_u + OFF_STG;
    auto drain_dq = [&](int j) {
      mbar_wait(&mma2_done[j & 1], (j >> 1) & 1);
      tc_fence_after();
      float v[CW];
      tmem_ldn(tmem + TM_DQ + lane_addr + (uint32_t)cbase, v);
      tc_fence_before();
      mbar_arrive(dq_drained);
      if (threadIdx.x == 64) bulk_wait_read0();                // the previous reduction has read the staging tile
      asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");
      const uint32_t base = stg_u + (uint32_t)q4 * (QT * 128) + (uint32_t)(lane & 3) * 4;
      const uint32_t ch = (uint32_t)(lane >> 2);
#pragma unroll
      for (int t = 0; t < CW; ++t) {
        const uint32_t q = (uint32_t)(cbase + t);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + q * 128 + ((ch ^ (q & 7)) << 4)), "f"(v[t]) : "memory");
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");
      if (threadIdx.x == 64 && !(args.debug & 1)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_reduce_add_3d(&map_dq, stg_u + c * (QT * 128), h * DH + 32 * c, j * QT, b);
        bulk_commit();
      }
    };

    for (int i = 0; i < nq; ++i) {
      const int s = i & 1;
      const int qbase = i * QT;
      float nl = INFINITY, nd = 0.f;                           // next tile's lse / delta: in flight during this tile
      if (tq < QT && i + 1 < nq) {
        const int qq = qbase + QT + tq;
        if (qq < S) { nl = __ldg(lse_g + qq) * LOG2E; nd = __ldg(delta_g + qq); }
      }
      const float* lse2 = s_lse2 + s * QT + cbase;
      const float* dlt = s_delta + s * QT + cbase;
      const uint32_t prow = smem_u + OFF_P + s * PTILE + (uint32_t)r * 128;
      const uint32_t dsrow = smem_u + OFF_DS + s * PTILE + (uint32_t)r * 128;
      const bool live = (qbase + cbase < S) && (k0 + q4 * 32 < S) && !(args.debug & 2);   // warp-uniform
      mbar_wait(&st_full[s], (i >> 1) & 1);
      tc_fence_after();
      // (Two leaner formulations were measured and lost: sharing each dropout hash between the two lanes of a key pair by
      // shuffle, and carrying the keep decision in p's sign bit -- 94.4 vs 90.1 us per launch.)
      float p[CW];
      uint32_t keepmask = 0xffffffffu;
      if (live) {
        float sv_[CW];
        tmem_ldn(tmem + TM_ST + (uint32_t)(s * QT) + lane_addr + (uint32_t)cbase, sv_);
        uint32_t pk[CW / 2];
#pragma unroll
        for (int t = 0; t < CW; t += 2) {
          float pv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            p[t + u] = fast_ex2(fmaf(sv_[t + u], args.scale_log2, mterm) - lse2[t + u]);
            float keep = args.inv_keep;
            if (drop) {
              const uint32_t w = drop_word(key32, prow0 + (uint32_t)(qbase + cbase + t + u), half_pitch, kpair);
              if (((w >> kshift) & 0xffffu) < thr) { keep = 0.f; keepmask &= ~(1u << (t + u)); }
            }
            pv[u] = p[t + u] * keep;
          }
          pk[t >> 1] = pack2(pv[0], pv[1], bf16);
        }
#pragma unroll
        for (int c = 0; c < CW / 8; ++c)
          sts128u(prow + ((((uint32_t)(part * (CW / 8) + c)) ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < CW / 8; ++c) sts128u(prow + ((((uint32_t)(part * (CW / 8) + c)) ^ sw) << 4), 0u, 0u, 0u, 0u);
      }
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
      if (live) {
        float dp[CW];
        tmem_ldn(tmem + TM_DP + lane_addr + (uint32_t)cbase, dp);
        uint32_t dk[CW / 2];
#pragma unroll
        for (int t = 0; t < CW; t += 2) {
          const float k0_ = ((keepmask >> t) & 1u) ? args.inv_keep : 0.f, k1_ = ((keepmask >> (t + 1)) & 1u) ? args.inv_keep : 0.f;
          const float d0 = p[t] * fmaf(dp[t], k0_, -dlt[t]) * args.scale;
          const float d1 = p[t + 1] * fmaf(dp[t + 1], k1_, -dlt[t + 1]) * args.scale;
          dk[t >> 1] = pack2(d0, d1, bf16);
        }
#pragma unroll
        for (int c = 0; c < CW / 8; ++c)
          sts128u(dsrow + ((((uint32_t)(part * (CW / 8) + c)) ^ sw) << 4), dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < CW / 8; ++c) sts128u(dsrow + ((((uint32_t)(part * (CW / 8) + c)) ^ sw) << 4), 0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();                                 // generic-proxy smem writes -> visible to the MMAs
      tc_fence_before();
      mbar_arrive(&pds_full[s]);
      if (tq < QT && i + 1 < nq) { s_lse2[(s ^ 1) * QT + tq] = nl; s_delta[(s ^ 1) * QT + tq] = nd; }
      if (i >= 1) drain_dq(i - 1);                              // (its barriers also publish the vectors above)
      else asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");
    }
    drain_dq(nq - 1);

    // epilogue: dV_j, dK_j (lane = key row, EW dh columns per thread) -> 128-byte-swizzled [128 keys x 64] tiles in the
    // Q / dO buffers -> TMA stores into the V and K blocks of dqkv16 (clipped at S); bias-gradient column sums
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t tm = which == 0 ? TM_DV : TM_DK;
      const int blk = which == 0 ? 2 : 1;                       // column block of dqkv16: Q | K | V
      const uint32_t tile_u = smem_u + (which == 0 ? OFF_Q : OFF_DO);
      float* bsum = args.dbias + blk * d_model + h * DH + part * EW;
#pragma unroll 1
      for (int c = 0; c < EW / 16; ++c) {
        float v[16];
        const int col = part * EW + c * 16;
        tmem_ld16f(tmem + tm + lane_addr + (uint32_t)col, v);
        if (!(args.debug & 4)) {
          const uint32_t rb = tile_u + (uint32_t)(col >> 6) * KSUB + (uint32_t)r * 128;
          const uint32_t ch0 = (uint32_t)((col & 63) >> 3);
          sts128u(rb + ((ch0 ^ sw) << 4), pack2(v[0], v[1], bf16), pack2(v[2], v[3], bf16), pack2(v[4], v[5], bf16),
                  pack2(v[6], v[7], bf16));
          sts128u(rb + (((ch0 + 1) ^ sw) << 4), pack2(v[8], v[9], bf16), pack2(v[10], v[11], bf16),
                  pack2(v[12], v[13], bf16), pack2(v[14], v[15], bf16));
        }
        if (args.debug & 8) continue;
        if (!key_ok) {
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = 0.f;
        }
        bwd::warp_colsum16(v, lane, bsum + c * 16, args.inv_scale16);
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(SMT) : "memory");
    if (threadIdx.x == 64) {
      if (!(args.debug & 4)) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          tma_store_3d(&map_out, smem_u + OFF_Q + t * KSUB, 2 * d_model + h * DH + 64 * t, k0, b);
          tma_store_3d(&map_out, smem_u + OFF_DO + t * KSUB, d_model + h * DH + 64 * t, k0, b);
        }
        bulk_commit();
      }
      bulk_wait_read0();                                        // smem may go once every bulk store / reduction has read it
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
}  // namespace bwd2

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D 16-bit tensor map [rows, cols] (row pitch ld elements), box {64 cols, box_rows}, 128-byte swizzle
int make_map16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int dtype,
                  const char* what) {
  static EncodeTiledFn enc = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return (EncodeTiledFn) nullptr;
  }();
  MMT_ARG_CHECK(enc != nullptr, MMT_E_UNSUPPORTED, "cuTensorMapEncodeTiled unavailable");
  MMT_ARG_CHECK(((uintptr_t)base % 16) == 0 && (ld * 2) % 16 == 0, MMT_E_ALIGN,
                "tensor map %s needs a 16-byte aligned base and row pitch", what);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, estr[2] = {1, 1};
  CUresult r = enc(map, dtype == MMT_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMT_ARG_CHECK(r == CUDA_SUCCESS, MMT_E_UNSUPPORTED, "cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return 0;
}

// [B, S, cols] view of a [B*S, ld] tensor (16-bit, or fp32 when dtype < 0) for TMA STORES / REDUCTIONS: the box (128 bytes
// of columns x box_rows tokens x 1) is clipped at S, so a ragged last tile does not spill into the next batch item's rows
int make_map_3d(CUtensorMap* map, const void* base, int64_t B, int64_t S, int64_t cols, int64_t ld, int box_rows, int dtype,
                const char* what) {
  static EncodeTiledFn enc = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return (EncodeTiledFn) nullptr;
  }();
  MMT_ARG_CHECK(enc != nullptr, MMT_E_UNSUPPORTED, "cuTensorMapEncodeTiled unavailable");
  const int es = dtype < 0 ? 4 : 2;
  MMT_ARG_CHECK(((uintptr_t)base % 16) == 0 && (ld * es) % 16 == 0, MMT_E_ALIGN,
                "tensor map %s needs a 16-byte aligned base and row pitch", what);
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * es, (cuuint64_t)S * ld * es};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows, 1}, estr[3] = {1, 1, 1};
  const CUtensorMapDataType cdt = dtype < 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                  : dtype == MMT_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, cdt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMT_ARG_CHECK(r == CUDA_SUCCESS, MMT_E_UNSUPPORTED, "cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return 0;
}

}  // namespace
}  // namespace mmt

using namespace mmt;

extern "C" int mmt_attention16_fwd(const void* qkv16, const float* mask, int32_t B, int32_t H, int32_t S, int32_t dh,
                                   float scale, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site,
                                   void* ctx16, float* lse, int32_t dtype, void* stream) {
  MMT_ARG_CHECK(qkv16 && mask && ctx16, MMT_E_ARG, "mmt_attention16_fwd: null pointer");
  MMT_ARG_CHECK(dh == DH, MMT_E_SHAPE, "mmt_attention16_fwd: head dim %d unsupported (only %d)", dh, DH);
  MMT_ARG_CHECK(B > 0 && H > 0 && S > 0 && B <= 65535 && H <= 65535, MMT_E_SHAPE, "mmt_attention16_fwd: bad shape B=%d H=%d S=%d", B, H, S);
  MMT_ARG_CHECK((int64_t)B * H * S * ((S + 1) / 2) < (1ll << 32), MMT_E_SHAPE, "mmt_attention16_fwd: dropout counter overflow");
  MMT_ARG_CHECK(p_drop >= 0.f && p_drop < 1.f, MMT_E_ARG, "mmt_attention16_fwd: p_drop=%f", (double)p_drop);
  MMT_ARG_CHECK(dtype == MMT_DT_F16 || dtype == MMT_DT_BF16, MMT_E_ARG, "mmt_attention16_fwd: bad dtype %d", dtype);
  const int64_t rows = (int64_t)B * S, cols = 3LL * H * DH;
  CUtensorMap mq, mk;
  int rc = make_map16_2d(&mq, qkv16, rows, cols, cols, fwd::QM, dtype, "Q");
  if (rc) return rc;
  rc = make_map16_2d(&mk, qkv16, rows, cols, cols, fwd::KB, dtype, "K/V");
  if (rc) return rc;
  CUtensorMap mo;
  rc = make_map_3d(&mo, ctx16, B, S, (int64_t)H * DH, (int64_t)H * DH, fwd::QM, dtype, "context");
  if (rc) return rc;
  AttArgs a;
  a.mask = mask; a.ctx16 = reinterpret_cast<uint16_t*>(ctx16); a.lse = lse;
  a.B = B; a.H = H; a.S = S;
  a.scale_log2 = scale * LOG2E;
  a.p_drop = p_drop; a.inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  a.seed = seed; a.ctr = seed_ctr; a.site = site; a.bf16 = dtype == MMT_DT_BF16 ? 1 : 0;
  rc = ensure_dynamic_smem((const void*)fwd::attention16_fwd_kernel, fwd::SMEM, "attention16_fwd smem attribute");
  if (rc) return rc;
  dim3 grid((S + fwd::QM - 1) / fwd::QM, H, B);
  launch_pdl(fwd::attention16_fwd_kernel, grid, dim3(fwd::THREADS), fwd::SMEM, (cudaStream_t)stream, mq, mk, mo, a);
  MMT_LAUNCH_CHECK("attention16_fwd_kernel");
  return 0;
}

extern "C" int mmt_attention16_bwd(const void* qkv16, const void* ctx16, const void* dctx16, const float* lse,
                                   const float* mask, int32_t B, int32_t H, int32_t S, int32_t dh, float scale,
                                   float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float scale16,
                                   void* dqkv16, float* dq32, float* delta, float* dbias, int32_t dtype, void* stream) {
  MMT_ARG_CHECK(qkv16 && ctx16 && dctx16 && lse && mask && dqkv16 && dq32 && delta && dbias, MMT_E_ARG,
                "mmt_attention16_bwd: null pointer");
  MMT_ARG_CHECK(dh == DH, MMT_E_SHAPE, "mmt_attention16_bwd: head dim %d unsupported (only %d)", dh, DH);
  MMT_ARG_CHECK(B > 0 && H > 0 && S > 0 && B <= 65535 && H <= 65535, MMT_E_SHAPE, "mmt_attention16_bwd: bad shape B=%d H=%d S=%d", B, H, S);
  MMT_ARG_CHECK((int64_t)B * H * S * ((S + 1) / 2) < (1ll << 32), MMT_E_SHAPE, "mmt_attention16_bwd: dropout counter overflow");
  MMT_ARG_CHECK(p_drop >= 0.f && p_drop < 1.f && scale16 > 0.f, MMT_E_ARG, "mmt_attention16_bwd: p_drop=%f scale16=%f", (double)p_drop, (double)scale16);
  MMT_ARG_CHECK(dtype == MMT_DT_F16 || dtype == MMT_DT_BF16, MMT_E_ARG, "mmt_attention16_bwd: bad dtype %d", dtype);
  const int d_model = H * DH;
  CHECK_D(d_model);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * S;
  const int bf16 = dtype == MMT_DT_BF16 ? 1 : 0;
  {
    int64_t blocks = (rows + 7) / 8;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    launch_pdl(bwd::attn_delta16_kernel, dim3((int)blocks), dim3(256), 0, st, reinterpret_cast<const uint16_t*>(dctx16),
               reinterpret_cast<const uint16_t*>(ctx16), rows, S, H, delta, bf16);
    MMT_LAUNCH_CHECK("attn_delta16_kernel");
  }
  CUtensorMap mqkv, mdo, mout, mdq;
  int rc = make_map_3d(&mout, dqkv16, B, S, 3LL * d_model, 3LL * d_model, 128, dtype, "dQKV");
  if (rc) return rc;
  rc = make_map_3d(&mdq, dq32, B, S, d_model, d_model, 128, -1, "dQ fp32");
  if (rc) return rc;
  rc = make_map16_2d(&mqkv, qkv16, rows, 3LL * d_model, 3LL * d_model, 128, dtype, "QKV");
  if (rc) return rc;
  rc = make_map16_2d(&mdo, dctx16, rows, d_model, d_model, 128, dtype, "dO");
  if (rc) return rc;
  AttBwdArgs a;
  a.mask = mask; a.lse = lse; a.delta = delta;
  a.dqkv16 = reinterpret_cast<uint16_t*>(dqkv16); a.dq32 = dq32; a.dbias = dbias;
  a.B = B; a.H = H; a.S = S;
  a.scale = scale; a.scale_log2 = scale * LOG2E;
  a.p_drop = p_drop; a.inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  a.inv_scale16 = 1.0f / scale16;
  a.seed = seed; a.ctr = seed_ctr; a.site = site; a.bf16 = bf16;
  static const int dbg = [] { const char* e = getenv("MMT_ATT_BWD_DEBUG"); return e ? atoi(e) : 0; }();
  a.debug = dbg;
  static const int nsw = [] {                          // softmax warps per CTA (A/B switch; 8 is the measured default)
    const char* e = getenv("MMT_ATT_BWD_WARPS");
    return e && atoi(e) == 16 ? 16 : 8;
  }();
  static const int dq_force = [] { const char* e = getenv("MMT_ATT_BWD_DQ"); return e ? atoi(e) : -1; }();
  static const int v1 = [] { const char* e = getenv("MMT_ATT_BWD_V1"); return e ? atoi(e) : 0; }();   // 1: the serial kernel
  const int nkt = (S + bwd::KT - 1) / bwd::KT;
  if (!v1 && dq_force < 0) {
    // pipelined kernel: 64-query tiles, dQ always through the fp32 reduction + finish kernel
    CUtensorMap mq64, mdo64, mdq64;
    rc = make_map16_2d(&mq64, qkv16, rows, 3LL * d_model, 3LL * d_model, bwd2::QT, dtype, "Q");
    if (rc) return rc;
    rc = make_map16_2d(&mdo64, dctx16, rows, d_model, d_model, bwd2::QT, dtype, "dO");
    if (rc) return rc;
    rc = make_map_3d(&mdq64, dq32, B, S, d_model, d_model, bwd2::QT, -1, "dQ fp32");
    if (rc) return rc;
    a.dq_mode = 0;
    if (nsw == 16) {
      rc = ensure_dynamic_smem((const void*)bwd2::attention16_bwd2_kernel<16>, bwd2::SMEM, "attention16_bwd2 smem attribute");
      if (rc) return rc;
      launch_pdl(bwd2::attention16_bwd2_kernel<16>, dim3(nkt, H, B), dim3(64 + 16 * 32), bwd2::SMEM, st, mqkv, mq64, mdo64, mout,
                 mdq64, a);
    } else {
      rc = ensure_dynamic_smem((const void*)bwd2::attention16_bwd2_kernel<8>, bwd2::SMEM, "attention16_bwd2 smem attribute");
      if (rc) return rc;
      launch_pdl(bwd2::attention16_bwd2_kernel<8>, dim3(nkt, H, B), dim3(64 + 8 * 32), bwd2::SMEM, st, mqkv, mq64, mdo64, mout,
                 mdq64, a);
    }
    MMT_LAUNCH_CHECK("attention16_bwd2_kernel");
    const int d4 = d_model / 4;
    const int bx = d4 <= 128 ? d4 : 128;
    const int by = 256 / bx > 0 ? 256 / bx : 1;
    const int64_t blocks = (rows + bwd::FIN_ROWS - 1) / bwd::FIN_ROWS;
    launch_pdl(bwd::attn_dq_finish_kernel, dim3((int)blocks), dim3(bx, by), 0, st, reinterpret_cast<float4*>(dq32), rows, d4,
               reinterpret_cast<uint16_t*>(dqkv16), dbias, 1.0f / scale16, bf16);
    MMT_LAUNCH_CHECK("attn_dq_finish_kernel");
    return 0;
  }
  // see the kernel.  The 2-CTA hand-over (MMT_ATT_BWD_DQ=2) is correct but measured SLOWER than the reduction + finish kernel
  // (134 vs 114 us per layer at B=64, S=218: its per-lane fp32 loads / stores sit on each CTA's critical path), so it is opt-in.
  a.dq_mode = nkt == 1 ? 1 : 0;
  if (dq_force == 2 && nkt == 2) a.dq_mode = 2;
  if (dq_force == 0 || dq_force == 3) a.dq_mode = 0;
  const int cluster_x = (a.dq_mode == 2 || (dq_force == 3 && nkt == 2)) ? 2 : 1;   // 3: timing experiment (atomics, but clustered)
  dim3 grid(nkt, H, B);
  if (nsw == 16) {
    rc = ensure_dynamic_smem((const void*)bwd::attention16_bwd_kernel<16>, bwd::SMEM, "attention16_bwd smem attribute");
    if (rc) return rc;
    launch_pdl_cluster(bwd::attention16_bwd_kernel<16>, grid, dim3(64 + 16 * 32), bwd::SMEM, st, cluster_x, mqkv, mdo, mout, mdq, a);
  } else {
    rc = ensure_dynamic_smem((const void*)bwd::attention16_bwd_kernel<8>, bwd::SMEM, "attention16_bwd smem attribute");
    if (rc) return rc;
    launch_pdl_cluster(bwd::attention16_bwd_kernel<8>, grid, dim3(64 + 8 * 32), bwd::SMEM, st, cluster_x, mqkv, mdo, mout, mdq, a);
  }
  MMT_LAUNCH_CHECK("attention16_bwd_kernel");
  if (a.dq_mode == 0) {
    const int d4 = d_model / 4;
    const int bx = d4 <= 128 ? d4 : 128;
    const int by = 256 / bx > 0 ? 256 / bx : 1;
    const int64_t blocks = (rows + bwd::FIN_ROWS - 1) / bwd::FIN_ROWS;
    launch_pdl(bwd::attn_dq_finish_kernel, dim3((int)blocks), dim3(bx, by), 0, st, reinterpret_cast<float4*>(dq32), rows, d4,
               reinterpret_cast<uint16_t*>(dqkv16), dbias, 1.0f / scale16, bf16);
    MMT_LAUNCH_CHECK("attn_dq_finish_kernel");
  }
  return 0;
}
