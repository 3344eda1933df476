// Fused flash-style self-attention forward on tcgen05 (model/bert.py:136-172, one kernel):
//     ctx[b, i, h*dh:(h+1)*dh] = dropout(softmax(Q K^T / sqrt(dh) + (1 - mask) * -10000)) V
// The S x S score / probability matrices are NEVER written to HBM: S = Q K^T is accumulated in
// TMEM, the softmax warps turn it into P in place (tcgen05.ld -> registers -> tcgen05.st), and
// P is consumed as the A operand of the second tensor-core product straight from TMEM.
//
// One CTA per (b, h, 128-query tile); keys are processed in blocks of KB = 224 (S = 218 fits one
// block; longer sequences take several with an online-softmax rescale of the TMEM accumulator).
//   warp 0    TMA producer: Q tile [128 x dh] (K-major, 4 swizzle-128B sub-tiles), then per key block
//             K [KB x dh] (K-major) and -- once the score MMAs have drained K -- V [KB x dh] into
//             the SAME buffer as an MN-major operand (swizzle-128B/32B-atom boxes)
//   warp 1    MMA issuer: S = Q K^T (kind::tf32, M=128, N=KB, 16 k-steps, operands in smem), then
//             O += P V (A = P from TMEM, B = V from smem, KB/8 k-steps)
//   warps 2-9 softmax: TWO threads per query row (each TMEM lane quarter is served by two warps that
//             take half of the key columns each; row max / sum are exchanged through shared memory):
//             scale + additive mask, running max / sum, exp2, Philox dropout (16 random bits per
//             element, identical keying to the unfused path), write P, rescale O if the running max
//             moved; finally O / l -> ctx and log-sum-exp for the backward pass.
//             The softmax is the long pole of a tile (~25 ALU instructions per score with dropout
//             against 2 x 128 MMA flops), hence the second warp set and the cheaper random bits.
// Shared memory: Q 64 KB + K|V 112 KB = 176 KB; TMEM: S/P 224 columns + O 128 columns.
// Q and K arrive as four 32-column slices of dh on separate barriers, so the score MMAs start after
// the first quarter of the operand bytes.
#include "tc_ptx.cuh"

namespace mmt {
namespace {
using namespace tc;

constexpr int DH = 128;                 // head dim (4 sub-tiles of 32 fp32 = 128 B rows)
constexpr int QM = 128;                 // query rows per CTA
constexpr int KB = 224;                 // keys per block (UMMA N, multiple of 16, <= 256)
constexpr int ATT_THREADS = 320;                // TMA warp, MMA warp, 8 softmax warps
constexpr int SM_THREADS = 256;
constexpr uint32_t Q_BYTES = QM * DH * 4;          // 64 KB
constexpr uint32_t KV_BYTES = KB * DH * 4;         // 112 KB
constexpr uint32_t TM_S = 0, TM_O = 256;           // TMEM column offsets (512 allocated)

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
        "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

struct AttArgs {
  const float* mask;      // [B, S] 1 = attend
  float* ctx;             // [B*S, H*DH]
  float* lse;             // [B, H, S] natural-log sum-exp of the masked, scaled scores
  int B, H, S;
  float scale_log2;       // (1/sqrt(dh)) * tf32 compensation * log2(e)
  float mask_log2;        // -10000 * log2(e)
  float out_scale;        // tf32 compensation of the P V product
  float p_drop, inv_keep;
  uint64_t seed;
  uint32_t site;
  const uint64_t* ctr;
};

// map_qk: qkv viewed as [B*S rows, 3*H*DH cols], box {32 cols, rows<=256}, SWIZZLE_128B
// map_v : same tensor, box {32 cols, 32 rows}, SWIZZLE_128B_ATOM_32B (MN-major B operand)
// WRITE_P (training, S <= KB): the normalised probabilities P and their dropped copy Pd are ALSO
// streamed to HBM (TMA stores from a per-warp staging tile) for the backward pass, which then needs
// neither the Q K^T product nor the softmax again; the row sum is taken in an extra exp pass so that
// what is written (and what feeds the P V product) is already normalised.
template <bool WRITE_P>
__global__ void __launch_bounds__(ATT_THREADS, 1) attention_fwd_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                       const __grid_constant__ CUtensorMap map_k,
                                                                       const __grid_constant__ CUtensorMap map_v,
                                                                       const __grid_constant__ CUtensorMap map_p,
                                                                       const __grid_constant__ CUtensorMap map_pd,
                                                                       const AttArgs args) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sq = smem;                       // 4 sub-tiles [128 rows x 128 B]
  uint8_t* skv = smem + Q_BYTES;            // K: 4 sub-tiles [224 rows x 128 B];  V: 4 n-chunks [224 k-rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_BYTES + KV_BYTES);
  uint64_t* qk_full = bars + 0;             // [4] Q and K sub-tile t (32 of the dh columns) have landed
  uint64_t* v_full = bars + 4;
  uint64_t* k_free = bars + 5;              // score MMAs have finished reading K
  uint64_t* v_free = bars + 6;              // P V MMAs have finished reading V
  uint64_t* s_full = bars + 7;              // scores ready in TMEM
  uint64_t* p_full = bars + 8;              // probabilities written (and O rescaled) -- 256 arrivals
  uint64_t* o_full = bars + 9;              // P V accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* smask = reinterpret_cast<float*>(bars + 16);      // additive mask of the current key block (log2 domain)
  float* xch = smask + KB;                                 // [2][2][128] row max / row sum exchange between column halves
  float* pstage = xch + 4 * QM;                            // WRITE_P: per softmax warp, P and Pd tiles [32 rows][16 cols]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, S = args.S;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * QM;
  const int nblk = (S + KB - 1) / KB;
  const int d_model = H * DH;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 10; ++i) mbar_init(&bars[i], i == 8 ? SM_THREADS : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                           // prologue done: now wait for the producer grid

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int row_q = b * S + q0;
      for (int j = 0; j < nblk; ++j) {
        const int row_k = b * S + j * KB;
        mbar_wait(v_free, (j & 1) ^ 1);                       // previous block's V no longer read
#pragma unroll
        for (int t = 0; t < 4; ++t) {                          // per 32-column slice of dh: the MMAs start on slice 0
          mbar_arrive_expect_tx(&qk_full[t], (j == 0 ? Q_BYTES / 4 : 0) + KV_BYTES / 4);
          if (j == 0) tma_load_2d(sq + t * (QM * 128), &map_q, &qk_full[t], h * DH + 32 * t, row_q);
          tma_load_2d(skv + t * (KB * 128), &map_k, &qk_full[t], d_model + h * DH + 32 * t, row_k);
        }
        mbar_wait(k_free, j & 1);                             // scores done: K's buffer can take V
        mbar_arrive_expect_tx(v_full, KV_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c)                            // n-chunk c: dh columns 32c..32c+31
#pragma unroll
          for (int kb = 0; kb < KB / 32; ++kb)
            tma_load_2d(skv + c * (KB * 128) + kb * (32 * 128), &map_v, v_full, 2 * d_model + h * DH + 32 * c,
                        row_k + 32 * kb);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_tf32(QM, KB, false, false);     // S[128 x KB] = Q K^T
      const uint32_t idesc_o = make_idesc_tf32(QM, DH, false, true);      // O[128 x dh] += P V (V MN-major)
      for (int j = 0; j < nblk; ++j) {
        if (j > 0) mbar_wait(o_full, (j - 1) & 1);            // S/P columns are free again
#pragma unroll
        for (int ks = 0; ks < DH / 8; ++ks) {
          if ((ks & 3) == 0) { mbar_wait(&qk_full[ks >> 2], j & 1); tc_fence_after(); }
          const uint64_t da = make_smem_desc(smem_u32(sq) + (ks >> 2) * (QM * 128) + (ks & 3) * 32, 16, 1024, 2);
          const uint64_t db = make_smem_desc(smem_u32(skv) + (ks >> 2) * (KB * 128) + (ks & 3) * 32, 16, 1024, 2);
          umma_tf32(tmem + TM_S, da, db, idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(k_free);
        umma_commit(s_full);
        mbar_wait(v_full, j & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        for (int ks = 0; ks < KB / 8; ++ks) {
          // V: n-chunks KB*128 B apart (LBO), 4-row k-groups 512 B apart (SBO), 8 keys per step = 1024 B
          const uint64_t db = make_smem_desc(smem_u32(skv) + ks * 1024, KB * 128, 512, 1);
          umma_tf32_ts(tmem + TM_O, tmem + TM_S + ks * 8, db, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(v_free);
        umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / epilogue (warps 2..9) =====================
    const int q = warp & 3;                                    // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                          // which half of the key columns / of dh it handles
    const int r = q * 32 + lane;                               // query row within the tile == TMEM lane
    const int qi = q0 + r;
    const bool row_ok = qi < S;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int cbase = half * (KB / 2);                         // first key column of this thread's half
    const float* mrow = args.mask + (int64_t)b * S;
    const uint32_t prow = (uint32_t)(((int64_t)b * H + h) * S + qi);   // Philox row id (== unfused path)
    const uint64_t seed = args.seed + (args.ctr ? *args.ctr : 0);
    float m_run = -INFINITY, l_run = 0.f;                      // l_run: this half's share of the row sum
    for (int j = 0; j < nblk; ++j) {
      const int key0 = j * KB;
      // additive mask terms of this key block, shared by the softmax threads (named barrier 1)
      for (int t = threadIdx.x - 64; t < KB; t += SM_THREADS) {
        const int key = key0 + t;
        smask[t] = key < S ? (1.0f - __ldg(mrow + key)) * args.mask_log2 : -INFINITY;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: block max of the masked, scaled scores (log2 domain) over this thread's columns
      float m_part = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < KB / 32; ++c) {
        float v[16];
        tmem_ld16(tmem + TM_S + lane_addr + cbase + c * 16, v);
#pragma unroll
        for (int t = 0; t < 16; t += 4) {
          const float4 mk = *reinterpret_cast<const float4*>(smask + cbase + c * 16 + t);
          m_part = fmaxf(fmaxf(m_part, fmaf(v[t], args.scale_log2, mk.x)), fmaf(v[t + 1], args.scale_log2, mk.y));
          m_part = fmaxf(fmaxf(m_part, fmaf(v[t + 2], args.scale_log2, mk.z)), fmaf(v[t + 3], args.scale_log2, mk.w));
        }
      }
      xch[half * QM + r] = m_part;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float m_new = fmaxf(m_run, fmaxf(m_part, xch[(half ^ 1) * QM + r]));
      const float alpha = (j == 0) ? 0.f : fast_ex2(m_run - m_new);
      float norm2 = m_new;                                     // exponent offset of pass 2
      if (WRITE_P) {
        // pass 1.5 (single key block): the row sum at the final max, so that pass 2 emits normalised P
        float l_part = 0.f;
#pragma unroll 1
        for (int c = 0; c < KB / 32; ++c) {
          float v[16];
          tmem_ld16(tmem + TM_S + lane_addr + cbase + c * 16, v);
#pragma unroll
          for (int t = 0; t < 16; t += 4) {
            const float4 mk = *reinterpret_cast<const float4*>(smask + cbase + c * 16 + t);
            l_part += (fast_ex2(fmaf(v[t], args.scale_log2, mk.x) - m_new) + fast_ex2(fmaf(v[t + 1], args.scale_log2, mk.y) - m_new)) +
                      (fast_ex2(fmaf(v[t + 2], args.scale_log2, mk.z) - m_new) + fast_ex2(fmaf(v[t + 3], args.scale_log2, mk.w) - m_new));
          }
        }
        xch[(2 + half) * QM + r] = l_part;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        norm2 = m_new + log2f(l_part + xch[(2 + (half ^ 1)) * QM + r]);
      }
      float* stg_p = pstage + (warp - 2) * (2 * 32 * 16);      // this warp's P tile, Pd tile right behind it
      // pass 2: p = 2^(x - m_new), row sum (un-dropped), dropout, write P in place of S
      float l_blk = 0.f;
#pragma unroll 1
      for (int c = 0; c < KB / 32; ++c) {
        float v[16];
        float pv[WRITE_P ? 16 : 1];
        tmem_ld16(tmem + TM_S + lane_addr + cbase + c * 16, v);
#pragma unroll
        for (int t8 = 0; t8 < 16; t8 += 8) {
          float sc[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
          if (args.p_drop > 0.f)
            dropout_scale8_h16(seed, args.site, prow, (uint32_t)((key0 + cbase + c * 16 + t8) >> 3), args.p_drop,
                               args.inv_keep, sc);
          const float4 mk0 = *reinterpret_cast<const float4*>(smask + cbase + c * 16 + t8);
          const float4 mk1 = *reinterpret_cast<const float4*>(smask + cbase + c * 16 + t8 + 4);
          const float mk[8] = {mk0.x, mk0.y, mk0.z, mk0.w, mk1.x, mk1.y, mk1.z, mk1.w};
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float p = fast_ex2(fmaf(v[t8 + u], args.scale_log2, mk[u]) - norm2);   // 2^-inf = 0 past S
            l_blk += p;
            if constexpr (WRITE_P) pv[t8 + u] = p;
            v[t8 + u] = p * sc[u];
          }
        }
        tmem_st16(tmem + TM_S + lane_addr + cbase + c * 16, v);
        if constexpr (WRITE_P) {
          if (lane == 0) bulk_wait_read0();                    // the previous chunk's stores have read the staging tiles
          __syncwarp();
#pragma unroll
          for (int t = 0; t < 16; t += 4) {
            *reinterpret_cast<float4*>(stg_p + lane * 16 + t) = make_float4(pv[t], pv[t + 1], pv[t + 2], pv[t + 3]);
            if (args.p_drop > 0.f)
              *reinterpret_cast<float4*>(stg_p + 32 * 16 + lane * 16 + t) = make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int bh = b * H + h;
            tma_store_3d(&map_p, stg_p, key0 + cbase + c * 16, q0 + q * 32, bh);
            if (args.p_drop > 0.f) tma_store_3d(&map_pd, stg_p + 32 * 16, key0 + cbase + c * 16, q0 + q * 32, bh);
            bulk_commit();
          }
        }
      }
      // online-softmax rescale of the running accumulator (only when there was a previous block):
      // this thread's half of the dh columns
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);                        // previous P V has landed in O
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < DH / 32; ++c) {
          float v[16];
          tmem_ld16(tmem + TM_O + lane_addr + half * (DH / 2) + c * 16, v);
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] *= alpha;
          tmem_st16(tmem + TM_O + lane_addr + half * (DH / 2) + c * 16, v);
        }
      }
      tmem_st_wait();
      l_run = l_run * alpha + l_blk;
      m_run = WRITE_P ? norm2 : m_new;                         // WRITE_P: log2 of the full normaliser
      tc_fence_before();
      mbar_arrive(p_full);
      asm volatile("bar.sync 1, 256;" ::: "memory");          // smask / xch are rewritten by the next block
    }
    // epilogue: O / l -> ctx (this thread's half of dh), log-sum-exp for backward
    xch[half * QM + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_tot = WRITE_P ? 1.0f : l_run + xch[(half ^ 1) * QM + r];    // WRITE_P: already normalised
    mbar_wait(o_full, (nblk - 1) & 1);
    tc_fence_after();
    const float inv_l = args.out_scale / l_tot;
    float* orow = args.ctx + ((int64_t)b * S + qi) * d_model + h * DH + half * (DH / 2);
#pragma unroll 1
    for (int c = 0; c < DH / 32; ++c) {
      float v[16];
      tmem_ld16(tmem + TM_O + lane_addr + half * (DH / 2) + c * 16, v);
      if (row_ok) {
#pragma unroll
        for (int t = 0; t < 16; t += 4)
          *reinterpret_cast<float4*>(orow + c * 16 + t) =
              make_float4(v[t] * inv_l, v[t + 1] * inv_l, v[t + 2] * inv_l, v[t + 3] * inv_l);
      }
    }
    if (row_ok && half == 0 && args.lse)
      args.lse[((int64_t)b * H + h) * S + qi] = (WRITE_P ? m_run : m_run + log2f(l_tot)) * 0.69314718055994530942f;
    if (WRITE_P && lane == 0) bulk_wait0();                    // P / Pd stores complete before the CTA retires
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int make_tf32_map2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                    int box_rows, bool atom32, const char* what);
int make_f32_store_map3d(CUtensorMap* map, float* base, int64_t cols, int64_t rows, int64_t batches, int64_t ld,
                         int64_t batch_stride, int box_cols, int box_rows, const char* what);

}  // namespace mmt

using namespace mmt;

extern "C" int mmt_attention_fwd(const float* qkv, const float* mask, int32_t B, int32_t H, int32_t S,
                                 int32_t dh, float scale, float p_drop, uint64_t seed, uint32_t site,
                                 float* ctx, float* lse, float* probs, float* probs_drop, int32_t ld_p,
                                 void* stream) {
  MMT_ARG_CHECK(qkv && mask && ctx, MMT_E_ARG, "mmt_attention_fwd: null pointer");
  MMT_ARG_CHECK(dh == DH, MMT_E_SHAPE, "mmt_attention_fwd: head dim %d unsupported (only %d)", dh, DH);
  MMT_ARG_CHECK(B > 0 && H > 0 && S > 0 && B <= 65535 && H <= 65535, MMT_E_SHAPE, "mmt_attention_fwd: bad shape B=%d H=%d S=%d", B, H, S);
  MMT_ARG_CHECK(p_drop >= 0.f && p_drop < 1.f, MMT_E_ARG, "mmt_attention_fwd: p_drop=%f", (double)p_drop);
  const bool write_p = probs != nullptr;
  if (write_p) {
    MMT_ARG_CHECK(S <= KB, MMT_E_UNSUPPORTED, "mmt_attention_fwd: probabilities can be saved for S <= %d only (S=%d)", KB, S);
    MMT_ARG_CHECK(ld_p >= S && (ld_p & 3) == 0, MMT_E_SHAPE, "mmt_attention_fwd: ld_p=%d (S=%d)", ld_p, S);
    MMT_ARG_CHECK(p_drop == 0.f || probs_drop != nullptr, MMT_E_ARG, "mmt_attention_fwd: probs_drop missing");
  }
  const int64_t rows = (int64_t)B * S, cols = 3LL * H * DH;
  CUtensorMap mq, mk, mv, mp, mpd;
  int rc = make_tf32_map2d(&mq, qkv, rows, cols, cols, 32, QM, false, "Q");
  if (rc) return rc;
  rc = make_tf32_map2d(&mk, qkv, rows, cols, cols, 32, KB, false, "K");
  if (rc) return rc;
  rc = make_tf32_map2d(&mv, qkv, rows, cols, cols, 32, 32, true, "V");
  if (rc) return rc;
  mp = mq; mpd = mq;                                           // placeholders when nothing is saved
  if (write_p) {
    rc = make_f32_store_map3d(&mp, probs, ld_p, S, (int64_t)B * H, ld_p, (int64_t)S * ld_p, 16, 32, "P");
    if (rc) return rc;
    mpd = mp;
    if (p_drop > 0.f) {
      rc = make_f32_store_map3d(&mpd, probs_drop, ld_p, S, (int64_t)B * H, ld_p, (int64_t)S * ld_p, 16, 32, "Pd");
      if (rc) return rc;
    }
  }
  AttArgs a;
  a.mask = mask; a.ctx = ctx; a.lse = lse;
  a.B = B; a.H = H; a.S = S;
  a.scale_log2 = scale * tc::kTf32TruncComp * 1.44269504088896340736f;
  a.mask_log2 = -10000.0f * 1.44269504088896340736f;
  a.out_scale = tc::kTf32TruncComp;
  a.p_drop = p_drop; a.inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  a.seed = seed; a.site = site; a.ctr = g_step_ctr;
  constexpr size_t smem_base = Q_BYTES + KV_BYTES + 1024 + 128 + KB * 4 + 4 * QM * 4;
  constexpr size_t smem_save = smem_base + 8 * 2 * 32 * 16 * 4;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_base);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_save);
    if (e != cudaSuccess) return cuda_status(e, "attention_fwd smem attribute");
    configured = true;
  }
  dim3 grid((S + QM - 1) / QM, H, B);
  if (write_p)
    launch_pdl(attention_fwd_kernel<true>, dim3(grid), dim3(ATT_THREADS), smem_save, (cudaStream_t)stream, mq, mk, mv, mp, mpd, a);
  else
    launch_pdl(attention_fwd_kernel<false>, dim3(grid), dim3(ATT_THREADS), smem_base, (cudaStream_t)stream, mq, mk, mv, mp, mpd, a);
  MMT_LAUNCH_CHECK("attention_fwd_kernel");
  return 0;
}
