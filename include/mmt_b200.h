/* mmt_b200 -- C ABI of the B200-native (sm_100a) MMT hot path.
 *
 * The reference (gabeur/mmt) is pure Python/PyTorch and has NO native interface; the seam a
 * maintainer binds is the Python plugin surface `model.model.{CENet, sharded_cross_view_inner_product}`
 * and `model.loss.MaxMarginRankingLoss` (reference train.py:86-93, trainer/trainer.py:27,178-204;
 * SURVEY.md §8(b)).  Each entry point below replaces the ATen op sequence of the cited reference
 * lines; the ctypes binding that calls them lives in mmt_b200/_lib.py and INTEGRATION.md shows
 * the reference-side stub.
 *
 * Conventions
 *   - plain pointers and sizes only: every pointer is a DEVICE pointer to fp32 (or int32 where
 *     noted) owned by the caller (PyTorch owns all memory); `stream` is a cudaStream_t passed as
 *     void*.  All launches are asynchronous on that stream; nothing synchronises.
 *   - return 0 on success, a negative MMT_E* code for argument/shape errors (checked before any
 *     launch), a positive value = cudaError_t.  mmt_last_error() returns the message.
 *   - row-major everywhere; "rows" of the video token matrix are (b * S + s).
 *   - dropout masks are a pure function of (seed, site, row, col) (Philox4x32-10) so backward
 *     entry points regenerate the forward's mask from the same (seed, site).
 */
#ifndef MMT_B200_H_
#define MMT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMT_E_ARG (-1)      /* null pointer / bad enum */
#define MMT_E_SHAPE (-2)    /* unsupported or inconsistent shape */
#define MMT_E_ALIGN (-3)    /* pointer / stride alignment the kernel needs is not met */
#define MMT_E_UNSUPPORTED (-4)

int mmt_version(void);
/* Copies the last error message of this thread's most recent failing call. */
int mmt_last_error(char* buf, size_t len);
/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
int64_t mmt_launch_count(void);
/* Optional device-resident step counter (uint64, caller-owned; NULL disables): every kernel that
 * takes a dropout `seed` adds *dev_counter to it and mmt_adam_step adds it to `step`, so a train
 * step captured in a CUDA graph draws fresh dropout masks / bias corrections on each replay when
 * the graph increments the counter. */
int mmt_set_step_counter(const uint64_t* dev_counter);

/* ---------------------------------------------------------------------------------------------
 * GEMM:  C(m,n) = epilogue( alpha * sum_k A(m,k) * B(n,k) + bias[n] + add(m,n) )
 * Replaces every torch addmm / matmul / bmm on the path: ReduceDim (model/model.py:723-726),
 * Q/K/V, attention-output, intermediate and output dense layers (model/bert.py:137-143,186,218,
 * 234), GatedEmbeddingUnit / ContextGating fc (model/model.py:698,746), moe_fc_txt (:277-279),
 * the per-expert similarity matmuls (:823-824) and all of their autograd backward GEMMs.
 *
 * Operands are addressed by element strides so that transposed (wgrad / dgrad) and head-strided
 * (attention) operands need no copies:
 *   A(m,k) = A[m*a_ms + (k / a_kb)*a_kbs + (k % a_kb)*a_ks]      (a_kb == 0: A[m*a_ms + k*a_ks])
 *   B(n,k) = B[n*b_ns + k*b_ks]
 *   C(m,n) = C[(m / c_mb)*c_mbs + (m % c_mb)*c_ms + n]           (c_mb == 0: C[m*c_ms + n])
 * `add` and `aux` use C's addressing.  Batched: z in [0,batch): z0 = z / batch_inner,
 * z1 = z % batch_inner, operand X is offset by z0*x_bs0 + z1*x_bs1.
 * With MMT_GEMM_SPLIT_K set, weight-gradient shaped problems (few output tiles, long K, dense
 * un-batched C) are split along K across CTAs and reduced with fp32 atomics into a zeroed C.
 * ------------------------------------------------------------------------------------------- */
enum { MMT_EPI_NONE = 0,
       MMT_EPI_GELU = 1,    /* aux <- pre-activation u, C <- gelu_erf(u)        (bert.py:218-219,53) */
       MMT_EPI_DGELU = 2 }; /* C <- value * gelu_erf'(aux)   (autograd of bert.py:53)               */
enum { MMT_GEMM_SPLIT_K = 1 };
enum { MMT_PREC_FP32 = 0,   /* CUDA-core FMA, fp32 operands and accumulation (exact-order class)     */
       MMT_PREC_TF32 = 1,   /* tcgen05 kind::tf32 tensor-core tiles, TMA-fed, fp32 accumulate in TMEM */
       MMT_PREC_BF16 = 2 }; /* EXPERIMENTAL 16-bit operand mode: A and B point to bf16 data (strides in elements,
                               contiguous along k or along m / n, un-batched, M >= 256), kind::f16 MMAs, fp32 C /
                               epilogue operands.
                               Not used by the train step (it cannot hold the 1e-3 bar); mmt_cast_bf16 makes
                               the operand copies */

typedef struct mmt_gemm_desc {
  int32_t M, N, K;
  const float* A; int64_t a_ms, a_ks; int32_t a_kb; int64_t a_kbs;
  const float* B; int64_t b_ns, b_ks;
  float* C;       int64_t c_ms;       int32_t c_mb; int64_t c_mbs;
  const float* bias;
  const float* add;
  float* aux;
  int32_t epilogue;
  float alpha;
  int32_t batch, batch_inner;
  int64_t a_bs0, a_bs1, b_bs0, b_bs1, c_bs0, c_bs1;
  int64_t bias_bs;      /* bias of batch z starts at bias + z*bias_bs */
  int32_t precision;
  float* colsum;        /* optional (MMT_PREC_TF32 only): colsum[z1*colsum_bs + n] += sum_m C(m,n) of the
                           FINAL output values, accumulated atomically by the epilogue -- the bias
                           gradient of the layer whose input gradient this GEMM produces.  Caller zeroes it. */
  int64_t colsum_bs;
  int32_t flags;        /* MMT_GEMM_SPLIT_K: allow a split-K schedule (fp32 atomics, run-to-run
                           summation order not fixed); forward GEMMs leave it clear so that the
                           forward pass is bit-reproducible */
} mmt_gemm_desc;

int mmt_gemm(const mmt_gemm_desc* d, void* stream);

/* out[n] (+)= sum_r X[(r / rb)*rbs + (r % rb)*ld + n]  (rb == 0: X[r*ld + n])
 * -- bias gradients (autograd of every `+ b` above). */
int mmt_colsum(const float* X, int64_t rows, int32_t n, int64_t ld, int32_t rb, int64_t rbs,
               float* out, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Token assembly + BertEmbeddings  (model/model.py:485-567 + model/bert.py:87-105), one kernel.
 * proj   [M, B, T+1, d]  raw ReduceDim GEMM outputs, expert-major: row (k, b, 0) is the projected
 *                   max-pooled feature ([AGG] token), rows (k, b, 1..T) the projected frames.
 *                   This kernel gathers them into token order, L2-normalises (model.py:725, eps
 *                   1e-12), adds position/type embeddings, LayerNorms and applies dropout.
 * ft,ind [M,B,T]    features_t / features_ind per expert (sorted expert order).
 * type_idx [M]      int32 token-type id per expert (utils/util.py:154-247).
 * Outputs: h [B*S,d]; mask [B*S] (1 = attend); pos_ids,type_ids [B*S] int32; inv_norm [B*S];
 *          mean,rstd [B*S] (LayerNorm statistics, saved for backward).
 * ------------------------------------------------------------------------------------------- */
int mmt_embed_ln_fwd(const float* proj, const float* ft, const float* ind, const int32_t* type_idx,
                     const float* pos_emb, const float* type_emb, const float* gamma,
                     const float* beta, int32_t B, int32_t M, int32_t T, int32_t d,
                     int32_t max_pos, float eps, float p_drop, uint64_t seed, uint32_t site,
                     float* h, float* mask, int32_t* pos_ids, int32_t* type_ids, float* inv_norm,
                     float* mean, float* rstd, void* stream);
/* Backward of the above: dh [B*S,d] -> dproj [M,B,T+1,d] (gradient w.r.t. the raw GEMM outputs)
 * and ACCUMULATES into dpos_emb [max_pos,d], dtype_emb [type_vocab,d], dgamma, dbeta [d]. */
int mmt_embed_ln_bwd(const float* dh, const float* proj, const int32_t* pos_ids,
                     const int32_t* type_ids, const float* inv_norm, const float* mean,
                     const float* rstd, const float* pos_emb, const float* type_emb,
                     const float* gamma, int32_t B, int32_t M, int32_t T, int32_t d, float p_drop,
                     uint64_t seed, uint32_t site, float* dproj, float* dpos_emb,
                     float* dtype_emb, float* dgamma, float* dbeta, void* stream);

/* ---------------------------------------------------------------------------------------------
 * y = LayerNorm(dropout(t) + r)   (BertSelfOutput / BertOutput, model/bert.py:186-188, 234-236)
 * t is overwritten in place with z = dropout(t) + r (kept for backward).
 * ------------------------------------------------------------------------------------------- */
int mmt_res_ln_fwd(float* t_inout_z, const float* r, const float* gamma, const float* beta,
                   int64_t rows, int32_t d, float eps, float p_drop, uint64_t seed, uint32_t site,
                   float* y, float* mean, float* rstd, void* stream);
/* dz = LN'(dy) (+ dy2 if non-null: a second upstream gradient added to dy first);
 * dt = dropout-mask * dz (written only when p_drop > 0, else dt may alias / be NULL);
 * ACCUMULATES dgamma, dbeta and dbias (= column sum of dt, the preceding dense layer's bias grad). */
int mmt_res_ln_bwd(const float* dy, const float* dy2, const float* z, const float* mean,
                   const float* rstd, const float* gamma, int64_t rows, int32_t d, float p_drop,
                   uint64_t seed, uint32_t site, float* dz, float* dt, float* dgamma, float* dbeta,
                   float* dbias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention probabilities over materialised scores (model/bert.py:147-164), fp32 path:
 *   Psoft[b,h,i,:] = softmax(scores[b,h,i,:] * scale + (1 - mask[b,:]) * -10000)
 *   Pdrop          = dropout(Psoft)            (written only when p_drop > 0)
 * Rows are padded to `ld` (multiple of 4, >= S) floats; padded columns are written as 0.
 * In-place (scores == Psoft) is allowed.  The tensor-core path fuses this into mmt_attention_*.
 * ------------------------------------------------------------------------------------------- */
int mmt_softmax_mask_fwd(const float* scores, const float* mask, int32_t B, int32_t H, int32_t S,
                         int32_t ld, float scale, float p_drop, uint64_t seed, uint32_t site,
                         float* Psoft, float* Pdrop, void* stream);
/* In place on dP (= gradient w.r.t. Pdrop): dScores = scale * Psoft * (dA - sum_j dA*Psoft),
 * dA = dropout-mask * dP. */
int mmt_softmax_mask_bwd(float* dP_inout, const float* Psoft, int32_t B, int32_t H, int32_t S,
                         int32_t ld, float scale, float p_drop, uint64_t seed, uint32_t site,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused flash-style self-attention forward (model/bert.py:136-172) on tcgen05 tensor cores:
 *   ctx[b,i,h*dh:(h+1)*dh] = dropout(softmax(Q K^T * scale + (1 - mask[b,:]) * -10000)) V
 * qkv [B*S, 3*H*dh] holds Q | K | V column blocks (the fused QKV projection's output); scores and
 * probabilities are formed in TMEM.  With probs == NULL (inference) nothing of size S x S touches
 * HBM.  With probs != NULL (training, S <= 224) the normalised probabilities [B,H,S,ld_p] and, when
 * p_drop > 0, their dropped copy probs_drop are additionally streamed out (write-only) for the
 * backward pass -- what the reference's autograd keeps (bert.py:155-160) -- so that the backward
 * repeats neither Q K^T nor the softmax.  lse [B,H,S] (may be NULL) receives the log-sum-exp of the
 * masked, scaled scores.  dh must be 128.  The dropout mask is the same function of
 * (seed, site, (b*H+h)*S+i, j/8) that mmt_softmax_mask_fwd uses (16 random bits per element).
 * ------------------------------------------------------------------------------------------- */
int mmt_attention_fwd(const float* qkv, const float* mask, int32_t B, int32_t H, int32_t S,
                      int32_t dh, float scale, float p_drop, uint64_t seed, uint32_t site,
                      float* ctx, float* lse, float* probs, float* probs_drop, int32_t ld_p,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Expert read-out + L2 normalisation (model/model.py:583-587, 621-623):
 * v[b,k,:] = normalize(h[b, 1 + k*(T+1), :]).  Backward scatters into dh (other rows zeroed).
 * ------------------------------------------------------------------------------------------- */
int mmt_readout_norm_fwd(const float* h, int32_t B, int32_t S, int32_t M, int32_t T, int32_t d,
                         float* v, float* inv_norm, void* stream);
int mmt_readout_norm_bwd(const float* dv, const float* v, const float* inv_norm, int32_t B,
                         int32_t S, int32_t M, int32_t T, int32_t d, float* dh, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Text GatedEmbeddingUnit tail (model/model.py:697-702, 745-750, 624-625), all M experts at once.
 * X [R, M*d] = fc(text), G [R, M*d] = cg.fc(X)  (the two GEMMs are mmt_gemm calls).
 * y = normalize(normalize(X * sigmoid(BN(G))))  -> E [R, M, d]
 * training != 0: BatchNorm1d uses batch statistics over the R rows (biased variance) and updates
 * running_mean/var in place (momentum, unbiased variance); else uses the running statistics.
 * Saves bn_mean, bn_rstd [M*d], inv_n1, inv_n2 [R*M] and Y [R, M*d] (pre-normalisation) for backward.
 * ------------------------------------------------------------------------------------------- */
int mmt_geu_gate_fwd(const float* X, const float* G, const float* bn_w, const float* bn_b,
                     float* run_mean, float* run_var, int32_t R, int32_t M, int32_t d,
                     int32_t training, float momentum, float bn_eps, float* E, float* Y,
                     float* bn_mean, float* bn_rstd, float* inv_n1, float* inv_n2, void* stream);
/* dE -> dX_direct [R,M*d] (through the x * sigmoid path) and dG [R,M*d] (through BatchNorm);
 * ACCUMULATES dbn_w, dbn_b [M*d]. */
int mmt_geu_gate_bwd(const float* dE, const float* X, const float* G, const float* Y,
                     const float* E, const float* bn_w, const float* bn_b, const float* bn_mean,
                     const float* bn_rstd, const float* inv_n1, const float* inv_n2, int32_t R,
                     int32_t M, int32_t d, int32_t training, float* dX, float* dG, float* dbn_w,
                     float* dbn_b, void* stream);

/* Elementwise dropout out = mask * in / (1-p) (moe_txt_dropout, model/model.py:274). */
int mmt_dropout(const float* in, float* out, int64_t rows, int32_t n, float p, uint64_t seed,
                uint32_t site, void* stream);
/* Text mixture weights (model/model.py:276-281, 618): w [R,M] = L1norm(softmax(logits)) over
 * M <= 32; logits / dlogits rows are `ld` floats apart (ld >= M; padding columns of dlogits are
 * written as 0 so the padded matrix can feed the tensor-core GEMMs). */
int mmt_moe_softmax_fwd(const float* logits, int32_t R, int32_t M, int32_t ld, float* w, void* stream);
int mmt_moe_softmax_bwd(const float* dw, const float* w, int32_t R, int32_t M, int32_t ld,
                        float* dlogits, void* stream);

/* ---------------------------------------------------------------------------------------------
 * sharded_cross_view_inner_product (model/model.py:789-837) given the per-expert dot products
 * dots [M, Nq, Nv] (an mmt_gemm batch): sims[i,j] = sum_m (tw[i,m]*vw[j,m]/norm[i,j]) * dots[m,i,j],
 * norm = sum_m tw*vw with exact zeros replaced by 1e-5.  caps > 1 and merge_avg != 0 averages the
 * caps consecutive caption rows of each video (Nq = Nv*caps) into sims [Nv,Nv]; else sims [Nq,Nv].
 * ------------------------------------------------------------------------------------------- */
int mmt_sims_combine_fwd(const float* dots, const float* tw, const float* vw, int32_t Nq,
                         int32_t Nv, int32_t M, int32_t caps, int32_t merge_avg, float* sims,
                         void* stream);
/* dsims -> ddots [M,Nq,Nv] and dtw [Nq,M] (vw carries no gradient: vid_wgh='none'). */
int mmt_sims_combine_bwd(const float* dsims, const float* dots, const float* tw, const float* vw,
                         int32_t Nq, int32_t Nv, int32_t M, int32_t caps, int32_t merge_avg,
                         float* ddots, float* dtw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MaxMarginRankingLoss (model/loss.py:38-65): loss = mean over the 2n(n-1) off-diagonal terms
 * relu(m - x_ii + x_ij), relu(m - x_ii + x_ji)  (fix_norm != 0), or over all 2n^2 terms.
 * loss is a device scalar; dx (may be NULL) receives d loss / d x  [n,n].
 * HBM-bound streaming kernel: one read of x, one write of dx.
 * ------------------------------------------------------------------------------------------- */
int mmt_max_margin_fwd_bwd(const float* x, int32_t n, float margin, int32_t fix_norm, float* loss,
                           float* dx, float* workspace /* >= 2n+2 floats, zeroed by the call */,
                           void* stream);

/* Fused Adam over a flat parameter buffer (torch.optim.Adam semantics, train.py:95-98):
 * grad_scale multiplies the gradient first (1/world_size for replicated head gradients). */
int mmt_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                  void* stream);

/* fp32 -> bf16 (round to nearest even) copy of n elements, n % 4 == 0, 16-byte aligned buffers:
 * operand producer of the experimental MMT_PREC_BF16 mode. */
int mmt_cast_bf16(const float* in, void* out_bf16, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Retrieval ranks on the device (model/metric.py:26-150 t2v_metrics, :152-230 v2t_metrics): 0-based
 * rank of the ground truth with ties averaged, by exact counting on the fp32 similarities (no sort).
 * sims [Nq, Nv] (rows = captions, video-major, Nq = Nv * caps).  v2t = 0: ranks[Nq], rank of video
 * i / caps in row i.  v2t = 1: ranks[Nv], for video v the best rank of its own captions among ALL
 * captions of column v; valid[Nq] (may be NULL) marks existing captions -- missing ones are neither
 * ranked nor candidates (the reference moves them to distance 1e8); +inf if a video has none.
 * ------------------------------------------------------------------------------------------- */
int mmt_retrieval_ranks(const float* sims, const int32_t* valid, int32_t Nq, int32_t Nv, int32_t v2t,
                        float* ranks, void* stream);

/* =============================================================================================
 * 16-bit operand path (the train step's default).  Every dense product of the step takes fp16 (or,
 * for BASELINE config 5, bf16) operands that their PRODUCERS wrote with round-to-nearest-even
 * conversion, accumulates in fp32 in TMEM and writes fp32 and / or 16-bit results.  fp16 has tf32's
 * 10-bit mantissa at half the bytes; gradients travel in 16-bit tensors multiplied by a power-of-two
 * `scale16` that the consuming GEMM's alpha divides out again (fp16 range), statistics, residual
 * streams, parameter gradients and the loss stay fp32.
 * `seed_ctr` (may be NULL): device-resident uint64 added to `seed` by the kernel, so a recorded
 * launch sequence / CUDA graph draws fresh dropout masks on every replay.
 * ============================================================================================= */
enum { MMT_DT_F16 = 0, MMT_DT_BF16 = 1 };

/* C(m,n) = epilogue(alpha * sum_k A(m,k) B(n,k)) with 16-bit A, B:
 *   v = alpha*acc + bias[n];  GELU: aux16 <- gelu_erf'(v), v = gelu_erf(v);  DGELU: v *= aux16 (the stored
 *   derivative: activation and derivative share one erf / exp evaluation in the forward epilogue);
 *   v *= dropout_mask(seed, site, m, n/4) / (1-p);  v += add(m,n);
 *   C32 <- v;  C16 <- rn16(v * out16_scale);  colsum[n] += colsum_scale * sum_m v.
 * A(m,k) = A[m*a_ld + k] (a_mn = 0) or A[k*a_ld + m] (a_mn = 1, "MN-major": dgrad / wgrad operands are
 * read in place, nothing is transposed in HBM); B likewise.  Pitches and batch strides are in ELEMENTS
 * and must be multiples of 8 (16 bytes); bases 16-byte aligned.  C32, C16, add and aux16 share the row
 * index m, their own pitches, and the batch offsets z0*c_bs0 + z1*c_bs1 (elements).
 * MMT_GEMM_SPLIT_K: plain fp32 C32 only; (tile, k-range) work items reduced with fp32 atomics.
 * Replaces the same reference lines as mmt_gemm (model/bert.py:137-143,186,218,234; model/model.py:
 * 698,723-726,746,277-279 and their autograd products). */
typedef struct mmt_gemm16_desc {
  int32_t M, N, K, dtype;
  const void* A; int64_t a_ld; int32_t a_mn;
  const void* B; int64_t b_ld; int32_t b_mn;
  float* C32; int64_t c32_ld;
  void* C16; int64_t c16_ld; float out16_scale;
  const float* bias;
  const float* add; int64_t add_ld;
  void* aux16; int64_t aux_ld;
  int32_t epilogue; float alpha;
  float p_drop; uint32_t site; uint64_t seed; const uint64_t* seed_ctr;
  int32_t batch, batch_inner;
  int64_t a_bs0, a_bs1, b_bs0, b_bs1, c_bs0, c_bs1, bias_bs;
  float* colsum; float colsum_scale; int64_t colsum_bs;
  int32_t flags;
} mmt_gemm16_desc;
int mmt_gemm16(const mmt_gemm16_desc* d, void* stream);

/* out[r, c] = rn16(mask * in[r*in_ld + c] * scale) for c < cols, 0 for cols <= c < out_cols (row pitch out_ld
 * elements).  p_drop > 0 applies the (seed, site, r, c/4) dropout mask of mmt_dropout (moe_txt_dropout,
 * model/model.py:274).  The per-step weight copy (rows = 1) and every small fp32 -> 16-bit operand copy.
 * out_lo (may be NULL; same shape as out): the second term of a two-term split, x*scale ~= out + out_lo / 2048,
 * for the few small products that need more than one 11-bit significand (the similarity's gradient products,
 * the root of every video / text gradient): A B = Ah Bh + (Ah Bl + Al Bh) / 2048 to ~2^-21. */
int mmt_cast16(const float* in, int64_t rows, int32_t cols, int64_t in_ld, void* out, void* out_lo, int32_t out_cols,
               int64_t out_ld, float scale, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site,
               int32_t dtype, void* stream);

/* ReduceDim operands (model/model.py:426-437): per expert k the max-pooled row and the T frame rows are
 * packed as one 16-bit [B, T+1, ld[k]] matrix (row 0 = maxpool = the [AGG] token's input), columns
 * in[k]..ld[k] zero -- ONE launch for all experts (replaces M torch.cat + M casts). */
#define MMT_MAX_EXPERTS 16
typedef struct mmt_pack_desc {
  const float* feats[MMT_MAX_EXPERTS];   /* [B, T, in[k]] */
  const float* maxp[MMT_MAX_EXPERTS];    /* [B, in[k]] */
  void* out[MMT_MAX_EXPERTS];            /* [B, T+1, ld[k]] 16-bit */
  int32_t in[MMT_MAX_EXPERTS], ld[MMT_MAX_EXPERTS];
  int32_t n, B, T, dtype;
} mmt_pack_desc;
int mmt_pack_inputs16(const mmt_pack_desc* d, void* stream);

/* mmt_embed_ln_fwd that additionally writes h16 [B*S, d] = rn16(h) (the QKV GEMM's operand). */
int mmt_embed_ln16_fwd(const float* proj, const float* ft, const float* ind, const int32_t* type_idx,
                       const float* pos_emb, const float* type_emb, const float* gamma, const float* beta,
                       int32_t B, int32_t M, int32_t T, int32_t d, int32_t max_pos, float eps, float p_drop,
                       uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float* h, void* h16, float* mask,
                       int32_t* pos_ids, int32_t* type_ids, float* inv_norm, float* mean, float* rstd,
                       int32_t dtype, void* stream);
/* mmt_embed_ln_bwd that additionally writes dproj16 = rn16(dproj * scale16) (ReduceDim wgrad operand). */
int mmt_embed_ln16_bwd(const float* dh, const float* proj, const int32_t* pos_ids, const int32_t* type_ids,
                       const float* inv_norm, const float* mean, const float* rstd, const float* pos_emb,
                       const float* type_emb, const float* gamma, int32_t B, int32_t M, int32_t T, int32_t d,
                       float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float* dproj,
                       void* dproj16, float scale16, float* dpos_emb, float* dtype_emb, float* dgamma,
                       float* dbeta, int32_t dtype, void* stream);

/* y = LayerNorm(z) (model/bert.py:188, 236) where z = dropout(dense) + residual was written by the GEMM
 * epilogue; writes y (fp32, the next residual) and y16 (the next GEMM operand), saves mean / rstd. */
int mmt_ln16_fwd(const float* z, const float* gamma, const float* beta, int64_t rows, int32_t d, float eps,
                 float* y, void* y16, float* mean, float* rstd, int32_t dtype, void* stream);
/* dz = LN'(dy (+ dy2)) in fp32 (the residual branch's gradient); dt16 = rn16(dropout_mask * dz * scale16)
 * (the dense layer's output gradient: operand of its dgrad / wgrad GEMMs); ACCUMULATES dgamma, dbeta and
 * dbias (= column sums of mask * dz). */
int mmt_ln16_bwd(const float* dy, const float* dy2, const float* z, const float* mean, const float* rstd,
                 const float* gamma, int64_t rows, int32_t d, float p_drop, uint64_t seed,
                 const uint64_t* seed_ctr, uint32_t site, float* dz, void* dt16, float scale16, float* dgamma,
                 float* dbeta, float* dbias, int32_t dtype, void* stream);

/* Fused self-attention (model/bert.py:136-172) on 16-bit operands, nothing of size S x S in HBM:
 * forward  ctx16[b,i,h*dh:(h+1)*dh] = dropout(softmax(Q K^T * scale + (1-mask) * -10000)) V, lse [B,H,S];
 * backward recomputes the probabilities from qkv16 and lse (the dropout decisions from (seed, site)) and writes
 *          dqkv16 = scale16-domain gradients of Q | K | V (same layout as qkv16), accumulating the QKV bias
 *          gradient dbias [3*H*dh] (fp32, divided by scale16).  dctx16 carries scale16 already.
 * dh must be 128, d = H*dh.  dq32 is a zeroed fp32 workspace [B*S, H*dh] the key-tile CTAs reduce dQ into (TMA
 * reduce-add); the call leaves it zeroed again.  delta [B,H,S] is scratch (dO . O per query). */
int mmt_attention16_fwd(const void* qkv16, const float* mask, int32_t B, int32_t H, int32_t S, int32_t dh,
                        float scale, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site,
                        void* ctx16, float* lse, int32_t dtype, void* stream);
int mmt_attention16_bwd(const void* qkv16, const void* ctx16, const void* dctx16, const float* lse,
                        const float* mask, int32_t B, int32_t H, int32_t S, int32_t dh, float scale, float p_drop,
                        uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float scale16, void* dqkv16,
                        float* dq32, float* delta, float* dbias, int32_t dtype, void* stream);

/* Adam that also refreshes the 16-bit weight copy: p16[i] = rn16(p[i]) after the update (p16 may be NULL). */
int mmt_adam16_step(float* p, const float* g, float* m, float* v, void* p16, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, const uint64_t* step_ctr,
                    float grad_scale, int32_t dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Text encoder: the reference's `txt_bert` is transformers' BertModel (bert-base-cased
 * geometry, model/model.py:150-193 construction / freezing, :350-387 call).  Its encoder layers have the
 * video encoder's algebra and run on the same kernels (mmt_gemm16 with fused epilogues, mmt_ln16_*); the
 * entry points below are the parts that differ: token-id embeddings and attention over W <= 128 tokens with
 * dh = 64.
 * ------------------------------------------------------------------------------------------- */
/* h[r] = dropout(LayerNorm(word[ids[r]] + pos[r % W] + type0)), rows = R*W; writes fp32 h and 16-bit h16. */
int mmt_txt_embed_ln_fwd(const int32_t* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                         const float* beta, int64_t rows, int32_t W, int32_t vocab, int32_t d, float eps, float p_drop,
                         uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float* h, void* h16, float* mean,
                         float* rstd, int32_t dtype, void* stream);
/* backward: ACCUMULATES (atomics) dword [vocab,d], dpos [.,d], dtype0 [d], dgamma, dbeta; table pointers may be NULL
 * (frozen embeddings). */
int mmt_txt_embed_ln_bwd(const float* dh, const int32_t* ids, const float* word, const float* pos, const float* type0,
                         const float* mean, const float* rstd, const float* gamma, int64_t rows, int32_t W, int32_t vocab,
                         int32_t d, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site, float* dword,
                         float* dpos, float* dtype0, float* dgamma, float* dbeta, void* stream);
/* ctx16 = dropout(softmax(Q K^T * scale + (1 - mask) * -10000)) V per (caption, head); qkv16 [R*W, 3*H*64]. */
int mmt_txt_attention_fwd(const void* qkv16, const float* mask, int32_t R, int32_t H, int32_t W, int32_t dh, float scale,
                          float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site, void* ctx16, int32_t dtype,
                          void* stream);
/* dqkv16 (Q | K | V blocks, in dctx16's scale16 domain) from qkv16 and dctx16; probabilities are recomputed. */
int mmt_txt_attention_bwd(const void* qkv16, const void* dctx16, const float* mask, int32_t R, int32_t H, int32_t W,
                          int32_t dh, float scale, float p_drop, uint64_t seed, const uint64_t* seed_ctr, uint32_t site,
                          void* dqkv16, int32_t dtype, void* stream);
/* out[n] += scale * sum_r X16[r*ld + n]: bias gradients from a 16-bit gradient tensor. */
int mmt_colsum16(const void* X16, int64_t rows, int32_t n, int64_t ld, float scale, float* out, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMT_B200_H_ */
