"""Kernel sequencing of the MMT hot path (forward and hand-written backward) over the C ABI.

This is the host side of the product path: pure pointer plumbing.  Every arithmetic step is a
kernel in libmmt_b200.so; PyTorch only allocates the buffers.  Reference semantics (file:line) are
cited at each stage; the math is restated in oracle/mmt_oracle.py and SURVEY.md Appendix A.
"""
import math
import os

import torch

from . import _lib
from ._lib import EPI_DGELU, EPI_GELU, PREC_FP32, PREC_TF32, check, gemm, ptr, stream_ptr

# dropout "sites" (Philox stream ids); layer l uses SITE_LAYER + 4*l + {0: attn probs, 1: after
# attention output dense, 2: after FFN output dense}
SITE_EMBED = 1
SITE_MOE_TXT = 2
SITE_LAYER = 16


class Config:
  """Static geometry of one model instance."""

  def __init__(self, layout, vid_bert_params, type_idx, txt_dropout):
    self.layout = layout
    self.mods = layout.mods
    self.M = len(layout.mods)
    self.d = layout.d
    self.ff = layout.ff
    self.L = layout.L
    self.H = vid_bert_params["num_attention_heads"]
    self.dh = self.d // self.H
    self.text_dim = layout.text_dim
    self.max_pos = layout.max_pos
    self.eps = float(vid_bert_params["layer_norm_eps"])
    self.p_hidden = float(vid_bert_params["hidden_dropout_prob"])
    self.p_attn = float(vid_bert_params["attention_probs_dropout_prob"])
    self.p_txt = float(txt_dropout)
    self.type_idx = type_idx            # list of int per expert (sorted order)
    self.in_dims = [layout.expert_dims[m]["dim"] for m in layout.mods]
    # PREC_TF32 (default): every linear layer and the attention matmuls run on the tcgen05
    # tensor-core kernels (tf32 operands, fp32 accumulate; outputs within 1e-3 of the fp32
    # reference, tests/test_gpu_parity.py).  PREC_FP32: CUDA-core fp32 FMAs everywhere (~1e-6).
    #   PREC_F16 (default) / PREC_BF16: the 16-bit operand path of engine16.py -- producers emit round-to-nearest
    #   fp16 (tf32's 10-bit mantissa at half the bytes) or bf16 (BASELINE config 5) operand copies, fp32
    #   accumulation / statistics / residuals / gradients, fused attention forward AND backward.
    self.precision = _lib.PREC_F16
    self.attn_precision = None          # attention matmuls (fp32 / tf32 modes); None -> same as `precision`
    self.w16 = None                     # engine16.Weights16 (16-bit weight copies), owned by the module
    self.seed_ctr = None                # device pointer of a uint64 step counter added to every dropout seed
    self.scale16_override = None
    # flash-style fused attention forward (scores / probabilities never leave TMEM); the backward
    # pass recomputes the probabilities.  Needs the tf32 attention path and dh == 128.
    self.fused_attention = True
    # training: the fused forward also streams P / dropout(P) out for the backward pass (S <= 224);
    # False = keep nothing of size S x S and recompute Q K^T + softmax in the backward instead
    self.save_attention_probs = os.environ.get("MMT_SAVE_PROBS", "1") != "0"


  def enc_spec(self):
    """engine16.EncSpec of the video encoder (parameter names of reference model/bert.py)."""
    from .engine16 import EncSpec
    if self.__dict__.get("_enc_spec") is None:
      self._enc_spec = EncSpec(self.layout, "vid_bert.encoder.layer.%d.", "layer_norm", self.d, self.ff, self.H, self.L,
                               self.eps, SITE_LAYER)
    return self._enc_spec

  @property
  def scale16(self):
    """Power-of-two factor carried by 16-bit gradient tensors (fp16 range); 1 for bf16."""
    if self.scale16_override is not None:
      return float(self.scale16_override)
    return 65536.0 if self.precision == _lib.PREC_F16 else 1.0


def _empty(shape, like, dtype=torch.float32):
  return torch.empty(shape, device=like.device, dtype=dtype)


_SIDE = {}


def _side_streams(device, n):
  key = (device.index, n)
  if key not in _SIDE:
    _SIDE[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
  return _SIDE[key]


class Saved:
  """Activations kept between forward and backward (all torch-owned device tensors)."""
  pass


def _attention_probs(cfg, lib, st, qkv, mask, B, H, S, Sp, d, dh, scale, p_att, seed, site, aprec):
  """Materialised attention probabilities (unfused path, and the backward recompute of the fused
  one): scores = Q K^T per (b, h) (bert.py:147), then scale + mask + softmax + dropout."""
  P = _empty((B, H, S, Sp), qkv)
  gemm(S, S, dh, qkv, 3 * d, 1, qkv, 3 * d, 1, P, Sp, b_off=d, batch=B * H, batch_inner=H,
       a_bs=(S * 3 * d, dh), b_bs=(S * 3 * d, dh), c_bs=(H * S * Sp, S * Sp), precision=aprec)
  Pd = _empty((B, H, S, Sp), qkv) if p_att > 0 else P
  check(lib.mmt_softmax_mask_fwd(ptr(P), ptr(mask), B, H, S, Sp, scale, p_att, seed, site, ptr(P),
                                 ptr(Pd) if p_att > 0 else None, st), "mmt_softmax_mask_fwd")
  return P, Pd


def video_forward(cfg, flat, feats, maxp, ft, ind, training, seed):
  """Video encoder forward: ReduceDim -> token assembly -> BertModel -> AGG read-out.

  flat  : flat fp32 parameter buffer (params.Layout)
  feats : list (sorted experts) of [B, T, in_m];  maxp: list of [B, in_m]
  ft/ind: [M, B, T] features_t / features_ind
  Returns (vid [B,M,d] L2-normalised expert embeddings, Saved).
  """
  if _lib.is16(cfg.precision):
    from . import engine16
    return engine16.video_forward(cfg, flat, feats, maxp, ft, ind, training, seed)
  L = cfg.layout
  d, ff, H, dh, M = cfg.d, cfg.ff, cfg.H, cfg.dh, cfg.M
  B, T = feats[0].shape[0], feats[0].shape[1]
  S = 1 + M * (T + 1)
  BS = B * S
  lib = _lib.load()
  prec = cfg.precision
  aprec = cfg.precision if cfg.attn_precision is None else cfg.attn_precision
  p_hid = cfg.p_hidden if training else 0.0
  p_att = cfg.p_attn if training else 0.0
  sv = Saved()
  sv.B, sv.T, sv.S, sv.training, sv.seed = B, T, S, training, seed
  sv.p_hid, sv.p_att = p_hid, p_att

  # ---- K1: ReduceDim (model.py:426-437, 723-726).  Per expert the max-pooled row and the T frame
  # rows are packed as one [B, T+1, in] operand so the projection (and its weight gradient) is a
  # single dense GEMM; outputs are expert-major [M, B, T+1, d] and gathered into token order by
  # the embedding kernel.
  # The M projections are independent and each fills only ~1/5 of the SMs (1984 x 512 outputs), so
  # they are issued round-robin on a few side streams and joined before the embedding kernel.
  R1 = B * (T + 1)
  proj = _empty((M, R1, d), flat)
  sv.xpack = [_empty((B, T + 1, cfg.in_dims[k]), flat) for k in range(M)]   # allocated on the main stream
  main = torch.cuda.current_stream() if flat.is_cuda else None
  side = _side_streams(flat.device, min(4, M)) if (flat.is_cuda and M > 1) else []
  if side:
    fork = torch.cuda.Event()
    fork.record(main)
  for k, mod in enumerate(cfg.mods):
    w_off = L.off("video_dim_reduce.%s.fc.weight" % mod)
    b_off = L.off("video_dim_reduce.%s.fc.bias" % mod)
    din = cfg.in_dims[k]
    xp = sv.xpack[k]

    def project():
      torch.cat((maxp[k].unsqueeze(1), feats[k]), 1, out=xp)    # [B, T+1, in] (data movement only)
      gemm(R1, d, din, xp, din, 1, flat, din, 1, proj, d, b_off=w_off, bias=flat, bias_off=b_off,
           c_off=k * R1 * d, precision=prec)

    if side:
      st_k = side[k % len(side)]
      st_k.wait_event(fork)
      with torch.cuda.stream(st_k):
        project()
    else:
      project()
  for st_k in side:
    main.wait_stream(st_k)
  sv.proj = proj

  # ---- K2+K3: token assembly + BertEmbeddings (model.py:485-567, bert.py:87-105) ----
  st = stream_ptr()
  h = _empty((BS, d), flat)
  sv.mask = _empty((BS,), flat)
  sv.pos_ids = _empty((BS,), flat, torch.int32)
  sv.type_ids = _empty((BS,), flat, torch.int32)
  sv.inv_norm = _empty((BS,), flat)
  sv.mean0, sv.rstd0 = _empty((BS,), flat), _empty((BS,), flat)
  e = "vid_bert.embeddings."
  check(lib.mmt_embed_ln_fwd(
      ptr(proj), ptr(ft), ptr(ind), ptr(cfg.type_idx_dev), ptr(flat, L.off(e + "position_embeddings.weight")),
      ptr(flat, L.off(e + "token_type_embeddings.weight")), ptr(flat, L.off(e + "layer_norm.weight")),
      ptr(flat, L.off(e + "layer_norm.bias")), B, M, T, d, cfg.max_pos, cfg.eps, p_hid, seed,
      SITE_EMBED, ptr(h), ptr(sv.mask), ptr(sv.pos_ids), ptr(sv.type_ids), ptr(sv.inv_norm),
      ptr(sv.mean0), ptr(sv.rstd0), st), "mmt_embed_ln_fwd")

  # ---- encoder layers (bert.py:249-256) ----
  Sp = (S + 3) // 4 * 4
  scale = 1.0 / math.sqrt(dh)
  sv.Sp = Sp
  sv.layers = []
  for l in range(cfg.L):
    p = "vid_bert.encoder.layer.%d." % l
    ls = Saved()
    ls.h_in = h
    # K4: fused QKV projection, [BS, 3d] (bert.py:137-143)
    qkv = _empty((BS, 3 * d), flat)
    gemm(BS, 3 * d, d, h, d, 1, flat, d, 1, qkv, 3 * d, b_off=L.off(p + "attention.self.query.weight"),
         bias=flat, bias_off=L.off(p + "attention.self.query.bias"), precision=prec)
    fused = cfg.fused_attention and aprec == PREC_TF32 and dh == 128
    ctx = _empty((BS, d), flat)
    if fused:
      # K5 fused: softmax(Q K^T / sqrt(dh) + mask) V in one tcgen05 kernel (bert.py:147-170)
      P = Pd = None
      if training and cfg.save_attention_probs and S <= 224:
        P = _empty((B, H, S, Sp), flat)
        Pd = _empty((B, H, S, Sp), flat) if p_att > 0 else P
      check(lib.mmt_attention_fwd(ptr(qkv), ptr(sv.mask), B, H, S, dh, scale, p_att, seed,
                                  SITE_LAYER + 4 * l, ptr(ctx), None, ptr(P) if P is not None else None,
                                  ptr(Pd) if (P is not None and p_att > 0) else None, Sp, st),
            "mmt_attention_fwd")
    else:
      P, Pd = _attention_probs(cfg, lib, st, qkv, sv.mask, B, H, S, Sp, d, dh, scale, p_att, seed,
                               SITE_LAYER + 4 * l, aprec)
      # ctx = P V, heads merged by the output addressing (bert.py:166-170)
      gemm(S, dh, S, Pd, Sp, 1, qkv, 1, 3 * d, ctx, d, b_off=2 * d, batch=B * H, batch_inner=H,
           a_bs=(H * S * Sp, S * Sp), b_bs=(S * 3 * d, dh), c_bs=(S * d, dh), precision=aprec)
    # K6: attention output dense + dropout + residual + LN (bert.py:186-188)
    z1 = _empty((BS, d), flat)
    gemm(BS, d, d, ctx, d, 1, flat, d, 1, z1, d, b_off=L.off(p + "attention.output.dense.weight"),
         bias=flat, bias_off=L.off(p + "attention.output.dense.bias"), precision=prec)
    a = _empty((BS, d), flat)
    ls.mean1, ls.rstd1 = _empty((BS,), flat), _empty((BS,), flat)
    check(lib.mmt_res_ln_fwd(ptr(z1), ptr(h), ptr(flat, L.off(p + "attention.output.layer_norm.weight")),
                             ptr(flat, L.off(p + "attention.output.layer_norm.bias")), BS, d, cfg.eps,
                             p_hid, seed, SITE_LAYER + 4 * l + 1, ptr(a), ptr(ls.mean1), ptr(ls.rstd1),
                             st), "mmt_res_ln_fwd")
    # K7: FFN up + erf-GELU (bert.py:218-219, 53); u = pre-activation kept for backward
    u, f = _empty((BS, ff), flat), _empty((BS, ff), flat)
    gemm(BS, ff, d, a, d, 1, flat, d, 1, f, ff, b_off=L.off(p + "intermediate.dense.weight"),
         bias=flat, bias_off=L.off(p + "intermediate.dense.bias"), epilogue=EPI_GELU, aux=u,
         precision=prec)
    # K8: FFN down + dropout + residual + LN (bert.py:234-236)
    z2 = _empty((BS, d), flat)
    gemm(BS, d, ff, f, ff, 1, flat, ff, 1, z2, d, b_off=L.off(p + "output.dense.weight"),
         bias=flat, bias_off=L.off(p + "output.dense.bias"), precision=prec)
    hn = _empty((BS, d), flat)
    ls.mean2, ls.rstd2 = _empty((BS,), flat), _empty((BS,), flat)
    check(lib.mmt_res_ln_fwd(ptr(z2), ptr(a), ptr(flat, L.off(p + "output.layer_norm.weight")),
                             ptr(flat, L.off(p + "output.layer_norm.bias")), BS, d, cfg.eps, p_hid,
                             seed, SITE_LAYER + 4 * l + 2, ptr(hn), ptr(ls.mean2), ptr(ls.rstd2), st),
          "mmt_res_ln_fwd")
    ls.qkv, ls.P, ls.Pd, ls.ctx, ls.z1, ls.a, ls.u, ls.f, ls.z2 = qkv, P, Pd, ctx, z1, a, u, f, z2
    sv.layers.append(ls)
    h = hn
  sv.h_last = h

  # ---- K10: expert read-out + L2 norm (model.py:583-587, 621-623) ----
  vid = _empty((B, M, d), flat)
  sv.vinv = _empty((B * M,), flat)
  check(lib.mmt_readout_norm_fwd(ptr(h), B, S, M, T, d, ptr(vid), ptr(sv.vinv), st),
        "mmt_readout_norm_fwd")
  sv.vid = vid
  return vid, sv


def head_forward(cfg, flat, bufs, text, training, seed):
  """Text head forward: M GatedEmbeddingUnits (+BatchNorm over the R rows) and mixture weights.

  bufs : flat fp32 buffer with BatchNorm running statistics [2*M*d]
  text : [R, text_dim]  (txt_bert CLS features, R = B*caps)        model.py:371-379
  Returns (txt [R,M,d], tw [R,M], Saved).
  """
  if _lib.is16(cfg.precision):
    from . import engine16
    return engine16.head_forward(cfg, flat, bufs, text, training, seed)
  L = cfg.layout
  d, M = cfg.d, cfg.M
  R = text.shape[0]
  st = stream_ptr()
  lib = _lib.load()
  prec = cfg.precision
  p_txt = cfg.p_txt if training else 0.0
  sv = Saved()
  sv.R, sv.training, sv.seed, sv.p_txt, sv.text = R, training, seed, p_txt, text

  # ---- K11: text GatedEmbeddingUnits, all experts at once (model.py:413-417, 697-702, 745-750) ----
  td = cfg.text_dim
  m0 = cfg.mods[0]
  X = _empty((R, M * d), flat)
  gemm(R, M * d, td, text, td, 1, flat, td, 1, X, M * d, b_off=L.off("text_GU.%s.fc.weight" % m0),
       bias=flat, bias_off=L.off("text_GU.%s.fc.bias" % m0), precision=prec)
  G = _empty((R, M * d), flat)
  gemm(R, d, d, X, M * d, 1, flat, d, 1, G, M * d, b_off=L.off("text_GU.%s.cg.fc.weight" % m0),
       bias=flat, bias_off=L.off("text_GU.%s.cg.fc.bias" % m0), bias_bs=d, batch=M,
       a_bs=(d, 0), b_bs=(d * d, 0), c_bs=(d, 0), precision=prec)
  txt = _empty((R, M, d), flat)
  sv.Y = _empty((R, M * d), flat)
  sv.bn_mean, sv.bn_rstd = _empty((M * d,), flat), _empty((M * d,), flat)
  sv.n1, sv.n2 = _empty((R * M,), flat), _empty((R * M,), flat)
  check(lib.mmt_geu_gate_fwd(
      ptr(X), ptr(G), ptr(flat, L.off("text_GU.%s.cg.batch_norm.weight" % m0)),
      ptr(flat, L.off("text_GU.%s.cg.batch_norm.bias" % m0)), ptr(bufs), ptr(bufs, M * d), R, M, d,
      1 if training else 0, 0.1, 1e-5, ptr(txt), ptr(sv.Y), ptr(sv.bn_mean), ptr(sv.bn_rstd),
      ptr(sv.n1), ptr(sv.n2), st), "mmt_geu_gate_fwd")
  sv.X, sv.G, sv.txt = X, G, txt

  # ---- K12: text mixture weights (model.py:273-281, 618) ----
  if p_txt > 0:
    tdrop = _empty((R, td), flat)
    check(lib.mmt_dropout(ptr(text), ptr(tdrop), R, td, p_txt, seed, SITE_MOE_TXT, st), "mmt_dropout")
  else:
    tdrop = text
  Mp = (M + 3) // 4 * 4                                # logits rows padded to 16 B for the TMA path
  logits = _empty((R, Mp), flat)
  gemm(R, M, td, tdrop, td, 1, flat, td, 1, logits, Mp, b_off=L.off("moe_fc_txt.%s.weight" % m0),
       bias=flat, bias_off=L.off("moe_fc_txt.%s.bias" % m0), precision=prec)
  tw = _empty((R, M), flat)
  check(lib.mmt_moe_softmax_fwd(ptr(logits), R, M, Mp, ptr(tw), st), "mmt_moe_softmax_fwd")
  sv.tdrop, sv.tw = tdrop, tw
  return txt, tw, sv


def zero_small_grads(cfg, gflat):
  """The small-parameter region of the flat gradient is accumulated with atomics: zero it once
  per step, before head_backward / video_backward."""
  gflat[:cfg.layout.small_numel].zero_()


def head_backward(cfg, flat, gflat, sv, dtxt, dtw, need_dtext=True):
  """Backward of head_forward: text-head parameter gradients into `gflat`; returns
  d loss / d text [R, text_dim] (or None)."""
  if _lib.is16(cfg.precision):
    from . import engine16
    return engine16.head_backward(cfg, flat, gflat, sv, dtxt, dtw, need_dtext)
  L = cfg.layout
  d, M = cfg.d, cfg.M
  R = sv.R
  td = cfg.text_dim
  st = stream_ptr()
  lib = _lib.load()
  seed = sv.seed
  m0 = cfg.mods[0]
  prec = cfg.precision

  def colsum(X, rows, n, ld, out_off, rb=0, rbs=0, x_off=0):
    check(lib.mmt_colsum(ptr(X, x_off), rows, n, ld, rb, rbs, ptr(gflat, out_off), 1, st), "mmt_colsum")

  dtext = None
  if dtw is not None:
    Mp = (M + 3) // 4 * 4
    dlog = _empty((R, Mp), flat)
    check(lib.mmt_moe_softmax_bwd(ptr(dtw), ptr(sv.tw), R, M, Mp, ptr(dlog), st), "mmt_moe_softmax_bwd")
    # dW_moe [M, td] = dlog^T @ tdrop ; db = colsum(dlog)
    gemm(M, td, R, dlog, 1, Mp, sv.tdrop, 1, td, gflat, td, c_off=L.off("moe_fc_txt.%s.weight" % m0),
         precision=prec)
    colsum(dlog, R, M, Mp, L.off("moe_fc_txt.%s.bias" % m0))
    if need_dtext:
      dtext = _empty((R, td), flat)
      gemm(R, td, M, dlog, Mp, 1, flat, 1, td, dtext, td, b_off=L.off("moe_fc_txt.%s.weight" % m0),
           precision=prec)
      if sv.p_txt > 0:
        check(lib.mmt_dropout(ptr(dtext), ptr(dtext), R, td, sv.p_txt, seed, SITE_MOE_TXT, st),
              "mmt_dropout")
  if dtxt is not None:
    dX = _empty((R, M * d), flat)
    dG = _empty((R, M * d), flat)
    check(lib.mmt_geu_gate_bwd(
        ptr(dtxt), ptr(sv.X), ptr(sv.G), ptr(sv.Y), ptr(sv.txt),
        ptr(flat, L.off("text_GU.%s.cg.batch_norm.weight" % m0)),
        ptr(flat, L.off("text_GU.%s.cg.batch_norm.bias" % m0)), ptr(sv.bn_mean), ptr(sv.bn_rstd),
        ptr(sv.n1), ptr(sv.n2), R, M, d, 1 if sv.training else 0, ptr(dX), ptr(dG),
        ptr(gflat, L.off("text_GU.%s.cg.batch_norm.weight" % m0)),
        ptr(gflat, L.off("text_GU.%s.cg.batch_norm.bias" % m0)), st), "mmt_geu_gate_bwd")
    # cg.fc: dW2_m [d,d] = dG_m^T @ X_m ; db2 = colsum(dG) ; dX += dG_m @ W2_m
    gemm(d, d, R, dG, 1, M * d, sv.X, 1, M * d, gflat, d, c_off=L.off("text_GU.%s.cg.fc.weight" % m0),
         batch=M, a_bs=(d, 0), b_bs=(d, 0), c_bs=(d * d, 0), precision=prec)
    colsum(dG, R, M * d, M * d, L.off("text_GU.%s.cg.fc.bias" % m0))
    gemm(R, d, d, dG, M * d, 1, flat, 1, d, dX, M * d, b_off=L.off("text_GU.%s.cg.fc.weight" % m0),
         add=dX, batch=M, a_bs=(d, 0), b_bs=(d * d, 0), c_bs=(d, 0), precision=prec)
    # fc: dW1 [M*d, td] = dX^T @ text ; db1 = colsum(dX) ; dtext += dX @ W1
    gemm(M * d, td, R, dX, 1, M * d, sv.text, 1, td, gflat, td, c_off=L.off("text_GU.%s.fc.weight" % m0),
         precision=prec)
    colsum(dX, R, M * d, M * d, L.off("text_GU.%s.fc.bias" % m0))
    if need_dtext:
      if dtext is None:
        dtext = _empty((R, td), flat)
        gemm(R, td, M * d, dX, M * d, 1, flat, 1, td, dtext, td, b_off=L.off("text_GU.%s.fc.weight" % m0),
             precision=prec)
      else:
        gemm(R, td, M * d, dX, M * d, 1, flat, 1, td, dtext, td, b_off=L.off("text_GU.%s.fc.weight" % m0),
             add=dtext, precision=prec)

  return dtext


def video_backward(cfg, flat, gflat, sv, dvid, on_layer_done=None):
  """Backward of video_forward: every video-side parameter gradient into `gflat` (small region
  accumulated -- zero it first with zero_small_grads -- big matrices overwritten).
  `on_layer_done(l)` is called once layer l's weight-matrix gradients have been enqueued (the
  data-parallel path starts their all-reduce there)."""
  if _lib.is16(cfg.precision):
    from . import engine16
    return engine16.video_backward(cfg, flat, gflat, sv, dvid, on_layer_done)
  L = cfg.layout
  d, ff, H, dh, M = cfg.d, cfg.ff, cfg.H, cfg.dh, cfg.M
  B, T, S, Sp = sv.B, sv.T, sv.S, sv.Sp
  BS = B * S
  st = stream_ptr()
  lib = _lib.load()
  prec = cfg.precision
  aprec = cfg.precision if cfg.attn_precision is None else cfg.attn_precision
  seed = sv.seed
  scale = 1.0 / math.sqrt(dh)

  def colsum(X, rows, n, ld, out_off, rb=0, rbs=0, x_off=0):
    check(lib.mmt_colsum(ptr(X, x_off), rows, n, ld, rb, rbs, ptr(gflat, out_off), 1, st), "mmt_colsum")

  dh_ = _empty((BS, d), flat)
  check(lib.mmt_readout_norm_bwd(ptr(dvid), ptr(sv.vid), ptr(sv.vinv), B, S, M, T, d, ptr(dh_), st),
        "mmt_readout_norm_bwd")
  for l in reversed(range(cfg.L)):
    p = "vid_bert.encoder.layer.%d." % l
    ls = sv.layers[l]
    p_hid, p_att = sv.p_hid, sv.p_att
    # --- LN2 backward: dz2 (to residual a), dt2 (to FFN-down output), db2 ---
    dz2 = _empty((BS, d), flat)
    dt2 = _empty((BS, d), flat) if p_hid > 0 else dz2
    check(lib.mmt_res_ln_bwd(ptr(dh_), None, ptr(ls.z2), ptr(ls.mean2), ptr(ls.rstd2),
                             ptr(flat, L.off(p + "output.layer_norm.weight")), BS, d, p_hid, seed,
                             SITE_LAYER + 4 * l + 2, ptr(dz2), ptr(dt2) if p_hid > 0 else None,
                             ptr(gflat, L.off(p + "output.layer_norm.weight")),
                             ptr(gflat, L.off(p + "output.layer_norm.bias")),
                             ptr(gflat, L.off(p + "output.dense.bias")), st), "mmt_res_ln_bwd")
    # FFN down: dW2 [d, ff] = dt2^T @ f ; du = (dt2 @ W2) * gelu'(u)
    gemm(d, ff, BS, dt2, 1, d, ls.f, 1, ff, gflat, ff, c_off=L.off(p + "output.dense.weight"),
         precision=prec, split_k=True)
    du = _empty((BS, ff), flat)
    fuse_cs = (prec == PREC_TF32)        # bias gradients as a fused column-sum epilogue (tensor-core path)
    gemm(BS, ff, d, dt2, d, 1, flat, 1, ff, du, ff, b_off=L.off(p + "output.dense.weight"),
         epilogue=EPI_DGELU, aux=ls.u, precision=prec,
         colsum=gflat if fuse_cs else None, colsum_off=L.off(p + "intermediate.dense.bias"))
    # FFN up: dW1 [ff, d] = du^T @ a ; db1 = colsum(du) ; da = du @ W1
    gemm(ff, d, BS, du, 1, ff, ls.a, 1, d, gflat, d, c_off=L.off(p + "intermediate.dense.weight"),
         precision=prec, split_k=True)
    if not fuse_cs:
      colsum(du, BS, ff, ff, L.off(p + "intermediate.dense.bias"))
    da = _empty((BS, d), flat)
    gemm(BS, d, ff, du, ff, 1, flat, 1, d, da, d, b_off=L.off(p + "intermediate.dense.weight"),
         precision=prec)
    # --- LN1 backward on (da + dz2) ---
    dz1 = _empty((BS, d), flat)
    dt1 = _empty((BS, d), flat) if p_hid > 0 else dz1
    check(lib.mmt_res_ln_bwd(ptr(da), ptr(dz2), ptr(ls.z1), ptr(ls.mean1), ptr(ls.rstd1),
                             ptr(flat, L.off(p + "attention.output.layer_norm.weight")), BS, d, p_hid,
                             seed, SITE_LAYER + 4 * l + 1, ptr(dz1), ptr(dt1) if p_hid > 0 else None,
                             ptr(gflat, L.off(p + "attention.output.layer_norm.weight")),
                             ptr(gflat, L.off(p + "attention.output.layer_norm.bias")),
                             ptr(gflat, L.off(p + "attention.output.dense.bias")), st), "mmt_res_ln_bwd")
    # attention output dense: dWo = dt1^T @ ctx ; dctx = dt1 @ Wo
    gemm(d, d, BS, dt1, 1, d, ls.ctx, 1, d, gflat, d, c_off=L.off(p + "attention.output.dense.weight"),
         precision=prec, split_k=True)
    dctx = _empty((BS, d), flat)
    gemm(BS, d, d, dt1, d, 1, flat, 1, d, dctx, d, b_off=L.off(p + "attention.output.dense.weight"),
         precision=prec)
    # --- attention backward (materialised probabilities; recomputed if the forward was fused) ---
    if ls.P is None:
      ls.P, ls.Pd = _attention_probs(cfg, lib, st, ls.qkv, sv.mask, B, H, S, Sp, d, dh, scale, p_att, seed,
                                     SITE_LAYER + 4 * l, aprec)
    dqkv = _empty((BS, 3 * d), flat)
    dP = _empty((B, H, S, Sp), flat)
    bsP = (H * S * Sp, S * Sp)
    bsQ = (S * 3 * d, dh)
    # dP = dctx @ V^T
    gemm(S, S, dh, dctx, d, 1, ls.qkv, 3 * d, 1, dP, Sp, b_off=2 * d, batch=B * H, batch_inner=H,
         a_bs=(S * d, dh), b_bs=bsQ, c_bs=bsP, precision=aprec)
    # dV = Pd^T @ dctx
    fuse_acs = (aprec == PREC_TF32)
    acs = dict(colsum=gflat, colsum_bs=dh) if fuse_acs else {}
    qb_off = L.off(p + "attention.self.query.bias")
    gemm(S, dh, S, ls.Pd, 1, Sp, dctx, 1, d, dqkv, 3 * d, c_off=2 * d, batch=B * H, batch_inner=H,
         a_bs=bsP, b_bs=(S * d, dh), c_bs=bsQ, precision=aprec,
         **(dict(acs, colsum_off=qb_off + 2 * d) if fuse_acs else {}))
    check(lib.mmt_softmax_mask_bwd(ptr(dP), ptr(ls.P), B, H, S, Sp, scale, p_att, seed,
                                   SITE_LAYER + 4 * l, st), "mmt_softmax_mask_bwd")
    # dQ = dS @ K ; dK = dS^T @ Q
    gemm(S, dh, S, dP, Sp, 1, ls.qkv, 1, 3 * d, dqkv, 3 * d, b_off=d, batch=B * H, batch_inner=H,
         a_bs=bsP, b_bs=bsQ, c_bs=bsQ, precision=aprec,
         **(dict(acs, colsum_off=qb_off) if fuse_acs else {}))
    gemm(S, dh, S, dP, 1, Sp, ls.qkv, 1, 3 * d, dqkv, 3 * d, c_off=d, batch=B * H, batch_inner=H,
         a_bs=bsP, b_bs=bsQ, c_bs=bsQ, precision=aprec,
         **(dict(acs, colsum_off=qb_off + d) if fuse_acs else {}))
    # QKV projection: dWqkv [3d, d] = dqkv^T @ h_in ; dbqkv ; dh = dz1 + dqkv @ Wqkv
    gemm(3 * d, d, BS, dqkv, 1, 3 * d, ls.h_in, 1, d, gflat, d,
         c_off=L.off(p + "attention.self.query.weight"), precision=prec, split_k=True)
    if not fuse_acs:
      colsum(dqkv, BS, 3 * d, 3 * d, L.off(p + "attention.self.query.bias"))
    dh_ = _empty((BS, d), flat)
    gemm(BS, d, 3 * d, dqkv, 3 * d, 1, flat, 1, d, dh_, d,
         b_off=L.off(p + "attention.self.query.weight"), add=dz1, precision=prec)
    if on_layer_done is not None:
      on_layer_done(l)

  # --- embeddings + token assembly backward ---
  R1 = B * (T + 1)
  dproj = _empty((M, R1, d), flat)
  e = "vid_bert.embeddings."
  check(lib.mmt_embed_ln_bwd(
      ptr(dh_), ptr(sv.proj), ptr(sv.pos_ids), ptr(sv.type_ids), ptr(sv.inv_norm), ptr(sv.mean0),
      ptr(sv.rstd0), ptr(flat, L.off(e + "position_embeddings.weight")),
      ptr(flat, L.off(e + "token_type_embeddings.weight")), ptr(flat, L.off(e + "layer_norm.weight")),
      B, M, T, d, sv.p_hid, seed, SITE_EMBED, ptr(dproj),
      ptr(gflat, L.off(e + "position_embeddings.weight")),
      ptr(gflat, L.off(e + "token_type_embeddings.weight")), ptr(gflat, L.off(e + "layer_norm.weight")),
      ptr(gflat, L.off(e + "layer_norm.bias")), st), "mmt_embed_ln_bwd")
  # --- ReduceDim weight gradients (inputs carry no gradient): dW [d, in] = dproj_k^T @ xpack_k ---
  for k, mod in enumerate(cfg.mods):
    w_off = L.off("video_dim_reduce.%s.fc.weight" % mod)
    b_off = L.off("video_dim_reduce.%s.fc.bias" % mod)
    din = cfg.in_dims[k]
    gemm(d, din, R1, dproj, 1, d, sv.xpack[k], 1, din, gflat, din, a_off=k * R1 * d, c_off=w_off,
         precision=prec, split_k=True)
    colsum(dproj, R1, d, d, b_off, x_off=k * R1 * d)


def encode_forward(cfg, flat, bufs, text, feats, maxp, ft, ind, training, seed):
  """Single-device composition: (vid [B,M,d], txt [R,M,d], tw [R,M], (video Saved, head Saved))."""
  vid, sv_v = video_forward(cfg, flat, feats, maxp, ft, ind, training, seed)
  txt, tw, sv_h = head_forward(cfg, flat, bufs, text, training, seed)
  return vid, txt, tw, (sv_v, sv_h)


def encode_backward(cfg, flat, gflat, sv, dvid, dtxt, dtw, need_dtext=True):
  """Backward of encode_forward; returns d loss / d text [R, text_dim] (or None)."""
  sv_v, sv_h = sv
  zero_small_grads(cfg, gflat)
  dtext = head_backward(cfg, flat, gflat, sv_h, dtxt, dtw, need_dtext)
  video_backward(cfg, flat, gflat, sv_v, dvid)
  return dtext


SPLIT_DOTS_MIN = 256      # training batches at least this large take the tensor-core split products


def sims_forward(vid, txt, vw, tw, caps, merge_avg, train_precision=None):
  """sharded_cross_view_inner_product (model.py:789-837).  vid [Nv,M,d], txt [Nq,M,d] (Nq = Nv*caps,
  video-major), vw [Nv,M], tw [Nq,M] -> (sims, dots [M,Nq,Nv]).
  The dot products are fp32 FMAs (ranking bit-exact against the reference on identical embeddings) except in
  TRAINING with a 16-bit `train_precision` and a large batch (the data-parallel global batch), where they are
  three-pass split tensor-core products of fp32-class accuracy (engine16.sims_dots_split)."""
  lib = _lib.load()
  Nv, M, d = vid.shape
  Nq = txt.shape[0]
  if train_precision is not None and _lib.is16(train_precision) and Nv >= SPLIT_DOTS_MIN and Nv % 4 == 0 and d % 8 == 0:
    from . import engine16
    dots = engine16.sims_dots_split(_lib.dt_of(train_precision), vid, txt)
  else:
    dots = torch.empty((M, Nq, Nv), device=vid.device, dtype=torch.float32)
    gemm(Nq, Nv, d, txt, M * d, 1, vid, M * d, 1, dots, Nv, batch=M, a_bs=(d, 0), b_bs=(d, 0),
         c_bs=(Nq * Nv, 0))
  rows = Nv if (merge_avg and caps > 1) else Nq
  sims = torch.empty((rows, Nv), device=vid.device, dtype=torch.float32)
  check(lib.mmt_sims_combine_fwd(ptr(dots), ptr(tw), ptr(vw), Nq, Nv, M, caps,
                                 1 if (merge_avg and caps > 1) else 0, ptr(sims), stream_ptr()),
        "mmt_sims_combine_fwd")
  return sims, dots


def sims_backward(dsims, dots, vid, txt, vw, tw, caps, merge_avg, precision=PREC_FP32, scale16=1.0):
  """Backward of sims_forward.  `precision` applies to the two gradient products only (the forward dot
  products always stay fp32 FMAs: ranking at the similarity boundary must be exact)."""
  lib = _lib.load()
  Nv, M, d = vid.shape
  Nq = txt.shape[0]
  ddots = torch.empty_like(dots)
  dtw = torch.empty((Nq, M), device=vid.device, dtype=torch.float32)
  check(lib.mmt_sims_combine_bwd(ptr(dsims), ptr(dots), ptr(tw), ptr(vw), Nq, Nv, M, caps,
                                 1 if (merge_avg and caps > 1) else 0, ptr(ddots), ptr(dtw),
                                 stream_ptr()), "mmt_sims_combine_bwd")
  # dtxt[:, m, :] = ddots_m @ vid_m ; dvid[:, m, :] = ddots_m^T @ txt_m
  if _lib.is16(precision):
    from . import engine16
    dvid, dtxt = engine16.sims_backward_products(_lib.dt_of(precision), ddots, vid, txt, scale16)
    return dvid, dtxt, dtw
  dtxt = torch.empty_like(txt)
  dvid = torch.empty_like(vid)
  tc = precision == PREC_TF32 and Nv % 4 == 0 and Nq % 4 == 0          # TMA strides: 16-byte multiples
  prec = PREC_TF32 if tc else PREC_FP32
  gemm(Nq, d, Nv, ddots, Nv, 1, vid, 1, M * d, dtxt, M * d, batch=M, a_bs=(Nq * Nv, 0), b_bs=(d, 0),
       c_bs=(d, 0), precision=prec)
  gemm(Nv, d, Nq, ddots, 1, Nv, txt, 1, M * d, dvid, M * d, batch=M, a_bs=(Nq * Nv, 0), b_bs=(d, 0),
       c_bs=(d, 0), precision=prec)
  return dvid, dtxt, dtw


def max_margin(sims, margin, fix_norm, want_grad=True):
  """MaxMarginRankingLoss (loss.py:38-65): returns (loss scalar tensor, d loss / d sims or None)."""
  lib = _lib.load()
  n = sims.shape[0]
  loss = torch.empty((), device=sims.device, dtype=torch.float32)
  dx = torch.empty_like(sims) if want_grad else None
  ws = torch.empty((n + 2,), device=sims.device, dtype=torch.float32)
  check(lib.mmt_max_margin_fwd_bwd(ptr(sims), n, float(margin), 1 if fix_norm else 0, ptr(loss),
                                   ptr(dx), ptr(ws), stream_ptr()), "mmt_max_margin_fwd_bwd")
  return loss, dx
