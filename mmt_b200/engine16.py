"""Kernel sequencing of the MMT hot path on 16-bit tensor-core operands (the train step's default).

Same stages, reference lines and saved-activation contract as engine.py (the fp32 / tf32 sequencing); the
differences are the data types on the wire:
  * every GEMM operand is a 16-bit tensor (fp16, or bf16 for BASELINE config 5) written by its PRODUCER with
    round-to-nearest conversion -- activations by the GEMM / LayerNorm / attention epilogues, weights by the
    fused Adam kernel (or one cast pass when another optimizer changed them), inputs by one packing kernel;
  * residual streams, LayerNorm / BatchNorm statistics, the similarity matrix, the loss and all parameter
    gradients stay fp32;
  * gradients w.r.t. activations travel in 16-bit tensors multiplied by `cfg.scale16` (2^16 for fp16, 1 for
    bf16); the GEMM that turns them into an fp32 result divides it out through `alpha`;
  * attention is two fused kernels (forward / backward); nothing of size S x S exists in HBM.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import EPI_DGELU, EPI_GELU, check, gemm16, cast16, ptr, stream_ptr
from .engine import SITE_EMBED, SITE_LAYER, SITE_MOE_TXT, Saved, _empty, _side_streams, zero_small_grads


def _pad8(n):
  return (n + 7) // 8 * 8


class Weights16:
  """16-bit copy of the flat parameter buffer (same element offsets) plus row-padded copies of the ReduceDim
  matrices whose input width is not a multiple of 8 (TMA needs 16-byte row pitches)."""

  def __init__(self, cfg, flat):
    self.dt = _lib.dt_of(cfg.precision)
    tdt = _lib.torch_dtype(self.dt)
    self.flat16 = torch.zeros(flat.numel(), device=flat.device, dtype=tdt)
    self.red = {}
    for k, din in enumerate(cfg.in_dims):
      if din % 8 != 0:
        self.red[k] = torch.zeros((cfg.d, _pad8(din)), device=flat.device, dtype=tdt)
    self.sig = None                         # parameter-version signature the copy was taken at

  def refresh(self, cfg, flat, sig, force=False):
    """flat -> flat16 (one pass over 100 MB: ~25 us) unless the copy is known to be current: `sig` is the sum of
    the parameters' torch version counters (every torch-side in-place change bumps one; FusedAdam, which
    bypasses torch, refreshes the copy inside its own kernel)."""
    if not force and self.sig is not None and self.sig == sig:
      return
    cast16(self.dt, flat, 1, flat.numel(), flat.numel(), self.flat16, flat.numel(), flat.numel())
    self.refresh_padded(cfg, flat)
    self.sig = sig

  def refresh_padded(self, cfg, flat):
    L = cfg.layout
    for k, w in self.red.items():
      din = cfg.in_dims[k]
      cast16(self.dt, flat, cfg.d, din, din, w, w.shape[1], w.shape[1],
             src_off=L.off("video_dim_reduce.%s.fc.weight" % cfg.mods[k]))

  def reduce_weight(self, cfg, k):
    """(tensor, element offset, row pitch) of expert k's ReduceDim weight copy."""
    if k in self.red:
      return self.red[k], 0, self.red[k].shape[1]
    return self.flat16, cfg.layout.off("video_dim_reduce.%s.fc.weight" % cfg.mods[k]), cfg.in_dims[k]


def _e16(shape, like, dt):
  return torch.empty(shape, device=like.device, dtype=_lib.torch_dtype(dt))


class EncSpec:
  """Static description of a post-LN BERT encoder stack (reference model/bert.py:136-256; the same algebra as
  transformers' BertLayer) whose parameters live in a flat buffer: the video encoder and the text encoder."""

  def __init__(self, layout, prefix, ln_name, d, ff, H, L, eps, site_base):
    self.layout, self.prefix, self.ln = layout, prefix, ln_name
    self.d, self.ff, self.H, self.L, self.eps, self.site_base = d, ff, H, L, eps, site_base
    self.dh = d // H

  def off(self, l, name):
    return self.layout.off((self.prefix % l) + name)


def layers_forward(E, flat, f16, dt, h, h16, mask, B, S, p_hid, p_att, seed, ctr):
  """L encoder layers on 16-bit operands.  h (fp32 residual stream) / h16 (its 16-bit copy) [B*S, d]; mask [B*S]
  (1 = attend).  Returns (h_last fp32, per-layer saved activations)."""
  lib = _lib.load()
  st = stream_ptr()
  d, ff, H, dh = E.d, E.ff, E.H, E.dh
  BS = B * S
  scale = 1.0 / math.sqrt(dh)
  layers = []
  for l in range(E.L):
    site = E.site_base + 4 * l
    ls = Saved()
    ls.h16 = h16
    # K4: fused QKV projection -> 16-bit only (bert.py:137-143)
    qkv16 = _e16((BS, 3 * d), flat, dt)
    gemm16(dt, BS, 3 * d, d, h16, d, 0, f16, d, 0, b_off=E.off(l, "attention.self.query.weight"),
           bias=flat, bias_off=E.off(l, "attention.self.query.bias"), C16=qkv16, c16_ld=3 * d)
    # K5: fused attention (bert.py:147-170)
    ctx16 = _e16((BS, d), flat, dt)
    lse = None
    if dh == 128:
      lse = _empty((B, H, S), flat)
      check(lib.mmt_attention16_fwd(ptr(qkv16), ptr(mask), B, H, S, dh, scale, p_att, seed, ctr, site, ptr(ctx16),
                                    ptr(lse), dt, st), "mmt_attention16_fwd")
    else:
      check(lib.mmt_txt_attention_fwd(ptr(qkv16), ptr(mask), B, H, S, dh, scale, p_att, seed, ctr, site, ptr(ctx16), dt,
                                      st), "mmt_txt_attention_fwd")
    # K6: attention output dense + dropout + residual in the GEMM epilogue, then LayerNorm (bert.py:186-188)
    z1 = _empty((BS, d), flat)
    gemm16(dt, BS, d, d, ctx16, d, 0, f16, d, 0, b_off=E.off(l, "attention.output.dense.weight"),
           bias=flat, bias_off=E.off(l, "attention.output.dense.bias"), p_drop=p_hid, seed=seed, seed_ctr=ctr,
           site=site + 1, add=h, add_ld=d, C32=z1, c32_ld=d)
    a = _empty((BS, d), flat)
    a16 = _e16((BS, d), flat, dt)
    ls.mean1, ls.rstd1 = _empty((BS,), flat), _empty((BS,), flat)
    check(lib.mmt_ln16_fwd(ptr(z1), ptr(flat, E.off(l, "attention.output.%s.weight" % E.ln)),
                           ptr(flat, E.off(l, "attention.output.%s.bias" % E.ln)), BS, d, E.eps, ptr(a), ptr(a16),
                           ptr(ls.mean1), ptr(ls.rstd1), dt, st), "mmt_ln16_fwd")
    # K7: FFN up + erf-GELU (bert.py:218-219, 53): f16 = activation, g16 = gelu'(pre-activation) for the backward
    g16, fa16 = _e16((BS, ff), flat, dt), _e16((BS, ff), flat, dt)
    gemm16(dt, BS, ff, d, a16, d, 0, f16, d, 0, b_off=E.off(l, "intermediate.dense.weight"),
           bias=flat, bias_off=E.off(l, "intermediate.dense.bias"), epilogue=EPI_GELU, aux16=g16, aux_ld=ff,
           C16=fa16, c16_ld=ff)
    # K8: FFN down + dropout + residual, LayerNorm (bert.py:234-236)
    z2 = _empty((BS, d), flat)
    gemm16(dt, BS, d, ff, fa16, ff, 0, f16, ff, 0, b_off=E.off(l, "output.dense.weight"),
           bias=flat, bias_off=E.off(l, "output.dense.bias"), p_drop=p_hid, seed=seed, seed_ctr=ctr,
           site=site + 2, add=a, add_ld=d, C32=z2, c32_ld=d)
    hn = _empty((BS, d), flat)
    hn16 = _e16((BS, d), flat, dt) if l + 1 < E.L else None
    ls.mean2, ls.rstd2 = _empty((BS,), flat), _empty((BS,), flat)
    check(lib.mmt_ln16_fwd(ptr(z2), ptr(flat, E.off(l, "output.%s.weight" % E.ln)),
                           ptr(flat, E.off(l, "output.%s.bias" % E.ln)), BS, d, E.eps, ptr(hn), ptr(hn16),
                           ptr(ls.mean2), ptr(ls.rstd2), dt, st), "mmt_ln16_fwd")
    ls.qkv16, ls.ctx16, ls.lse, ls.z1, ls.a16, ls.u16, ls.f16, ls.z2 = qkv16, ctx16, lse, z1, a16, g16, fa16, z2
    layers.append(ls)
    h, h16 = hn, hn16
  return h, layers


def layers_backward(E, flat, f16, gflat, dt, layers, mask, B, S, dh_, p_hid, p_att, seed, ctr, sg, ws, on_layer_done=None):
  """Backward of layers_forward: parameter gradients into gflat (weights overwritten, small vectors accumulated),
  returns d loss / d (encoder input) [B*S, d] fp32.  `ws`: any object that may carry the zeroed dq32 workspace."""
  lib = _lib.load()
  st = stream_ptr()
  d, ff, H, dh = E.d, E.ff, E.H, E.dh
  BS = B * S
  scale = 1.0 / math.sqrt(dh)
  inv = 1.0 / sg
  dq32 = None
  if dh == 128:
    dq32 = ws.__dict__.get("_dq32")
    if dq32 is None or dq32.numel() != BS * d or dq32.device != flat.device:
      dq32 = torch.zeros((BS, d), device=flat.device, dtype=torch.float32)     # kept zeroed by the attention backward
      ws._dq32 = dq32
  for l in reversed(range(E.L)):
    site = E.site_base + 4 * l
    ls = layers[l]
    # --- LN2 backward: dz2 (fp32, to the residual a), dt2 (16-bit, to the FFN-down GEMMs), bias gradient ---
    dz2 = _empty((BS, d), flat)
    dt2 = _e16((BS, d), flat, dt)
    check(lib.mmt_ln16_bwd(ptr(dh_), None, ptr(ls.z2), ptr(ls.mean2), ptr(ls.rstd2),
                           ptr(flat, E.off(l, "output.%s.weight" % E.ln)), BS, d, p_hid, seed, ctr, site + 2, ptr(dz2),
                           ptr(dt2), sg, ptr(gflat, E.off(l, "output.%s.weight" % E.ln)),
                           ptr(gflat, E.off(l, "output.%s.bias" % E.ln)), ptr(gflat, E.off(l, "output.dense.bias")), dt, st),
          "mmt_ln16_bwd")
    # FFN down: dW2 [d, ff] = dt2^T @ f ; du = (dt2 @ W2) * gelu'(u) (+ its column sums = FFN-up bias gradient)
    gemm16(dt, d, ff, BS, dt2, d, 1, ls.f16, ff, 1, alpha=inv, split_k=True, C32=gflat,
           c32_off=E.off(l, "output.dense.weight"), c32_ld=ff)
    du = _e16((BS, ff), flat, dt)
    gemm16(dt, BS, ff, d, dt2, d, 0, f16, ff, 1, b_off=E.off(l, "output.dense.weight"), epilogue=EPI_DGELU,
           aux16=ls.u16, aux_ld=ff, C16=du, c16_ld=ff, colsum=gflat,
           colsum_off=E.off(l, "intermediate.dense.bias"), colsum_scale=inv)
    # FFN up: dW1 [ff, d] = du^T @ a ; da = du @ W1
    gemm16(dt, ff, d, BS, du, ff, 1, ls.a16, d, 1, alpha=inv, split_k=True, C32=gflat,
           c32_off=E.off(l, "intermediate.dense.weight"), c32_ld=d)
    da = _empty((BS, d), flat)
    gemm16(dt, BS, d, ff, du, ff, 0, f16, d, 1, b_off=E.off(l, "intermediate.dense.weight"), alpha=inv,
           C32=da, c32_ld=d)
    # --- LN1 backward on (da + dz2) ---
    dz1 = _empty((BS, d), flat)
    dt1 = _e16((BS, d), flat, dt)
    check(lib.mmt_ln16_bwd(ptr(da), ptr(dz2), ptr(ls.z1), ptr(ls.mean1), ptr(ls.rstd1),
                           ptr(flat, E.off(l, "attention.output.%s.weight" % E.ln)), BS, d, p_hid, seed, ctr, site + 1,
                           ptr(dz1), ptr(dt1), sg, ptr(gflat, E.off(l, "attention.output.%s.weight" % E.ln)),
                           ptr(gflat, E.off(l, "attention.output.%s.bias" % E.ln)),
                           ptr(gflat, E.off(l, "attention.output.dense.bias")), dt, st), "mmt_ln16_bwd")
    # attention output dense: dWo = dt1^T @ ctx ; dctx = dt1 @ Wo (16-bit, scale16 domain)
    gemm16(dt, d, d, BS, dt1, d, 1, ls.ctx16, d, 1, alpha=inv, split_k=True, C32=gflat,
           c32_off=E.off(l, "attention.output.dense.weight"), c32_ld=d)
    dctx = _e16((BS, d), flat, dt)
    gemm16(dt, BS, d, d, dt1, d, 0, f16, d, 1, b_off=E.off(l, "attention.output.dense.weight"),
           C16=dctx, c16_ld=d)
    # --- fused attention backward: dqkv16 (scale16 domain) and the QKV bias gradient ---
    dqkv = _e16((BS, 3 * d), flat, dt)
    if dh == 128:
      delta = _empty((B, H, S), flat)
      check(lib.mmt_attention16_bwd(ptr(ls.qkv16), ptr(ls.ctx16), ptr(dctx), ptr(ls.lse), ptr(mask), B, H, S, dh,
                                    scale, p_att, seed, ctr, site, sg, ptr(dqkv), ptr(dq32), ptr(delta),
                                    ptr(gflat, E.off(l, "attention.self.query.bias")), dt, st), "mmt_attention16_bwd")
    else:
      check(lib.mmt_txt_attention_bwd(ptr(ls.qkv16), ptr(dctx), ptr(mask), B, H, S, dh, scale, p_att, seed, ctr, site,
                                      ptr(dqkv), dt, st), "mmt_txt_attention_bwd")
      check(lib.mmt_colsum16(ptr(dqkv), BS, 3 * d, 3 * d, inv, ptr(gflat, E.off(l, "attention.self.query.bias")), dt, st),
            "mmt_colsum16")
    # QKV projection: dWqkv [3d, d] = dqkv^T @ h_in ; dh = dz1 + dqkv @ Wqkv
    gemm16(dt, 3 * d, d, BS, dqkv, 3 * d, 1, ls.h16, d, 1, alpha=inv, split_k=True, C32=gflat,
           c32_off=E.off(l, "attention.self.query.weight"), c32_ld=d)
    dh_ = _empty((BS, d), flat)
    gemm16(dt, BS, d, 3 * d, dqkv, 3 * d, 0, f16, d, 1, b_off=E.off(l, "attention.self.query.weight"), alpha=inv,
           add=dz1, add_ld=d, C32=dh_, c32_ld=d)
    if on_layer_done is not None:
      on_layer_done(l)
  return dh_


def video_forward(cfg, flat, feats, maxp, ft, ind, training, seed):
  """ReduceDim -> token assembly -> BertModel -> AGG read-out on 16-bit operands (engine.video_forward)."""
  L = cfg.layout
  W = cfg.w16
  dt = W.dt
  f16 = W.flat16
  d, ff, H, dh, M = cfg.d, cfg.ff, cfg.H, cfg.dh, cfg.M
  B, T = feats[0].shape[0], feats[0].shape[1]
  S = 1 + M * (T + 1)
  BS = B * S
  lib = _lib.load()
  st = stream_ptr()
  ctr = cfg.seed_ctr
  p_hid = cfg.p_hidden if training else 0.0
  p_att = cfg.p_attn if training else 0.0
  sv = Saved()
  sv.B, sv.T, sv.S, sv.training, sv.seed = B, T, S, training, seed
  sv.p_hid, sv.p_att = p_hid, p_att

  # ---- K1: ReduceDim (model.py:426-437, 723-726): one packing launch, then one GEMM per expert ----
  R1 = B * (T + 1)
  pd = _lib.PackDesc()
  sv.xpack = []
  for k in range(M):
    ld = _pad8(cfg.in_dims[k])
    xp = _e16((B, T + 1, ld), flat, dt)
    sv.xpack.append(xp)
    pd.feats[k], pd.maxp[k], pd.out[k] = ptr(feats[k]), ptr(maxp[k]), ptr(xp)
    pd.in_[k], pd.ld[k] = cfg.in_dims[k], ld
  pd.n, pd.B, pd.T, pd.dtype = M, B, T, dt
  check(lib.mmt_pack_inputs16(ctypes.byref(pd), st), "mmt_pack_inputs16")
  proj = _empty((M, R1, d), flat)
  # The M projections are independent and each fills only 16 of the 74 CTA pairs (1984 x 512 outputs): they are issued
  # round-robin on a few side streams and joined before the embedding kernel (7 x ~15 us back to back otherwise).
  main = torch.cuda.current_stream() if flat.is_cuda else None
  side = _side_streams(flat.device, min(4, M)) if (flat.is_cuda and M > 1) else []
  if side:
    fork = torch.cuda.Event()
    fork.record(main)
  for k, mod in enumerate(cfg.mods):
    w, w_off, w_ld = W.reduce_weight(cfg, k)
    ld = sv.xpack[k].shape[2]

    def project():
      gemm16(dt, R1, d, ld, sv.xpack[k], ld, 0, w, w_ld, 0, b_off=w_off, bias=flat,
             bias_off=L.off("video_dim_reduce.%s.fc.bias" % mod), C32=proj, c32_off=k * R1 * d, c32_ld=d)

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
  h = _empty((BS, d), flat)
  h16 = _e16((BS, d), flat, dt)
  sv.mask = _empty((BS,), flat)
  sv.pos_ids = _empty((BS,), flat, torch.int32)
  sv.type_ids = _empty((BS,), flat, torch.int32)
  sv.inv_norm = _empty((BS,), flat)
  sv.mean0, sv.rstd0 = _empty((BS,), flat), _empty((BS,), flat)
  e = "vid_bert.embeddings."
  check(lib.mmt_embed_ln16_fwd(
      ptr(proj), ptr(ft), ptr(ind), ptr(cfg.type_idx_dev), ptr(flat, L.off(e + "position_embeddings.weight")),
      ptr(flat, L.off(e + "token_type_embeddings.weight")), ptr(flat, L.off(e + "layer_norm.weight")),
      ptr(flat, L.off(e + "layer_norm.bias")), B, M, T, d, cfg.max_pos, cfg.eps, p_hid, seed, ctr,
      SITE_EMBED, ptr(h), ptr(h16), ptr(sv.mask), ptr(sv.pos_ids), ptr(sv.type_ids), ptr(sv.inv_norm),
      ptr(sv.mean0), ptr(sv.rstd0), dt, st), "mmt_embed_ln16_fwd")

  # ---- encoder layers (bert.py:249-256) ----
  h, sv.layers = layers_forward(cfg.enc_spec(), flat, f16, dt, h, h16, sv.mask, B, S, p_hid, p_att, seed, ctr)

  # ---- K10: expert read-out + L2 norm (model.py:583-587, 621-623) ----
  vid = _empty((B, M, d), flat)
  sv.vinv = _empty((B * M,), flat)
  check(lib.mmt_readout_norm_fwd(ptr(h), B, S, M, T, d, ptr(vid), ptr(sv.vinv), st), "mmt_readout_norm_fwd")
  sv.vid = vid
  return vid, sv


def head_forward(cfg, flat, bufs, text, training, seed):
  """Text head forward on 16-bit operands (engine.head_forward)."""
  L = cfg.layout
  W = cfg.w16
  dt, f16 = W.dt, W.flat16
  d, M = cfg.d, cfg.M
  R = text.shape[0]
  st = stream_ptr()
  lib = _lib.load()
  ctr = cfg.seed_ctr
  p_txt = cfg.p_txt if training else 0.0
  sv = Saved()
  sv.R, sv.training, sv.seed, sv.p_txt = R, training, seed, p_txt
  td = cfg.text_dim
  m0 = cfg.mods[0]

  # ---- K11: text GatedEmbeddingUnits, all experts at once (model.py:413-417, 697-702, 745-750) ----
  text16 = _e16((R, td), flat, dt)
  cast16(dt, text, R, td, td, text16, td, td)
  X = _empty((R, M * d), flat)
  X16 = _e16((R, M * d), flat, dt)
  gemm16(dt, R, M * d, td, text16, td, 0, f16, td, 0, b_off=L.off("text_GU.%s.fc.weight" % m0),
         bias=flat, bias_off=L.off("text_GU.%s.fc.bias" % m0), C32=X, c32_ld=M * d, C16=X16, c16_ld=M * d)
  G = _empty((R, M * d), flat)
  gemm16(dt, R, d, d, X16, M * d, 0, f16, d, 0, b_off=L.off("text_GU.%s.cg.fc.weight" % m0),
         bias=flat, bias_off=L.off("text_GU.%s.cg.fc.bias" % m0), bias_bs=d, batch=M, a_bs=(d, 0),
         b_bs=(d * d, 0), c_bs=(d, 0), C32=G, c32_ld=M * d)
  txt = _empty((R, M, d), flat)
  sv.Y = _empty((R, M * d), flat)
  sv.bn_mean, sv.bn_rstd = _empty((M * d,), flat), _empty((M * d,), flat)
  sv.n1, sv.n2 = _empty((R * M,), flat), _empty((R * M,), flat)
  check(lib.mmt_geu_gate_fwd(
      ptr(X), ptr(G), ptr(flat, L.off("text_GU.%s.cg.batch_norm.weight" % m0)),
      ptr(flat, L.off("text_GU.%s.cg.batch_norm.bias" % m0)), ptr(bufs), ptr(bufs, M * d), R, M, d,
      1 if training else 0, 0.1, 1e-5, ptr(txt), ptr(sv.Y), ptr(sv.bn_mean), ptr(sv.bn_rstd),
      ptr(sv.n1), ptr(sv.n2), st), "mmt_geu_gate_fwd")
  sv.X, sv.X16, sv.G, sv.txt, sv.text16 = X, X16, G, txt, text16

  # ---- K12: text mixture weights (model.py:273-281, 618) ----
  if p_txt > 0:
    tdrop16 = _e16((R, td), flat, dt)
    cast16(dt, text, R, td, td, tdrop16, td, td, p_drop=p_txt, seed=seed, seed_ctr=ctr, site=SITE_MOE_TXT)
  else:
    tdrop16 = text16
  Mp = _pad8(M)
  logits = _empty((R, Mp), flat)
  gemm16(dt, R, M, td, tdrop16, td, 0, f16, td, 0, b_off=L.off("moe_fc_txt.%s.weight" % m0),
         bias=flat, bias_off=L.off("moe_fc_txt.%s.bias" % m0), C32=logits, c32_ld=Mp)
  tw = _empty((R, M), flat)
  check(lib.mmt_moe_softmax_fwd(ptr(logits), R, M, Mp, ptr(tw), st), "mmt_moe_softmax_fwd")
  sv.tdrop16, sv.tw = tdrop16, tw
  return txt, tw, sv


def head_backward(cfg, flat, gflat, sv, dtxt, dtw, need_dtext=True):
  """Backward of head_forward: text-head parameter gradients into `gflat`; returns d loss / d text or None."""
  L = cfg.layout
  W = cfg.w16
  dt, f16 = W.dt, W.flat16
  d, M = cfg.d, cfg.M
  R = sv.R
  td = cfg.text_dim
  st = stream_ptr()
  lib = _lib.load()
  seed, ctr = sv.seed, cfg.seed_ctr
  m0 = cfg.mods[0]
  sg = cfg.scale16
  inv = 1.0 / sg

  def colsum(X, rows, n, ld, out_off):
    check(lib.mmt_colsum(ptr(X), rows, n, ld, 0, 0, ptr(gflat, out_off), 1, st), "mmt_colsum")

  dtext = None
  if dtw is not None:
    Mp = _pad8(M)
    dlog = _empty((R, Mp), flat)
    check(lib.mmt_moe_softmax_bwd(ptr(dtw), ptr(sv.tw), R, M, Mp, ptr(dlog), st), "mmt_moe_softmax_bwd")
    dlog16 = _e16((R, Mp), flat, dt)
    cast16(dt, dlog, R, Mp, Mp, dlog16, Mp, Mp, scale=sg)
    # dW_moe [M, td] = dlog^T @ tdrop ; db = colsum(dlog)
    gemm16(dt, M, td, R, dlog16, Mp, 1, sv.tdrop16, td, 1, alpha=inv, C32=gflat,
           c32_off=L.off("moe_fc_txt.%s.weight" % m0), c32_ld=td)
    colsum(dlog, R, M, Mp, L.off("moe_fc_txt.%s.bias" % m0))
    if need_dtext:
      dtext = _empty((R, td), flat)
      gemm16(dt, R, td, M, dlog16, Mp, 0, f16, td, 1, b_off=L.off("moe_fc_txt.%s.weight" % m0), alpha=inv,
             p_drop=sv.p_txt, seed=seed, seed_ctr=ctr, site=SITE_MOE_TXT, C32=dtext, c32_ld=td)
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
    dG16 = _e16((R, M * d), flat, dt)
    cast16(dt, dG, R, M * d, M * d, dG16, M * d, M * d, scale=sg)
    # cg.fc: dW2_m [d,d] = dG_m^T @ X_m ; db2 = colsum(dG) ; dX += dG_m @ W2_m
    gemm16(dt, d, d, R, dG16, M * d, 1, sv.X16, M * d, 1, alpha=inv, batch=M, a_bs=(d, 0), b_bs=(d, 0),
           c_bs=(d * d, 0), C32=gflat, c32_off=L.off("text_GU.%s.cg.fc.weight" % m0), c32_ld=d)
    colsum(dG, R, M * d, M * d, L.off("text_GU.%s.cg.fc.bias" % m0))
    dX16 = _e16((R, M * d), flat, dt)
    gemm16(dt, R, d, d, dG16, M * d, 0, f16, d, 1, b_off=L.off("text_GU.%s.cg.fc.weight" % m0), alpha=inv,
           batch=M, a_bs=(d, 0), b_bs=(d * d, 0), c_bs=(d, 0), add=dX, add_ld=M * d, C32=dX, c32_ld=M * d,
           C16=dX16, c16_ld=M * d, out16_scale=sg)
    # fc: dW1 [M*d, td] = dX^T @ text ; db1 = colsum(dX) ; dtext += dX @ W1
    gemm16(dt, M * d, td, R, dX16, M * d, 1, sv.text16, td, 1, alpha=inv, C32=gflat,
           c32_off=L.off("text_GU.%s.fc.weight" % m0), c32_ld=td)
    colsum(dX, R, M * d, M * d, L.off("text_GU.%s.fc.bias" % m0))
    if need_dtext:
      first = dtext is None
      if first:
        dtext = _empty((R, td), flat)
      gemm16(dt, R, td, M * d, dX16, M * d, 0, f16, td, 1, b_off=L.off("text_GU.%s.fc.weight" % m0), alpha=inv,
             add=None if first else dtext, add_ld=td, C32=dtext, c32_ld=td)
  return dtext


def video_backward(cfg, flat, gflat, sv, dvid, on_layer_done=None):
  """Backward of video_forward (engine.video_backward) on 16-bit operands."""
  L = cfg.layout
  W = cfg.w16
  dt, f16 = W.dt, W.flat16
  d, ff, H, dh, M = cfg.d, cfg.ff, cfg.H, cfg.dh, cfg.M
  B, T, S = sv.B, sv.T, sv.S
  BS = B * S
  st = stream_ptr()
  lib = _lib.load()
  seed, ctr = sv.seed, cfg.seed_ctr
  scale = 1.0 / math.sqrt(dh)
  sg = cfg.scale16
  inv = 1.0 / sg
  p_hid, p_att = sv.p_hid, sv.p_att

  dh_ = _empty((BS, d), flat)
  check(lib.mmt_readout_norm_bwd(ptr(dvid), ptr(sv.vid), ptr(sv.vinv), B, S, M, T, d, ptr(dh_), st),
        "mmt_readout_norm_bwd")
  dh_ = layers_backward(cfg.enc_spec(), flat, f16, gflat, dt, sv.layers, sv.mask, B, S, dh_, p_hid, p_att, seed, ctr,
                        sg, cfg, on_layer_done)

  # --- embeddings + token assembly backward ---
  R1 = B * (T + 1)
  dproj = _empty((M, R1, d), flat)
  dproj16 = _e16((M, R1, d), flat, dt)
  e = "vid_bert.embeddings."
  check(lib.mmt_embed_ln16_bwd(
      ptr(dh_), ptr(sv.proj), ptr(sv.pos_ids), ptr(sv.type_ids), ptr(sv.inv_norm), ptr(sv.mean0),
      ptr(sv.rstd0), ptr(flat, L.off(e + "position_embeddings.weight")),
      ptr(flat, L.off(e + "token_type_embeddings.weight")), ptr(flat, L.off(e + "layer_norm.weight")),
      B, M, T, d, p_hid, seed, ctr, SITE_EMBED, ptr(dproj), ptr(dproj16), sg,
      ptr(gflat, L.off(e + "position_embeddings.weight")),
      ptr(gflat, L.off(e + "token_type_embeddings.weight")), ptr(gflat, L.off(e + "layer_norm.weight")),
      ptr(gflat, L.off(e + "layer_norm.bias")), dt, st), "mmt_embed_ln16_bwd")
  # --- ReduceDim weight gradients (inputs carry no gradient): dW [d, in] = dproj_k^T @ xpack_k ---
  for k, mod in enumerate(cfg.mods):
    din = cfg.in_dims[k]
    ld = sv.xpack[k].shape[2]
    gemm16(dt, d, din, R1, dproj16, d, 1, sv.xpack[k], ld, 1, a_off=k * R1 * d, alpha=inv, split_k=True,
           C32=gflat, c32_off=L.off("video_dim_reduce.%s.fc.weight" % mod), c32_ld=din)
    check(lib.mmt_colsum(ptr(dproj, k * R1 * d), R1, d, d, 0, 0,
                         ptr(gflat, L.off("video_dim_reduce.%s.fc.bias" % mod)), 1, st), "mmt_colsum")


def sims_dots_split(dt, vid, txt):
  """Per-expert dot products dots[m, i, j] = <txt[i, m], vid[j, m]> as three tensor-core passes over two-term splits
  (x = hi + lo / 2048): fp32-class accuracy (~2^-21 relative).  TRAINING ONLY and only for large batches (the
  data-parallel global batch): at N = 512 the fp32-FMA kernel takes 143 us per step on every rank.  Evaluation keeps
  the fp32-FMA products (engine.sims_forward), whose ranking is bit-exact against the reference's."""
  Nv, M, d = vid.shape
  Nq = txt.shape[0]
  v_h, v_l = _e16((Nv, M * d), vid, dt), _e16((Nv, M * d), vid, dt)
  t_h, t_l = _e16((Nq, M * d), vid, dt), _e16((Nq, M * d), vid, dt)
  cast16(dt, vid, Nv, M * d, M * d, v_h, M * d, M * d, out_lo=v_l)
  cast16(dt, txt, Nq, M * d, M * d, t_h, M * d, M * d, out_lo=t_l)
  dots = torch.empty((M, Nq, Nv), device=vid.device, dtype=torch.float32)
  kw = dict(batch=M, a_bs=(d, 0), b_bs=(d, 0), c_bs=(Nq * Nv, 0), c32_ld=Nv)
  lo = 1.0 / 2048.0
  for i, (a, b, al) in enumerate(((t_h, v_h, 1.0), (t_h, v_l, lo), (t_l, v_h, lo))):
    gemm16(dt, Nq, Nv, d, a, M * d, 0, b, M * d, 0, alpha=al, C32=dots, add=dots if i else None, add_ld=Nv, **kw)
  return dots


def sims_backward_products(dt, ddots, vid, txt, scale16):
  """The two gradient products of the similarity (engine.sims_backward):
  dtxt[:, m, :] = ddots_m @ vid_m ; dvid[:, m, :] = ddots_m^T @ txt_m.
  They are the ROOT of every video- and text-side gradient, and tiny (2 x 2*N^2*M*d FLOP), so both operands are
  split into two 16-bit terms (x = hi + lo / 2048) and each product is three tensor-core passes
  Ah Bh + (Ah Bl + Al Bh) / 2048 -- fp32-class accuracy (~2^-21) instead of one 11-bit rounding that every
  downstream gradient would inherit."""
  Nv, M, d = vid.shape
  Nq = txt.shape[0]
  Nvp = _pad8(Nv)
  inv = 1.0 / scale16
  dd_h, dd_l = _e16((M, Nq, Nvp), vid, dt), _e16((M, Nq, Nvp), vid, dt)
  cast16(dt, ddots, M * Nq, Nv, Nv, dd_h, Nvp, Nvp, scale=scale16, out_lo=dd_l)
  v_h, v_l = _e16((Nv, M * d), vid, dt), _e16((Nv, M * d), vid, dt)
  t_h, t_l = _e16((Nq, M * d), vid, dt), _e16((Nq, M * d), vid, dt)
  cast16(dt, vid, Nv, M * d, M * d, v_h, M * d, M * d, out_lo=v_l)
  cast16(dt, txt, Nq, M * d, M * d, t_h, M * d, M * d, out_lo=t_l)
  dtxt = torch.empty_like(txt)
  dvid = torch.empty_like(vid)
  kw = dict(batch=M, a_bs=(Nq * Nvp, 0), b_bs=(d, 0), c_bs=(d, 0), c32_ld=M * d)
  lo = inv / 2048.0
  for i, (a, b, al) in enumerate(((dd_h, v_h, inv), (dd_h, v_l, lo), (dd_l, v_h, lo))):
    gemm16(dt, Nq, d, Nv, a, Nvp, 0, b, M * d, 1, alpha=al, C32=dtxt, add=dtxt if i else None, add_ld=M * d, **kw)
  for i, (a, b, al) in enumerate(((dd_h, t_h, inv), (dd_h, t_l, lo), (dd_l, t_h, lo))):
    gemm16(dt, Nv, d, Nq, a, Nvp, 1, b, M * d, 1, alpha=al, C32=dvid, add=dvid if i else None, add_ld=M * d, **kw)
  return dvid, dtxt
