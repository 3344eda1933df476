"""Shared builders for the parity tests (CPU oracle vs CUDA path on identical inputs/weights)."""
import collections
import types

import torch

from oracle import mmt_oracle as O

VB_FULL = {"vocab_size_or_config_json_file": 10, "hidden_size": 512, "num_hidden_layers": 4,
           "num_attention_heads": 4, "intermediate_size": 3072, "hidden_act": "gelu",
           "hidden_dropout_prob": 0.0, "attention_probs_dropout_prob": 0.0,
           "max_position_embeddings": 32, "type_vocab_size": 19, "initializer_range": 0.02,
           "layer_norm_eps": 1e-12}


class TxtStub(torch.nn.Module):
  """Hot-path-only scope: the third-party text encoder is replaced on both
  sides by a module returning fixed features."""

  def __init__(self, hidden):
    super().__init__()
    self.hidden = hidden
    self.config = types.SimpleNamespace(hidden_size=hidden.shape[-1])

  def forward(self, input_ids, **kw):
    return (self.hidden,)


def make_case(modalities, B, T, layers=4, dropout=0.0, caps=1, seed=1234, dense=False,
              max_pos=32, type_vocab=19, face_dim=512, text_dim=768):
  ed = O.compute_dims(modalities, face_dim)
  vb = dict(VB_FULL, num_hidden_layers=layers, hidden_dropout_prob=dropout,
            attention_probs_dropout_prob=dropout, max_position_embeddings=max_pos,
            type_vocab_size=type_vocab)
  P = O.init_params(ed, vb, text_dim=text_dim, seed=seed + 1)
  batch = O.synth_batch(ed, B, T, caps=caps, seed=seed, dense=dense, text_dim=text_dim)
  cfg = {"expert_dims": ed, "vid_bert_params": vb, "txt_dropout": dropout,
         "test_caption_mode": "indep"}
  return ed, vb, P, batch, cfg


class TxtEmb(torch.nn.Module):
  """A tiny TRAINABLE stand-in for the text encoder: hidden states = an embedding of the token ids (CLS row =
  the embedding of the first token).  Exercises the gradient path into parameters outside the flat buffer."""

  def __init__(self, vocab=64, dim=768, seed=5):
    super().__init__()
    g = torch.Generator().manual_seed(seed)
    self.emb = torch.nn.Embedding(vocab, dim)
    with torch.no_grad():
      self.emb.weight.copy_(torch.randn(vocab, dim, generator=g))
    self.config = types.SimpleNamespace(hidden_size=dim)
    self.vocab = vocab

  def forward(self, input_ids, **kw):
    return (self.emb(input_ids % self.vocab),)


def build_cuda_net(ed, vb, P, batch, dropout=0.0, device="cuda", precision="fp32", txt_bert=None):
  from mmt_b200 import _lib
  from mmt_b200.model.model import CENet
  W = batch["token_ids"].shape[2]
  R = batch["text_feat"].shape[0]
  hidden = torch.zeros(R, W, batch["text_feat"].shape[1])
  hidden[:, 0] = batch["text_feat"]
  net = CENet(l2renorm=False, expert_dims=ed, tokenizer=None, keep_missing_modalities=True,
              test_caption_mode="indep", txt_inp="bertftn", txt_agg="bertftn", txt_wgh="emb",
              vid_wgh="none", vid_cont="bert", vid_inp="both", pos_enc="tint", out_tok="mxp",
              vid_bert_params=vb, txt_pro="gbn",
              txt_bert_params={"hidden_dropout_prob": dropout,
                               "attention_probs_dropout_prob": dropout},
              txt_bert=txt_bert if txt_bert is not None else TxtStub(hidden.to(device)))
  net.load_state_dict(P, strict=txt_bert is None)
  net.cfg.precision = {"fp32": _lib.PREC_FP32, "tf32": _lib.PREC_TF32, "f16": _lib.PREC_F16,
                       "bf16": _lib.PREC_BF16}[precision]
  return net.to(device)


def grad_errors(net, ref_grads):
  """Gradient parity of a CUDA net against oracle gradients {name: tensor}.
  Returns (whole-gradient max-norm relative error, whole-gradient rel-L2, worst per-tensor (name, err) with the
  per-tensor scale max(|g_ref|_max, 1e-2 * global max), worst per-tensor rel-L2 (name, err))."""
  gmax = max(float(g.abs().max()) for g in ref_grads.values())
  num = den = 0.0
  emax = 0.0
  worst, worst_l2 = ("", 0.0), ("", 0.0)
  for name, g in ref_grads.items():
    got = net._param(name).grad
    assert got is not None, name
    diff = got.detach().cpu().double() - g.double()
    emax = max(emax, float(diff.abs().max()))
    num += float((diff * diff).sum())
    den += float((g.double() * g.double()).sum())
    e = float(diff.abs().max()) / max(float(g.abs().max()), 1e-2 * gmax)
    if e > worst[1]:
      worst = (name, e)
    l2 = float(diff.norm()) / max(float(g.double().norm()), 1e-30)
    if l2 > worst_l2[1] and float(g.abs().max()) > 1e-2 * gmax:
      worst_l2 = (name, l2)
  return emax / gmax, (num / den) ** 0.5, worst, worst_l2


def oracle_step(P, batch, cfg):
  """Oracle forward + MaxMarginRankingLoss + backward (dropout off): conf, loss, {name: grad}."""
  Pr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
        for k, v in P.items()}
  ref = O.cenet_forward(Pr, batch, cfg, training=True, out="conf", text_feat=batch["text_feat"])
  conf = ref["cross_view_conf_matrix"]
  loss = O.max_margin_ranking_loss(conf, 0.05, True)
  loss.backward()
  grads = {k: v.grad for k, v in Pr.items() if getattr(v, "grad", None) is not None}
  return conf.detach(), float(loss.detach()), grads


def batch_kwargs(batch, device=None):
  kw = {}
  for k in ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool"):
    kw[k] = collections.OrderedDict(
        (m, v.clone().to(device) if device else v.clone()) for m, v in batch[k].items())
  kw["token_ids"] = batch["token_ids"].to(device) if device else batch["token_ids"]
  kw["query_masks"] = batch["query_masks"].to(device) if device else batch["query_masks"]
  return kw


def rel_err(got, ref):
  """max-norm relative error (1e-3 relative fp32, max-norm and rel-L2)."""
  got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
  scale = max(float(ref.abs().max()), 1e-30)
  return float((got - ref).abs().max()) / scale


def rel_l2(got, ref):
  got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
  return float((got - ref).norm()) / max(float(ref.norm()), 1e-30)
