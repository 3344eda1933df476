"""Synthetic workloads of the MMT hot path: expert tables, random parameters with the reference's names / shapes,
and minibatches with the reference's collate contract.

NEUTRAL module: neither product code nor oracle.  bench.py (both arms), the tests and the oracle draw their inputs
from here, so that the GPU arm's process never imports oracle/.  Nothing here computes the model.
"""
import collections

import torch

# utils/util.py:154-247 (the experts the published configs use)
EXPERT_TABLE = {
    "s3d": (1024, 1), "vggish": (128, 2), "face": (None, 3), "audio": (128, 4), "rgb": (2048, 5),
    "speech": (300, 6), "ocr": (300, 7), "flow": (1024, 8), "scene": (2208, 9),
}


def compute_dims(modalities, face_dim=512):
  """utils/util.py:154-247: sorted expert names -> {'dim', 'idx'}."""
  dims = collections.OrderedDict()
  for mod in sorted(modalities):
    in_dim, idx = EXPERT_TABLE[mod]
    dims[mod] = {"dim": face_dim if in_dim is None else in_dim, "idx": idx}
  return dims


def init_params(expert_dims, vid_bert_params, text_dim=768, same_dim=512, seed=0,
                dtype=torch.float32):
  """Random parameters with the reference's names/shapes and init rules:
  BERT Linear/Embedding N(0, initializer_range), zero bias, LN 1/0 (bert.py:361-369); other
  Linear layers use a small N(0, 0.02) too (any fixed init works for parity -- the same tensors
  are fed to both sides)."""
  g = torch.Generator().manual_seed(seed)
  std = vid_bert_params.get("initializer_range", 0.02)
  d, ff = vid_bert_params["hidden_size"], vid_bert_params["intermediate_size"]
  assert d == same_dim
  P = collections.OrderedDict()

  def lin(name, o, i, bias_std=0.0):
    P[name + ".weight"] = torch.randn(o, i, generator=g, dtype=dtype) * std
    P[name + ".bias"] = torch.randn(o, generator=g, dtype=dtype) * bias_std

  def ln(name):
    P[name + ".weight"] = 1.0 + 0.1 * torch.randn(d, generator=g, dtype=dtype)
    P[name + ".bias"] = 0.1 * torch.randn(d, generator=g, dtype=dtype)

  for mod, v in expert_dims.items():
    lin("video_dim_reduce.%s.fc" % mod, d, v["dim"], 0.02)
  P["vid_bert.embeddings.position_embeddings.weight"] = \
      torch.randn(vid_bert_params["max_position_embeddings"], d, generator=g, dtype=dtype) * std
  P["vid_bert.embeddings.token_type_embeddings.weight"] = \
      torch.randn(vid_bert_params["type_vocab_size"], d, generator=g, dtype=dtype) * std
  ln("vid_bert.embeddings.layer_norm")
  for l in range(vid_bert_params["num_hidden_layers"]):
    pre = "vid_bert.encoder.layer.%d." % l
    for n in ("query", "key", "value"):
      lin(pre + "attention.self." + n, d, d, 0.02)
    lin(pre + "attention.output.dense", d, d, 0.02)
    ln(pre + "attention.output.layer_norm")
    lin(pre + "intermediate.dense", ff, d, 0.02)
    lin(pre + "output.dense", d, ff, 0.02)
    ln(pre + "output.layer_norm")
  lin("vid_bert.pooler.dense", d, d)
  for mod in expert_dims:
    pre = "text_GU.%s." % mod
    lin(pre + "fc", d, text_dim, 0.02)
    lin(pre + "cg.fc", d, d, 0.02)
    P[pre + "cg.batch_norm.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g, dtype=dtype)
    P[pre + "cg.batch_norm.bias"] = 0.1 * torch.randn(d, generator=g, dtype=dtype)
    P[pre + "cg.batch_norm.running_mean"] = torch.zeros(d, dtype=dtype)
    P[pre + "cg.batch_norm.running_var"] = torch.ones(d, dtype=dtype)
    P[pre + "cg.batch_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
  for mod in expert_dims:
    lin("moe_fc_txt.%s" % mod, 1, text_dim, 0.02)
  return P


def synth_batch(expert_dims, b, t, w=30, caps=1, seed=1234, dense=False, text_dim=768,
                dtype=torch.float32):
  """Synthetic minibatch with the collate contract of SURVEY.md §8(b)/(d):
  features[m] ~ N(0,1) [B,T,in]; features_ind valid-first then padding (k ~ U{0..T}; k=0 with
  probability .3 for ocr/speech/face); features_t = 2+t for valid, 1 for padding
  (base_dataset.py:95, 779-781); max/avg pool over the valid rows, zeros when none
  (base_dataset.py:381-390, 800-822); padded rows are zero features."""
  g = torch.Generator().manual_seed(seed)
  batch = {k: collections.OrderedDict() for k in
           ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool")}
  for mod, v in expert_dims.items():
    x = torch.randn(b, t, v["dim"], generator=g, dtype=dtype)
    if dense:
      k = torch.full((b,), t)
    else:
      k = torch.randint(0, t + 1, (b,), generator=g)
      if mod in ("ocr", "speech", "face"):
        k = torch.where(torch.rand(b, generator=g) < 0.3, torch.zeros_like(k), k)
    ind = (torch.arange(t)[None, :] < k[:, None]).to(dtype)
    x = x * ind[:, :, None]
    ft = torch.where(ind > 0, 2.0 + torch.arange(t, dtype=dtype)[None, :],
                     torch.ones(b, t, dtype=dtype))
    neg = torch.where(ind[:, :, None] > 0, x, torch.full_like(x, -float("inf")))
    mx = neg.max(dim=1)[0]
    mx = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)
    av = x.sum(1) / k.clamp(min=1)[:, None].to(dtype)
    batch["features"][mod] = x
    batch["features_t"][mod] = ft
    batch["features_ind"][mod] = ind
    batch["features_maxpool"][mod] = mx
    batch["features_avgpool"][mod] = av
  tok = torch.zeros(b, caps, w, 2, dtype=torch.int32)
  ln_ = torch.randint(5, w + 1, (b, caps), generator=g)
  ids = torch.randint(1000, 20000, (b, caps, w), generator=g, dtype=torch.int32)
  valid = (torch.arange(w)[None, None, :] < ln_[:, :, None])
  ids[:, :, 0] = 101
  ids.scatter_(2, (ln_ - 1).unsqueeze(-1), torch.full((b, caps, 1), 102, dtype=torch.int32))
  tok[..., 0] = ids * valid
  tok[..., 1] = valid.to(torch.int32)
  batch["token_ids"] = tok
  batch["query_masks"] = torch.ones(b, caps, dtype=torch.int32)
  batch["text_feat"] = torch.randn(b * caps, text_dim, generator=g, dtype=dtype)
  return batch
