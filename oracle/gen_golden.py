"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY; runs in the build container (the reference tree does not travel to the
GPU box).  Re-run with:  python oracle/gen_golden.py

The reference has no tests / golden vectors of its own, so the oracle
(oracle/mmt_oracle.py) is pinned against these outputs of the reference itself:

  cenet_train.npz   CENet built from configs_pub/eccv20/MSRVTT_jsfusion_trainval.json with the
                    width shrunk (hidden 64, 2 layers, 3 experts) so the fixture stays < 2 MB;
                    model.train(), dropout 0; inputs, state_dict, per-stage activations,
                    conf matrix, MaxMarginRankingLoss, every parameter gradient, BN running stats.
  cenet_eval.npz    same weights, model.eval(), 2 captions per video, out='embds' and 'conf'
                    (merge 'indep').
  sims_loss.npz     sharded_cross_view_inner_product + MaxMarginRankingLoss (+ fix_norm=False)
                    on random embeddings at the real width (d=512, M=7), incl. a zero-weight row.
  metrics.npz       model.metric.t2v_metrics / v2t_metrics on a random sims matrix with ties.
"""
import collections
import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402
from oracle import mmt_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def tiny_config():
  cfg = json.load(open(os.path.join(refshim.REFERENCE_ROOT,
                                    "configs_pub/eccv20/MSRVTT_jsfusion_trainval.json")))
  cfg["experts"]["modalities"] = ["s3d", "vggish", "ocr"]
  a = cfg["arch"]["args"]
  a["same_dim"] = 64
  a["vid_bert_params"].update(hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                              intermediate_size=128, hidden_dropout_prob=0.0,
                              attention_probs_dropout_prob=0.0)
  a["txt_bert_params"] = {"hidden_dropout_prob": 0.0, "attention_probs_dropout_prob": 0.0,
                          "hidden_size": 48, "num_hidden_layers": 1, "num_attention_heads": 2,
                          "intermediate_size": 64}
  return cfg


def to_np(d, prefix):
  return {prefix + k: v.detach().cpu().numpy() for k, v in d.items()}


def clone_batch(batch):
  kw = {}
  for k in ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool"):
    kw[k] = collections.OrderedDict((m, v.clone()) for m, v in batch[k].items())
  kw["token_ids"] = batch["token_ids"].clone()
  kw["query_masks"] = batch["query_masks"].clone()
  return kw


def main():
  os.makedirs(OUT, exist_ok=True)
  ref_model, ref_loss, ref_metric, ref_util = refshim.load_reference()
  cfg = tiny_config()
  expert_dims = ref_util.compute_dims(cfg)
  torch.manual_seed(0)
  model = ref_model.CENet(expert_dims=expert_dims, tokenizer=None, **cfg["arch"]["args"])
  # non-trivial LN/BN affine parameters and biases so the fixture exercises them
  g = torch.Generator().manual_seed(7)
  with torch.no_grad():
    for n, p in model.named_parameters():
      if n.startswith("txt_bert"):
        continue
      if "layer_norm" in n or "batch_norm" in n or n.endswith("bias"):
        p.add_(0.1 * torch.randn(p.shape, generator=g))
  loss_fn = ref_loss.MaxMarginRankingLoss(**cfg["loss"]["args"])

  B, T, W, TD = 6, 5, 8, 48
  batch = O.synth_batch(expert_dims, B, T, w=W, caps=1, seed=11, text_dim=TD)
  batch["features_t"]["s3d"][0, 0] = 77.0     # exercises the clamp to max_pos-1 (model.py:516)
  hidden = torch.zeros(B, W, TD)
  hidden[:, 0] = batch["text_feat"]
  model.txt_bert = refshim.TxtBertStub(hidden)

  state0 = {k: v.clone() for k, v in model.state_dict().items() if not k.startswith("txt_bert")}

  # ---- train-mode forward / loss / backward, with per-stage hooks
  acts = {}

  def hook(name):
    def f(mod, inp, out):
      acts[name] = (out[0] if isinstance(out, tuple) else out).detach().clone()
    return f

  hs = [model.vid_bert.embeddings.register_forward_hook(hook("embeddings"))]
  for i, l in enumerate(model.vid_bert.encoder.layer):
    hs.append(l.register_forward_hook(hook("layer%d" % i)))
    hs.append(l.attention.register_forward_hook(hook("layer%d_attn" % i)))
  model.train()
  out = model(**clone_batch(batch), out="conf", device=torch.device("cpu"))
  conf = out["cross_view_conf_matrix"]
  loss = loss_fn(conf)
  loss.backward()
  for h in hs:
    h.remove()
  grads = {n: p.grad for n, p in model.named_parameters()
           if p.grad is not None and not n.startswith("txt_bert")}
  state1 = {k: v.clone() for k, v in model.state_dict().items()
            if "batch_norm.running" in k or "num_batches" in k}

  fx = {}
  fx.update(to_np(state0, "P/"))
  for k in ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool"):
    fx.update(to_np(batch[k], "in/%s/" % k))
  fx["in/token_ids"] = batch["token_ids"].numpy()
  fx["in/query_masks"] = batch["query_masks"].numpy()
  fx["in/text_feat"] = batch["text_feat"].numpy()
  fx.update(to_np(acts, "act/"))
  fx["out/conf"] = conf.detach().numpy()
  fx["out/loss"] = loss.detach().numpy()
  fx.update(to_np(grads, "grad/"))
  fx.update(to_np(state1, "P1/"))
  fx["cfg/json"] = np.frombuffer(json.dumps({
      "modalities": list(expert_dims.keys()), "face_dim": cfg["experts"]["face_dim"],
      "vid_bert_params": cfg["arch"]["args"]["vid_bert_params"], "same_dim": 64,
      "margin": cfg["loss"]["args"]["margin"], "text_dim": TD, "B": B, "T": T, "W": W}).encode(),
      dtype=np.uint8)
  np.savez_compressed(os.path.join(OUT, "cenet_train.npz"), **fx)
  print("cenet_train: loss", float(loss), "arrays", len(fx))

  # ---- eval mode, 2 captions per video
  model.load_state_dict({**model.state_dict(), **state0})
  model.eval()
  caps = 2
  batch2 = O.synth_batch(expert_dims, B, T, w=W, caps=caps, seed=12, text_dim=TD)
  hidden = torch.zeros(B * caps, W, TD)
  hidden[:, 0] = batch2["text_feat"]
  model.txt_bert = refshim.TxtBertStub(hidden)
  with torch.no_grad():
    e = model(**clone_batch(batch2), out="embds", device=torch.device("cpu"))
    c = model(**clone_batch(batch2), out="conf", device=torch.device("cpu"))
  fx = {}
  for k in ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool"):
    fx.update(to_np(batch2[k], "in/%s/" % k))
  fx["in/token_ids"] = batch2["token_ids"].numpy()
  fx["in/text_feat"] = batch2["text_feat"].numpy()
  fx.update(to_np({k: v for k, v in e.items()}, "out/"))
  fx["out/conf"] = c["cross_view_conf_matrix"].numpy()
  np.savez_compressed(os.path.join(OUT, "cenet_eval.npz"), **fx)
  print("cenet_eval: conf", tuple(c["cross_view_conf_matrix"].shape))

  # ---- similarity + loss at the real width
  g = torch.Generator().manual_seed(3)
  mods = ["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"]
  n, caps, d = 12, 2, 512
  vid = collections.OrderedDict((m, torch.nn.functional.normalize(
      torch.randn(n, d, generator=g), dim=-1)) for m in mods)
  txt = collections.OrderedDict((m, torch.nn.functional.normalize(
      torch.randn(n, caps, d, generator=g), dim=-1)) for m in mods)
  vw = torch.rand(n, len(mods), generator=g)
  vw[3] = 0.0                                   # all-zero video weights -> 1e-5 path (model.py:816)
  vw = torch.nn.functional.normalize(vw, p=1, dim=-1)
  tw = torch.softmax(torch.randn(n, caps, len(mods), generator=g), -1)
  fx = {}
  fx.update(to_np(vid, "vid/"))
  fx.update(to_np(txt, "txt/"))
  fx["vw"], fx["tw"] = vw.numpy(), tw.numpy()
  for merge in ("avg", "indep"):
    s = ref_model.sharded_cross_view_inner_product(
        vid_embds=vid, text_embds=copy.deepcopy(txt), vid_weights=vw, text_weights=tw,
        subspaces=mods, merge_caption_similiarities=merge)
    fx["sims_" + merge] = s.numpy()
  x = torch.tensor(fx["sims_avg"], requires_grad=True)
  for margin, fix in ((0.05, True), (0.2, True), (0.05, False)):
    l = ref_loss.MaxMarginRankingLoss(margin=margin, fix_norm=fix)(x)
    gx, = torch.autograd.grad(l, x)
    fx["loss_m%g_fix%d" % (margin, fix)] = l.detach().numpy()
    fx["dloss_m%g_fix%d" % (margin, fix)] = gx.numpy()
  np.savez_compressed(os.path.join(OUT, "sims_loss.npz"), **fx)
  print("sims_loss: ok")

  # ---- metrics (model/metric.py) with ties
  rng = np.random.RandomState(5)
  nv, caps = 20, 3
  sims = np.round(rng.randn(nv * caps, nv), 1).astype(np.float32)   # rounding creates ties
  qm = np.ones((nv, caps), dtype=np.int32)
  qm[2, 1] = 0
  qm[7, 2] = 0
  t2v = ref_metric.t2v_metrics(sims, qm)
  v2t = ref_metric.v2t_metrics(sims, qm)
  fx = {"sims": sims, "query_masks": qm}
  for k, v in t2v.items():
    if np.isscalar(v) or isinstance(v, (float, int, np.floating)):
      fx["t2v/" + k] = np.asarray(v, dtype=np.float64)
  for k, v in v2t.items():
    if np.isscalar(v) or isinstance(v, (float, int, np.floating)):
      fx["v2t/" + k] = np.asarray(v, dtype=np.float64)
  np.savez_compressed(os.path.join(OUT, "metrics.npz"), **fx)
  print("metrics:", {k: float(v) for k, v in fx.items() if k.startswith("t2v/")})


if __name__ == "__main__":
  main()
