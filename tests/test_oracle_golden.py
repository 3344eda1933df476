"""Pins the CPU oracle (oracle/mmt_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py -> tests/golden/*.npz).  CPU only."""
import collections
import json
import os

import numpy as np
import torch

from oracle import mmt_oracle as O


def _load(golden_dir, name):
  z = np.load(os.path.join(golden_dir, name))
  return {k: z[k] for k in z.files}


def _sub(fx, prefix):
  return collections.OrderedDict((k[len(prefix):], torch.from_numpy(v)) for k, v in fx.items()
                                 if k.startswith(prefix))


def _cfg(fx):
  c = json.loads(bytes(fx["cfg/json"]).decode())
  return c, {
      "expert_dims": O.compute_dims(c["modalities"], c["face_dim"]),
      "vid_bert_params": c["vid_bert_params"],
      "txt_dropout": 0.0,
      "test_caption_mode": "indep",
  }


def _batch(fx):
  b = {k: _sub(fx, "in/%s/" % k) for k in
       ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool")}
  return b


def test_cenet_train_forward_backward_matches_reference(golden_dir):
  fx = _load(golden_dir, "cenet_train.npz")
  c, cfg = _cfg(fx)
  P = _sub(fx, "P/")
  for k, v in P.items():
    if v.is_floating_point() and "running" not in k:
      v.requires_grad_(True)
  new_stats = {}
  out = O.cenet_forward(P, _batch(fx), cfg, training=True, out="conf", new_stats=new_stats,
                        text_feat=torch.from_numpy(fx["in/text_feat"]), return_intermediates=True)
  conf = out["cross_view_conf_matrix"]
  np.testing.assert_allclose(conf.detach().numpy(), fx["out/conf"], rtol=0, atol=2e-6)
  np.testing.assert_allclose(out["intermediates"]["last_layer"].detach().numpy(),
                             fx["act/layer1"], rtol=0, atol=2e-5)
  loss = O.max_margin_ranking_loss(conf, margin=c["margin"], fix_norm=True)
  np.testing.assert_allclose(loss.detach().numpy(), fx["out/loss"], rtol=1e-6, atol=1e-7)
  loss.backward()
  ref_grads = _sub(fx, "grad/")
  assert len(ref_grads) > 40
  gmax = max(float(g.abs().max()) for g in ref_grads.values())
  for name, g in ref_grads.items():
    got = P[name].grad
    assert got is not None, name
    # key.bias gradients are analytically zero (softmax shift invariance): pure rounding noise,
    # so the scale is floored at 1e-3 of the largest gradient.
    scale = max(float(g.abs().max()), 1e-3 * gmax)
    err = float((got - g).abs().max()) / scale
    assert err < 2e-4, (name, err)
  # parameters the reference leaves without gradient must get none from the oracle either
  for name, p in P.items():
    if p.requires_grad and name not in ref_grads:
      assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
  for name, v in _sub(fx, "P1/").items():
    if "num_batches" in name:
      continue
    np.testing.assert_allclose(new_stats[name].numpy(), v.numpy(), rtol=1e-5, atol=1e-6)


def test_token_assembly_matches_reference_embeddings(golden_dir):
  fx = _load(golden_dir, "cenet_train.npz")
  c, cfg = _cfg(fx)
  P = _sub(fx, "P/")
  out = O.cenet_forward(P, _batch(fx), cfg, training=True, out="conf",
                        text_feat=torch.from_numpy(fx["in/text_feat"]), return_intermediates=True)
  it = out["intermediates"]
  vb = cfg["vid_bert_params"]
  emb = O.bert_embeddings(it["tokens"], it["type_ids"], it["pos_ids"], P, "vid_bert.embeddings.",
                          vb["layer_norm_eps"], 0.0, False)
  np.testing.assert_allclose(emb.numpy(), fx["act/embeddings"], rtol=0, atol=5e-6)
  assert int(it["pos_ids"].max()) == vb["max_position_embeddings"] - 1   # the clamp was exercised


def test_cenet_eval_embds_and_conf(golden_dir):
  fx = _load(golden_dir, "cenet_eval.npz")
  c, cfg = _cfg(_load(golden_dir, "cenet_train.npz"))
  P = _sub(_load(golden_dir, "cenet_train.npz"), "P/")
  tf = torch.from_numpy(fx["in/text_feat"])
  e = O.cenet_forward(P, _batch(fx), cfg, training=False, out="embds", text_feat=tf)
  for k in ("vid_embds", "text_embds", "vid_weights", "text_weights"):
    np.testing.assert_allclose(e[k].numpy(), fx["out/" + k], rtol=0, atol=3e-6, err_msg=k)
  cm = O.cenet_forward(P, _batch(fx), cfg, training=False, out="conf", text_feat=tf)
  np.testing.assert_allclose(cm["cross_view_conf_matrix"].numpy(), fx["out/conf"], rtol=0,
                             atol=3e-6)


def test_sims_and_loss_golden(golden_dir):
  fx = _load(golden_dir, "sims_loss.npz")
  vid, txt = _sub(fx, "vid/"), _sub(fx, "txt/")
  mods = list(vid.keys())
  vw, tw = torch.from_numpy(fx["vw"]), torch.from_numpy(fx["tw"])
  for merge in ("avg", "indep"):
    s = O.sharded_cross_view_inner_product(vid, txt, vw, tw, mods, merge)
    np.testing.assert_allclose(s.numpy(), fx["sims_" + merge], rtol=0, atol=1e-6)
    # ranking indices identical (stable argsort = index tie-break)
    a = np.argsort(-s.numpy(), axis=1, kind="stable")
    b = np.argsort(-fx["sims_" + merge], axis=1, kind="stable")
    assert (a == b).all()
  for margin, fix in ((0.05, True), (0.2, True), (0.05, False)):
    x = torch.tensor(fx["sims_avg"], requires_grad=True)
    l = O.max_margin_ranking_loss(x, margin, fix)
    np.testing.assert_allclose(l.detach().numpy(), fx["loss_m%g_fix%d" % (margin, fix)],
                               rtol=1e-6)
    l.backward()
    np.testing.assert_allclose(x.grad.numpy(), fx["dloss_m%g_fix%d" % (margin, fix)], rtol=1e-5,
                               atol=1e-9)


def test_loss_n1_is_nan_like_reference():
  # model/loss.py:38-65 with n=1 and fix_norm: mean over an empty selection -> nan
  assert torch.isnan(O.max_margin_ranking_loss(torch.ones(1, 1), 0.05, True))


def test_retrieval_ranks_golden(golden_dir):
  fx = _load(golden_dir, "metrics.npz")
  ranks = O.retrieval_ranks(fx["sims"], fx["query_masks"])
  # metric.py reports 1-based ranks: R@k = 100 * mean(rank < k) on 0-based, MedR = median + 1
  np.testing.assert_allclose(100.0 * np.mean(ranks == 0), fx["t2v/R1"], rtol=1e-9)
  np.testing.assert_allclose(100.0 * np.mean(ranks < 5), fx["t2v/R5"], rtol=1e-9)
  np.testing.assert_allclose(100.0 * np.mean(ranks < 10), fx["t2v/R10"], rtol=1e-9)
  np.testing.assert_allclose(np.median(ranks) + 1, fx["t2v/MedR"], rtol=1e-9)
  np.testing.assert_allclose(np.mean(ranks) + 1, fx["t2v/MeanR"], rtol=1e-9)


def test_retrieval_metrics_v2t_and_all_keys_golden(golden_dir):
  """Every scalar the reference's t2v_metrics / v2t_metrics returned for the fixture (ties, two masked
  captions): the oracle's rank restatements + cols2metrics reproduce them."""
  fx = _load(golden_dir, "metrics.npz")
  qm = fx["query_masks"]
  t2v = O.cols2metrics(O.retrieval_ranks(fx["sims"], qm), int(qm.sum()))
  v2t = O.cols2metrics(O.retrieval_ranks_v2t(fx["sims"], qm), fx["sims"].shape[1])
  for k in ("R1", "R5", "R10", "R50", "MedR", "MeanR", "geometric_mean_R1-R5-R10"):
    np.testing.assert_allclose(t2v[k], fx["t2v/" + k], rtol=1e-9, err_msg="t2v/" + k)
    np.testing.assert_allclose(v2t[k], fx["v2t/" + k], rtol=1e-9, err_msg="v2t/" + k)
