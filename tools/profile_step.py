"""Workload for the ncu captures under profiles/: a few f16 train steps of the benchmark configuration (C2, B = 64,
dropout 0.1, FusedAdam), then the stand-alone MaxMarginRankingLoss forward at N = 16384.  Run under
`ncu -k regex:<kernel> ...`; never a source of timing numbers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import workloads as W  # noqa: E402
from mmt_b200 import _lib, engine  # noqa: E402
from mmt_b200.model.loss import MaxMarginRankingLoss  # noqa: E402
from mmt_b200.model.model import CENet  # noqa: E402
from mmt_b200.optim import FusedAdam  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
w = bench.WORKLOADS["C2"]
B = w["B"]
ed, batches = bench.make_batches(w, B, 1, 1234)
P = W.init_params(ed, bench.vb_params(w), seed=0)
feed = bench.TextFeed()
net = CENet(l2renorm=False, expert_dims=ed, tokenizer=None, keep_missing_modalities=True, test_caption_mode="indep",
            txt_inp="bertftn", txt_agg="bertftn", txt_wgh="emb", vid_wgh="none", vid_cont="bert", vid_inp="both",
            pos_enc="tint", out_tok="mxp", vid_bert_params=bench.vb_params(w), txt_pro="gbn",
            txt_bert_params={"hidden_dropout_prob": 0.1, "attention_probs_dropout_prob": 0.1}, txt_bert=feed)
net.load_state_dict(P, strict=True)
net.to(dev).train()
net.cfg.precision = _lib.PREC_F16
opt = FusedAdam(net, lr=5e-5)
crit = MaxMarginRankingLoss(0.05, True)
b = batches[0]
kw = {k: {m: v.to(dev) for m, v in b[k].items()} for k in ("features", "features_t", "features_ind", "features_avgpool",
                                                          "features_maxpool")}
kw["token_ids"], kw["query_masks"] = b["token_ids"].to(dev), b["query_masks"]
feed.cls = b["text_feat"].to(dev)
for _ in range(steps):
  opt.zero_grad()
  loss = crit(net(**kw, out="conf", device=dev)["cross_view_conf_matrix"])
  loss.backward()
  opt.step()
torch.cuda.synchronize()
x = torch.rand(16384, 16384, device=dev) * 2 - 1
for _ in range(2):
  engine.max_margin(x, 0.05, True, want_grad=False)
torch.cuda.synchronize()
print("done", float(loss))
