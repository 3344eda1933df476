"""Host-side cost of the eager train step (development aid): wall time per step when the host never waits for the
device (queue depth permitting) and a cProfile of a few steps.  Run on a GPU box."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import workloads as W  # noqa: E402
from mmt_b200 import _lib  # noqa: E402
from mmt_b200.model.loss import MaxMarginRankingLoss  # noqa: E402
from mmt_b200.model.model import CENet  # noqa: E402
from mmt_b200.optim import FusedAdam  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
w = bench.WORKLOADS["C2"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8          # small batch: the device is never the bottleneck
ed, batches = bench.make_batches(w, B, 1, 1234)
P = W.init_params(ed, bench.vb_params(w), seed=0)
feed = bench.TextFeed()
net = CENet(l2renorm=False, expert_dims=ed, tokenizer=None, keep_missing_modalities=True, test_caption_mode="indep",
            txt_inp="bertftn", txt_agg="bertftn", txt_wgh="emb", vid_wgh="none", vid_cont="bert", vid_inp="both",
            pos_enc="tint", out_tok="mxp", vid_bert_params=bench.vb_params(w), txt_pro="gbn",
            txt_bert_params={"hidden_dropout_prob": 0.1, "attention_probs_dropout_prob": 0.1}, txt_bert=feed)
net.load_state_dict(P, strict=True)
net.to(dev).train()
opt = FusedAdam(net, lr=5e-5)
crit = MaxMarginRankingLoss(0.05, True)
b = batches[0]
kw = {k: {m: v.to(dev) for m, v in b[k].items()} for k in ("features", "features_t", "features_ind", "features_avgpool",
                                                          "features_maxpool")}
kw["token_ids"], kw["query_masks"] = b["token_ids"].to(dev), b["query_masks"]
feed.cls = b["text_feat"].to(dev)


def step():
  opt.zero_grad()
  loss = crit(net(**kw, out="conf", device=dev)["cross_view_conf_matrix"])
  loss.backward()
  opt.step()


for _ in range(5):
  step()
torch.cuda.synchronize()
n = 30
t0 = time.perf_counter()
for _ in range(n):
  step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host time per eager step (B=%d): %.3f ms (device drained %.3f ms later); %d library launches per step" %
      (B, (t1 - t0) / n * 1e3, (t2 - t1) * 1e3, 0))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
  step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
