"""The reference's train-step sequence on the drop-in module (GPU).

`Trainer._train_epoch` (reference trainer/trainer.py:160-210) and the objects `train.py:84-108` builds are restated
verbatim below -- move_dict_to_device, optional warm-up dampening, optimizer.zero_grad, model(**minibatch, out=...,
device=..., debug=...), the loss on output["cross_view_conf_matrix"] (out='conf') or on the
sharded_cross_view_inner_product of the returned embeddings (out='embds'), backward, optimizer.step, loss.item(),
StepLR per epoch -- with torch.optim.Adam exactly as train.py:95-100 constructs it (over
filter(requires_grad, model.parameters())) and, next to it, mmt_b200.optim.FusedAdam.  Both must decrease the loss
on a fixed batch and agree with each other step by step (dropout off)."""
import collections
import math

import pytest
import torch

import mmt_test_helpers as H

pytestmark = pytest.mark.gpu

MODS = ["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"]


def move_dict_to_device(res, device):                      # reference trainer/trainer.py:21-35, verbatim semantics
  for key in list(res.keys()):
    value = res[key]
    if isinstance(value, (dict, collections.OrderedDict)):
      res[key] = move_dict_to_device(res[key], device)
    elif isinstance(value, torch.Tensor):
      res[key] = value.to(device)
  return res


def _train_iterations(net, loss_fn, optimizer, lr_scheduler, minibatch_cpu, out, n_iter, device):
  from mmt_b200.model.model import sharded_cross_view_inner_product
  modalities = net.modalities
  losses = []
  for batch_idx in range(n_iter):
    minibatch = move_dict_to_device({k: (dict(v) if isinstance(v, dict) else v) for k, v in minibatch_cpu.items()}, device)
    optimizer.zero_grad()
    output = net(**minibatch, out=out, device=device, debug=False)
    if out == "conf":
      loss = loss_fn(output["cross_view_conf_matrix"])
    else:
      vid_embds = collections.OrderedDict()
      text_embds = collections.OrderedDict()
      for idx, mod in enumerate(modalities):
        vid_embds[mod] = output["vid_embds"][:, idx]
        text_embds[mod] = output["text_embds"][:, idx]
      conf = sharded_cross_view_inner_product(vid_embds=vid_embds, text_embds=text_embds,
                                              vid_weights=output["vid_weights"], text_weights=output["text_weights"],
                                              subspaces=modalities, merge_caption_similiarities="avg")
      loss = loss_fn(conf)
    loss.backward()
    optimizer.step()
    losses.append(loss.item())
  lr_scheduler.step()                                       # trainer.py:245-246 (once per epoch)
  return losses


@pytest.mark.parametrize("out", ["conf", "embds"])
def test_reference_train_sequence_with_stock_adam_and_fused_adam(out):
  from mmt_b200.model.loss import MaxMarginRankingLoss
  from mmt_b200.optim import FusedAdam
  ed, vb, P, batch, cfg = H.make_case(MODS, 16, 30, layers=2)
  minibatch = H.batch_kwargs(batch)
  device = torch.device("cuda")
  runs = {}
  for kind in ("torch.optim.Adam", "FusedAdam"):
    net = H.build_cuda_net(ed, vb, P, batch, precision="f16").train()
    loss_fn = MaxMarginRankingLoss(margin=0.05, fix_norm=True)
    if kind == "FusedAdam":
      optimizer = FusedAdam(net, lr=1e-4, weight_decay=0.0)
    else:
      trainable_params = filter(lambda p: p.requires_grad, net.parameters())      # train.py:93
      optimizer = torch.optim.Adam(trainable_params, lr=1e-4, weight_decay=0.0)    # train.py:98, configs' optimizer
    lr_scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=0.95)  # configs' lr_scheduler
    runs[kind] = _train_iterations(net, loss_fn, optimizer, lr_scheduler, minibatch, out, 4, device)
    assert optimizer.param_groups[0]["lr"] == pytest.approx(0.95e-4)
    assert all(math.isfinite(x) for x in runs[kind])
    assert runs[kind][-1] < runs[kind][0], (kind, runs[kind])
  print(out, runs)
  for a, b in zip(runs["torch.optim.Adam"], runs["FusedAdam"]):
    assert abs(a - b) < 2e-4 * max(1.0, abs(a)), runs
