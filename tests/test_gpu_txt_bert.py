"""GPU parity of mmt_b200.model.txt_bert.TxtBert against transformers' BertModel -- the module
the reference instantiates (model/model.py:161) -- with the SAME random weights on both sides: last hidden state
and every parameter gradient.  The third-party encoder has no golden vectors in the reference ("parity unpinned"
at that boundary); this is module-for-module self-consistency on bert-base geometry."""
import math

import pytest
import torch

import mmt_test_helpers as H

pytestmark = pytest.mark.gpu


def _hf(layers, vocab=1000, seed=0, dropout=0.0):
  from transformers import BertConfig, BertModel
  torch.manual_seed(seed)
  cfg = BertConfig(vocab_size=vocab, hidden_size=768, num_hidden_layers=layers, num_attention_heads=12,
                   intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, hidden_dropout_prob=dropout,
                   attention_probs_dropout_prob=dropout)
  try:
    cfg._attn_implementation = "eager"
  except Exception:
    pass
  return BertModel(cfg).train()


@pytest.mark.parametrize("R,W,layers", [(8, 30, 2), (4, 100, 1), (64, 30, 12)])
def test_txt_bert_matches_transformers_bert_model(R, W, layers):
  from mmt_b200.model.txt_bert import TxtBert
  hf = _hf(layers)
  net = TxtBert.from_hf(hf).cuda().train()
  g = torch.Generator().manual_seed(R * 7 + W)
  ids = torch.randint(0, 1000, (R, W), generator=g)
  lens = torch.randint(5, W + 1, (R,), generator=g)
  mask = (torch.arange(W)[None, :] < lens[:, None]).long()
  pos = torch.arange(W)[None, :].expand(R, W)
  tt = torch.zeros(R, W, dtype=torch.long)
  # upstream gradient of the magnitude the ranking loss produces (~1/n per row): 16-bit gradient tensors carry the
  # fixed 2^16 scale of the train step, which assumes |d loss / d activation| << 1
  probe = torch.randn(R, 768, generator=g) * 1e-3
  # reference: CLS row of the last hidden state (what the reference consumes, model/model.py:378-379)
  out_ref = hf(ids, attention_mask=mask, token_type_ids=tt, position_ids=pos)[0]
  (out_ref[:, 0] * probe).sum().backward()
  out = net(ids.cuda(), attention_mask=mask.cuda(), token_type_ids=tt.cuda(), position_ids=pos.cuda(), head_mask=None)[0]
  (out[:, 0] * probe.cuda()).sum().backward()
  torch.cuda.synchronize()
  valid = mask.bool()
  e_cls = H.rel_err(out[:, 0], out_ref[:, 0])
  e_all = H.rel_err(out[valid.cuda()], out_ref[valid])
  ref_grads = {k: v.grad for k, v in hf.named_parameters() if v.grad is not None and k in net.layout.segments}
  g_max, g_l2, worst, worst_l2 = H.grad_errors(net, ref_grads)
  print("TxtBert R=%d W=%d L=%d: CLS max-rel %.2e, valid tokens max-rel %.2e | gradient (whole) max-norm %.2e rel-L2 %.2e | "
        "worst tensor %s %.2e" % (R, W, layers, e_cls, e_all, g_max, g_l2, worst[0], worst[1]))
  # every GEMM operand is rounded to an 11-bit significand once: zero-mean noise of ~4e-4 per GEMM that grows with
  # the square root of the depth (forward: 6 GEMMs per layer; a gradient that reaches the embeddings of a 12-layer
  # stack has crossed ~70 of them forward and backward)
  assert e_cls < (1e-3 if layers <= 2 else 2e-3) and e_all < (2e-3 if layers <= 2 else 3e-3)
  assert g_max < (2e-3 if layers <= 2 else 1e-2) and g_l2 < (2e-3 if layers <= 2 else 4e-3)


def test_txt_bert_dropout_and_fused_adam_in_cenet():
  """CENet with the native text encoder, dropout 0.1, FusedAdam over BOTH flat buffers: loss decreases, the 16-bit
  weight copies follow the fp32 master weights."""
  from mmt_b200.model.loss import MaxMarginRankingLoss
  from mmt_b200.model.txt_bert import TxtBert
  from mmt_b200.optim import FusedAdam
  ed, vb, P, batch, cfg = H.make_case(["s3d", "vggish"], 16, 14, layers=2, dropout=0.1)
  tb = TxtBert.from_hf(_hf(2, vocab=28996, dropout=0.1))
  net = H.build_cuda_net(ed, vb, P, batch, dropout=0.1, precision="f16", txt_bert=tb).train()
  opt = FusedAdam(net, lr=2e-5)
  assert len(opt.sub) == 1 and opt.other is None
  crit = MaxMarginRankingLoss(0.05, True)
  kw = H.batch_kwargs(batch, "cuda")
  losses = []
  for _ in range(6):
    opt.zero_grad()
    l = crit(net(**kw)["cross_view_conf_matrix"])
    l.backward()
    opt.step()
    losses.append(float(l))
  assert all(math.isfinite(x) for x in losses), losses
  assert losses[-1] < losses[0], losses
  assert torch.equal(net.txt_bert.w16.flat16, net.txt_bert.flat.half())
  assert net.txt_bert._param("encoder.layer.0.output.dense.weight").grad is not None
