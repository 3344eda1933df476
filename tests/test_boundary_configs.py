"""Boundary tests against the reference's own configuration files (CPU; need /root/reference, skipped elsewhere):
the drop-in CENet built from every configs_pub/eccv20/*.json through the reference's `compute_dims` has exactly the
reference module's state_dict -- names, shapes, trainable flags -- including the text encoder, whose native
implementation (mmt_b200.model.txt_bert.TxtBert) must mirror transformers' BertModel."""
import glob
import json
import os

import pytest
import torch

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="needs the reference checkout")

CONFIGS = sorted(glob.glob(os.path.join(refshim.REFERENCE_ROOT, "configs_pub", "eccv20", "*.json")))


@pytest.mark.parametrize("path", CONFIGS, ids=[os.path.basename(p) for p in CONFIGS])
def test_state_dict_equals_the_reference_module(path, monkeypatch):
  ref_model, _, _, ref_util = refshim.load_reference()
  cfg = json.load(open(path))
  expert_dims = ref_util.compute_dims(cfg)
  args = cfg["arch"]["args"]
  torch.manual_seed(0)
  ref = ref_model.CENet(expert_dims=expert_dims, tokenizer=None, **args)
  # our module, with the text encoder taken from the reference instance (same geometry / weights), converted to the
  # native TxtBert exactly as CENet does by default when it builds the encoder itself
  from mmt_b200.model.model import CENet
  from mmt_b200.model.txt_bert import TxtBert
  ours = CENet(expert_dims=expert_dims, tokenizer=None, txt_bert=TxtBert.from_hf(ref.txt_bert), **args)
  # the constructor froze the text encoder per `txt_agg` / `txt_inp` (model.py:164-193) on the module it was given
  sd_ref, sd = ref.state_dict(), ours.state_dict()
  skip = lambda k: k.endswith("position_ids")          # transformers buffer, not a parameter
  assert set(k for k in sd if not skip(k)) == set(k for k in sd_ref if not skip(k)), \
      set(sd) ^ set(sd_ref)
  for k, v in sd_ref.items():
    if skip(k):
      continue
    assert tuple(sd[k].shape) == tuple(v.shape), k
  tr_ref = {n: p.requires_grad for n, p in ref.named_parameters()}
  tr = {n: p.requires_grad for n, p in ours.named_parameters()}
  assert tr == tr_ref
  assert sum(p.numel() for p in ours.parameters() if p.requires_grad) == \
      sum(p.numel() for p in ref.parameters() if p.requires_grad)
  # a reference checkpoint loads, and round-trips bit-exactly through the flat buffers
  ours.load_state_dict({k: v for k, v in sd_ref.items() if not skip(k)}, strict=True)
  for k, v in ours.state_dict().items():
    if not skip(k):
      assert torch.equal(v, sd_ref[k]), k
