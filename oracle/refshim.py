"""Import the UNMODIFIED reference (/root/reference) with import-time shims.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (in the build container, where
/root/reference exists) to produce the golden fixtures under tests/golden/.  Nothing in the
product path (mmt_b200/) imports this module, and nothing that runs on the GPU box may, because
/root/reference does not exist there.

The shims only register empty stand-ins for third-party modules that are absent in this image
(tensorboardX, ipdb, h5py, pytorch_warmup, dominate, gensim) and alias the removed
`transformers.modeling_bert` module; no reference file is edited.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MMT_REFERENCE_ROOT", "/root/reference")


def reference_available():
  return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "model.py"))


def _shim(name, **attrs):
  if name in sys.modules:
    return
  m = types.ModuleType(name)
  m.__dict__.update(attrs)
  sys.modules[name] = m


def load_reference():
  """Returns (model.model, model.loss, model.metric, utils.util) of the reference."""
  if not reference_available():
    raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
  _shim("tensorboardX", SummaryWriter=object)
  _shim("ipdb", set_trace=lambda *a, **k: None)
  _shim("h5py")
  _shim("pytorch_warmup")
  _shim("dominate")
  _shim("dominate.tags")
  _shim("gensim")
  _shim("gensim.models")
  _shim("gensim.models.keyedvectors", KeyedVectors=object)
  _shim("gensim.scripts")
  _shim("gensim.scripts.glove2word2vec", glove2word2vec=None)

  from transformers import BertConfig
  from transformers.models.bert import modeling_bert as mb

  class TxtBertModel(mb.BertModel):
    # No network: random-init model of the requested geometry (bert-base-cased by default).
    @classmethod
    def from_pretrained(cls, name, **kw):
      return cls(BertConfig(vocab_size=28996, **kw))

  _shim("transformers.modeling_bert", BertModel=TxtBertModel)

  # The repo's own top-level package is also called `model` when mmt_b200 is used as a drop-in;
  # make sure we get the reference's.
  for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
    del sys.modules[k]
  import model.model as ref_model
  import model.loss as ref_loss
  import model.metric as ref_metric
  import utils.util as ref_util
  return ref_model, ref_loss, ref_metric, ref_util


import torch


class TxtBertStub(torch.nn.Module):
  """Stand-in for the third-party text encoder (hot-path-only scope).

  Returns a fixed [B*caps, W, text_dim] tensor as `last_hidden_state`; the reference reads
  `[0][:, 0]` (model.py:371-379).
  """

  def __init__(self, hidden):
    super().__init__()
    self.hidden = hidden
    self.config = types.SimpleNamespace(hidden_size=hidden.shape[-1])

  def forward(self, input_ids, **kw):
    return (self.hidden,)
