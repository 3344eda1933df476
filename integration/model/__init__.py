"""Drop-in `model` package for gabeur/mmt that needs no edit of the reference tree.

Put this directory's parent (`integration/`) AHEAD of the reference checkout on `sys.path`
(`PYTHONPATH=/path/to/mmt_b200_repo/integration:/path/to/mmt_b200_repo:/path/to/mmt python train.py ...`)
and point `MMT_REFERENCE_ROOT` at the checkout.  `model.model`, `model.loss` and `model.metric` then
resolve to the files next to this one (the B200 hot path: `CENet`, `sharded_cross_view_inner_product`,
`MaxMarginRankingLoss`, `t2v_metrics` / `v2t_metrics`), every other submodule the reference imports
(`model.bert`, `model.net_vlad`, `model.txt_embeddings`, ...) still comes from the reference, because
its `model/` directory is appended to this package's search path.  `train.py`, `trainer/trainer.py`
and the configs run unchanged (reference train.py:86-93, trainer/trainer.py:27).
"""
import os

_ref = os.environ.get("MMT_REFERENCE_ROOT")
if _ref:
  _ref_model = os.path.join(_ref, "model")
  if not os.path.isdir(_ref_model):
    raise ImportError("MMT_REFERENCE_ROOT=%s has no model/ directory" % _ref)
  __path__.append(_ref_model)           # submodules not overridden here are the reference's own
