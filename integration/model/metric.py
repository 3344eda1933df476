"""`model.metric`: retrieval metrics with the rank extraction on the GPU (reference model/metric.py:26-258)."""
from mmt_b200.model.metric import cols2metrics, retrieval_ranks, t2v_metrics, v2t_metrics  # noqa: F401
