"""`model.loss`: the published configs select MaxMarginRankingLoss (reference model/loss.py:32-65)."""
from mmt_b200.model.loss import MaxMarginRankingLoss  # noqa: F401
