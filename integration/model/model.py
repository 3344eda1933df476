"""`model.model` as the reference's train.py / trainer import it, served by the B200 hot path."""
from mmt_b200.model.model import CENet, sharded_cross_view_inner_product  # noqa: F401
