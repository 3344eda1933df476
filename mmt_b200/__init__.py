"""mmt_b200: B200-native (sm_100a) implementation of the gabeur/mmt training hot path.

`mmt_b200.model.model.{CENet, sharded_cross_view_inner_product}` and
`mmt_b200.model.loss.MaxMarginRankingLoss` are drop-ins for the reference's `model.model` /
`model.loss` symbols.  All arithmetic runs in libmmt_b200.so (C ABI:
include/mmt_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
