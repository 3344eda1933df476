"""Retrieval metrics of the eval path with the rank extraction on the GPU.

Same surface as the reference's model/metric.py: `t2v_metrics(sims, query_masks=None)`,
`v2t_metrics(sims, query_masks=None)` -> dict with R1, R5, R10, R50, MedR, MeanR,
geometric_mean_R1-R5-R10 and `cols` (metric.py:26-258).  The reference sorts every row of the
N_q x N_v distance matrix on the CPU and searches the ground truth in it; here the rank is counted
(#strictly better + (#equal - 1) / 2, which is what its tie "averaging" evaluates to) by
`mmt_retrieval_ranks` directly on the similarity matrix -- on the device it was computed on, if it
is still there.  Counting on the same fp32 values is exact: ranks agree with the reference bit for bit.
"""
import numpy as np
import torch

from .. import _lib


def _device_sims(sims):
  if isinstance(sims, np.ndarray):
    sims = torch.from_numpy(np.ascontiguousarray(sims, dtype=np.float32))
  if not sims.is_cuda:
    if not torch.cuda.is_available():
      raise RuntimeError("mmt_b200.model.metric needs a CUDA device (no CPU fallback)")
    sims = sims.to(torch.device("cuda", torch.cuda.current_device()))
  return sims.to(torch.float32).contiguous()


def retrieval_ranks(sims, query_masks=None, v2t=False):
  """0-based ranks (float64 numpy): [Nq] for t2v (masked queries NOT removed), [Nv] for v2t."""
  x = _device_sims(sims)
  if x.dim() != 2:
    raise AssertionError("expected a matrix")
  nq, nv = x.shape
  valid = None
  if query_masks is not None:
    qm = torch.as_tensor(np.asarray(query_masks).reshape(-1) if not torch.is_tensor(query_masks)
                         else query_masks.reshape(-1))
    if qm.numel() != nq:
      raise AssertionError("invalid query mask shape")
    valid = (qm != 0).to(device=x.device, dtype=torch.int32).contiguous()
  out = torch.empty(nv if v2t else nq, device=x.device, dtype=torch.float32)
  with torch.cuda.device(x.device):
    _lib.check(_lib.load().mmt_retrieval_ranks(_lib.ptr(x), _lib.ptr(valid), nq, nv, 1 if v2t else 0,
                                               _lib.ptr(out), _lib.stream_ptr()), "mmt_retrieval_ranks")
  return out.cpu().numpy().astype(np.float64)


def cols2metrics(cols, num_queries):
  """model/metric.py:246-258."""
  cols = np.asarray(cols, dtype=np.float64)
  metrics = {}
  metrics["R1"] = 100 * float(np.sum(cols == 0)) / num_queries
  metrics["R5"] = 100 * float(np.sum(cols < 5)) / num_queries
  metrics["R10"] = 100 * float(np.sum(cols < 10)) / num_queries
  metrics["R50"] = 100 * float(np.sum(cols < 50)) / num_queries
  metrics["MedR"] = np.median(cols) + 1
  metrics["MeanR"] = np.mean(cols) + 1
  stats = np.array([metrics[x] for x in ("R1", "R5", "R10")])
  metrics["geometric_mean_R1-R5-R10"] = float(np.exp(np.mean(np.log(stats)))) if np.all(stats > 0) else 0.0
  metrics["cols"] = [int(i) for i in list(cols)]
  return metrics


def t2v_metrics(sims, query_masks=None):
  """Text-to-video retrieval (metric.py:26-150): masked queries are dropped after ranking."""
  cols = retrieval_ranks(sims, None, v2t=False)
  nq = cols.size
  if query_masks is not None:
    qm = np.asarray(query_masks.cpu() if torch.is_tensor(query_masks) else query_masks).reshape(-1).astype(bool)
    assert qm.size == nq, "invalid query mask shape"
    cols = cols[qm]
    nq = int(qm.sum())
  return cols2metrics(cols, nq)


def v2t_metrics(sims, query_masks=None):
  """Video-to-text retrieval (metric.py:152-230): rank of the closest existing ground-truth caption."""
  cols = retrieval_ranks(sims, query_masks, v2t=True)
  return cols2metrics(cols, cols.size)
