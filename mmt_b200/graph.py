"""CUDA-graph capture of the whole train step (zero_grad -> forward -> loss -> backward -> Adam).

The step is ~150 kernel launches of 3-300 us each; replaying them as one graph removes the
per-launch host/driver latency (SURVEY.md §7 "launch count must collapse").  Everything the
step needs that changes from step to step lives in device memory:
  * inputs: static device tensors the caller refreshes with `copy_` (or `load()`),
  * dropout seeds and Adam's bias-correction step: a device counter the graph increments itself
    (`mmt_set_step_counter`), so every replay draws fresh masks.
"""
import torch

from . import _lib


class GraphedTrainStep:
  """graph = GraphedTrainStep(net, crit, opt, example_kwargs, example_text, set_text)

  example_kwargs : dict of CENet.forward kwargs holding DEVICE tensors (kept as the static inputs)
  set_text(t)    : hands the static text-feature tensor to whatever stands for txt_bert
  """

  def __init__(self, net, crit, opt, kwargs, text, set_text, warmup=3, share_with=None):
    import os
    if getattr(net, "_dp", False) and os.environ.get("MMT_GRAPH_DP", "0") != "1":
      # Capturing the data-parallel step (NCCL all-gather / all-reduce nodes inside the graph) is EXPERIMENTAL: the
      # first attempt on 2 B200s (torch 2.11, NCCL 2.28, TORCH_NCCL_ASYNC_ERROR_HANDLING=0) never finished its
      # capture.  Opt in with MMT_GRAPH_DP=1; the supported data-parallel path is the eager step.
      raise NotImplementedError("GraphedTrainStep: the data-parallel step is not captured (set MMT_GRAPH_DP=1 to try); "
                                "use the eager step with enable_data_parallel()")
    self.net, self.crit, self.opt = net, crit, opt
    self.kw, self.text = kwargs, text
    dev = net.flat.device
    # `share_with`: another GraphedTrainStep over the SAME net / optimizer whose step counter this one joins.  Two
    # captures with their own static input buffers let the host stage batch i+1 straight into the idle graph's
    # inputs (H2D, no device-to-device copy) while graph i runs; one shared counter keeps dropout seeds and Adam's
    # bias correction advancing by one per replay of either graph.
    self.shared = share_with is not None
    self.ctr = share_with.ctr if self.shared else torch.zeros(1, dtype=torch.int64, device=dev)
    self._attach()
    set_text(self.text)

    def one_step():
      opt.zero_grad()
      out = net(**self.kw, out="conf", device=dev)
      loss = crit(out["cross_view_conf_matrix"])
      loss.backward()
      opt.step()
      return loss

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):                       # eager warm-up on the capture stream
      for _ in range(warmup):
        one_step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph):
      self.ctr.add_(1)
      self.loss = one_step()
    # the python-side step counters advanced once during capture; replays advance the device one

  def _attach(self):
    """Point the step's kernels at the device counter: the 16-bit path takes it per call (net.cfg.seed_ctr,
    opt.step_ctr); the fp32 / tf32 entry points read the library-wide pointer."""
    if _lib.is16(self.net.cfg.precision):
      self.net.cfg.seed_ctr = _lib.ptr(self.ctr)
      if hasattr(getattr(self.net, "txt_bert", None), "seed_ctr"):
        self.net.txt_bert.seed_ctr = _lib.ptr(self.ctr)
      if hasattr(self.opt, "step_ctr"):
        self.opt.step_ctr = _lib.ptr(self.ctr)
    else:
      _lib.check(_lib.load().mmt_set_step_counter(_lib.ptr(self.ctr)), "mmt_set_step_counter")

  def _detach(self):
    if _lib.is16(self.net.cfg.precision):
      self.net.cfg.seed_ctr = None
      if hasattr(getattr(self.net, "txt_bert", None), "seed_ctr"):
        self.net.txt_bert.seed_ctr = None
      if hasattr(self.opt, "step_ctr"):
        self.opt.step_ctr = None
    else:
      _lib.check(_lib.load().mmt_set_step_counter(None), "mmt_set_step_counter")

  def __del__(self):
    try:
      if getattr(self, "ctr", None) is not None:
        self._detach()
    except Exception:
      pass

  def load(self, kwargs, text):
    """Refresh the static inputs from other (host-pinned or device) tensors."""
    for k, v in kwargs.items():
      if isinstance(v, dict):
        for m, t in v.items():
          self.kw[k][m].copy_(t, non_blocking=True)
      elif torch.is_tensor(v) and torch.is_tensor(self.kw.get(k)) and self.kw[k].is_cuda:
        self.kw[k].copy_(v, non_blocking=True)
    self.text.copy_(text, non_blocking=True)

  def replay(self):
    self.graph.replay()
    return self.loss

  def close(self):
    """Back to eager launches: detach the device counter and fold its value into the host-side
    step counts so Adam's bias correction continues where the replays left off."""
    torch.cuda.synchronize()
    if self.shared:                   # the owner of the counter does the bookkeeping
      self.ctr = None
      return
    n = int(self.ctr.item())
    self._detach()
    self.ctr = None
    if hasattr(self.opt, "t"):
      self.opt.t += n
    self.net._step += n
