"""Fused Adam over the model's flat parameter / gradient buffers (one kernel per contiguous run).

Semantics = torch.optim.Adam (reference train.py:95-98: Adam(lr, weight_decay) over
filter(requires_grad, model.parameters())): parameters that received no gradient (the unused
pooler, frozen parameters) are skipped, exactly as torch skips `p.grad is None`.
Parameters outside the flat buffer (e.g. the third-party text encoder) are handed to a regular
torch.optim.Adam.
"""
import torch

from . import _lib


class FusedAdam:

  def __init__(self, net, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    self.net = net
    self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
    self.grad_scale = grad_scale
    self.t = 0
    self.m = torch.zeros_like(net.flat)
    self.v = torch.zeros_like(net.flat)
    hot = set(id(p) for p in net._hot_params())
    others = [p for p in net.parameters() if id(p) not in hot and p.requires_grad]
    self.other = torch.optim.Adam(others, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay) \
        if others else None
    self._runs = None

  def _compute_runs(self):
    """Contiguous [offset, end) ranges of the flat buffer whose parameters are trained."""
    runs = []
    L = self.net.layout
    for name, seg in L.segments.items():
      p = self.net._param(name)
      if not p.requires_grad or name.startswith("vid_bert.pooler."):
        continue
      lo, hi = seg.offset, seg.offset + (seg.numel + 3) // 4 * 4
      if runs and runs[-1][1] >= lo:
        runs[-1][1] = max(runs[-1][1], hi)
      else:
        runs.append([lo, hi])
    for r in runs:
      r[1] = min(r[1], L.numel)
    return runs

  def zero_grad(self, set_to_none=True):
    for p in self.net._hot_params():
      p.grad = None
    if self.other is not None:
      self.other.zero_grad(set_to_none=set_to_none)

  def step(self):
    net = self.net
    if self.m.device != net.flat.device:
      self.m, self.v = self.m.to(net.flat.device), self.v.to(net.flat.device)
    if self._runs is None:
      self._runs = self._compute_runs()
    g = net._grad_flat()
    # gradients are normally views of net._gflat (EncodeFn publishes them without copies); if the
    # caller accumulated / replaced them, gather them back into the flat layout first
    views = net.__dict__.get("_grad_views")
    for i, p in enumerate(net._hot_params()):
      if p.grad is None:
        continue
      if views is not None and views[0] == g.data_ptr() and p.grad is views[1][i]:
        continue                                            # the common case: already a view of g
      v = net.layout.view(g, net._names[i])
      if p.grad.data_ptr() != v.data_ptr():
        v.copy_(p.grad)
    self.t += 1
    lib = _lib.load()
    st = _lib.stream_ptr()
    for lo, hi in self._runs:
      _lib.check(lib.mmt_adam_step(_lib.ptr(net.flat, lo), _lib.ptr(g, lo), _lib.ptr(self.m, lo),
                                   _lib.ptr(self.v, lo), hi - lo, self.lr, self.betas[0],
                                   self.betas[1], self.eps, self.wd, self.t, self.grad_scale, st),
                 "mmt_adam_step")
    if self.other is not None:
      self.other.step()
