"""Fused Adam over the model's flat parameter / gradient buffers (one kernel per contiguous run).

Semantics = torch.optim.Adam (reference train.py:95-98: Adam(lr, weight_decay) over
filter(requires_grad, model.parameters())): parameters that received no gradient in this step (the unused
pooler, frozen parameters, anything when no backward ran) are skipped, exactly as torch skips `p.grad is None`.
Parameters outside the flat buffer (e.g. the third-party text encoder) are handed to a regular torch.optim.Adam.

The object quacks like a torch optimizer where the reference's trainer touches one: `param_groups` (the
learning rate is read from `param_groups[0]["lr"]` at every step, so `torch.optim.lr_scheduler.StepLR` and
`pytorch_warmup` dampening -- train.py:101-108, trainer.py:245-246 -- drive it), `zero_grad`, `step`,
`state_dict` / `load_state_dict` (moments and step count, base_trainer.py:367, 442).

In the 16-bit operand modes the same kernel also writes the GEMMs' 16-bit weight copy, so no separate cast
pass runs in a FusedAdam-driven step.
"""
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
  """A torch.optim.Optimizer (so lr schedulers / warm-up wrappers accept it) whose step is one fused kernel per
  contiguous run of the flat parameter buffer."""

  def __init__(self, net, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    trainable = [p for p in net._hot_params() if p.requires_grad]
    super().__init__(trainable, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
    self.net = net
    self.grad_scale = grad_scale
    self.t = 0
    self.m = torch.zeros_like(net.flat)
    self.v = torch.zeros_like(net.flat)
    hot = set(id(p) for p in net._hot_params())
    # a text encoder that also lives in a flat buffer (mmt_b200.model.txt_bert.TxtBert) gets its own fused instance
    self.sub = []
    tb = getattr(net, "txt_bert", None)
    if tb is not None and hasattr(tb, "flat") and hasattr(tb, "_hot_params") and any(p.requires_grad for p in tb._hot_params()):
      self.sub.append(FusedAdam(tb, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale))
      hot |= set(id(p) for p in tb._hot_params())
    others = [p for p in net.parameters() if id(p) not in hot and p.requires_grad]
    self.other = torch.optim.Adam(others, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay) \
        if others else None
    self.param_groups[0].setdefault("initial_lr", lr)
    self._runs = None
    self.step_ctr = None        # device uint64 pointer added to the bias-correction step (CUDA-graph replays)

  # hyper-parameters are read from param_groups[0] at every step (lr schedulers mutate it)
  @property
  def lr(self):
    return self.param_groups[0]["lr"]

  @lr.setter
  def lr(self, v):
    self.param_groups[0]["lr"] = v

  def _sync_hyper(self, opt):
    for g in opt.param_groups:
      g.update({k: self.param_groups[0][k] for k in ("lr", "betas", "eps", "weight_decay")})

  def _compute_runs(self):
    """Contiguous [offset, end) ranges of the flat buffer whose parameters are trained."""
    runs = []
    L = self.net.layout
    for name, seg in L.segments.items():
      p = self.net._param(name)
      if not p.requires_grad or name.startswith(("vid_bert.pooler.", "pooler.")):
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
    for sub in self.sub:
      sub.zero_grad(set_to_none)
    for p in self.net._hot_params():
      p.grad = None
    if self.other is not None:
      self.other.zero_grad(set_to_none=set_to_none)

  def state_dict(self):
    return {"t": self.t, "m": self.m, "v": self.v, "param_groups": [{k: v for k, v in g.items() if k != "params"}
                                                                   for g in self.param_groups[:1]],
            "other": self.other.state_dict() if self.other is not None else None,
            "sub": [s_.state_dict() for s_ in self.sub]}

  def load_state_dict(self, sd):
    self.t = int(sd["t"])
    self.m.copy_(sd["m"])
    self.v.copy_(sd["v"])
    for k, v in sd["param_groups"][0].items():
      self.param_groups[0][k] = v
    if self.other is not None and sd.get("other") is not None:
      self.other.load_state_dict(sd["other"])
    for s_, d_ in zip(self.sub, sd.get("sub", [])):
      s_.load_state_dict(d_)

  @torch.no_grad()
  def step(self, closure=None):
    net = self.net
    for sub in self.sub:
      self._sync_hyper(sub)
      sub.step_ctr = self.step_ctr
      sub.step()
    if self.m.device != net.flat.device:
      self.m, self.v = self.m.to(net.flat.device), self.v.to(net.flat.device)
    if self._runs is None:
      self._runs = self._compute_runs()
    params = net._hot_params()
    g = net._grad_flat()
    # gradients are normally views of net._gflat (EncodeFn publishes them without copies); if the
    # caller accumulated / replaced them, gather them back into the flat layout first
    views = net.__dict__.get("_grad_views")
    fresh = False
    missing = []
    for i, p in enumerate(params):
      if p.grad is None:
        if p.requires_grad and not net._names[i].startswith(("vid_bert.pooler.", "pooler.")):
          missing.append(i)
        continue
      fresh = True
      if views is not None and views[0] == g.data_ptr() and p.grad is views[1][i]:
        continue                                            # the common case: already a view of g
      v = net.layout.view(g, net._names[i])
      if p.grad.data_ptr() != v.data_ptr():
        v.copy_(p.grad)
    if fresh:
      self.t += 1
      g_hp = self.param_groups[0]
      lr, (b1, b2), eps, wd = g_hp["lr"], g_hp["betas"], g_hp["eps"], g_hp["weight_decay"]
      if missing:
        # torch.optim.Adam skips parameters without a gradient: run only over the others (rare path)
        runs = []
        skip = set(missing)
        for i, p in enumerate(params):
          if p.grad is None or i in skip:
            continue
          seg = net.layout.segments[net._names[i]]
          runs.append([seg.offset, min(seg.offset + (seg.numel + 3) // 4 * 4, net.layout.numel)])
      else:
        runs = self._runs
      lib = _lib.load()
      st = _lib.stream_ptr()
      w16 = getattr(net, "w16", None)           # 16-bit weight copy of the module (None in the fp32 / tf32 modes)
      for lo, hi in runs:
        if w16 is not None:
          _lib.check(lib.mmt_adam16_step(_lib.ptr(net.flat, lo), _lib.ptr(g, lo), _lib.ptr(self.m, lo),
                                         _lib.ptr(self.v, lo), _lib.ptr(w16.flat16, lo), hi - lo, lr, b1, b2, eps, wd,
                                         self.t, self.step_ctr, self.grad_scale, w16.dt, st), "mmt_adam16_step")
        else:
          _lib.check(lib.mmt_adam_step(_lib.ptr(net.flat, lo), _lib.ptr(g, lo), _lib.ptr(self.m, lo),
                                       _lib.ptr(self.v, lo), hi - lo, lr, b1, b2, eps, wd, self.t,
                                       self.grad_scale, st), "mmt_adam_step")
      if w16 is not None and hasattr(net, "cfg"):
        w16.refresh_padded(net.cfg, net.flat)      # row-padded ReduceDim copies (two small casts)
    if self.other is not None:
      self._sync_hyper(self.other)
      self.other.step()
