"""The text encoder on the repo's kernels.

The reference's `txt_bert` is `transformers.BertModel.from_pretrained('bert-base-cased', **txt_bert_params)`
(reference model/model.py:150-193) and is called as
    txt_bert(input_ids, attention_mask=..., token_type_ids=..., position_ids=..., head_mask=None)[0]
(model/model.py:350-387).  `TxtBert` is a drop-in for that module on this call: same constructor geometry
(`BertConfig` fields), the SAME parameter names and shapes as transformers' BertModel (so `load_state_dict` of a
BertModel checkpoint -- e.g. the pretrained bert-base-cased weights -- works, and `TxtBert.from_hf(model)` copies
a live module), `.config.hidden_size`, and a tuple output whose element 0 is the last hidden state [R, W, d].

Inside, the parameters live in one flat fp32 buffer (views under the HF names), the encoder layers are
engine16.layers_forward / layers_backward -- the video encoder's fused tcgen05 GEMMs and LayerNorm kernels -- and
the embedding gather + LayerNorm and the short-sequence attention (dh = 64, W <= 128) are csrc/txtops.cu.  The
backward is hand-written like the video side's.  Only what the reference uses is implemented: token type 0,
positions 0..W-1 (what model.py:359-367 passes), no head mask, the pooler kept as (unused) parameters.
"""
import collections
import types

import torch
import torch.nn as nn

from .. import _lib, engine16
from ..engine import Saved, _empty
from ..params import Segment, _align

SITE_TXT_EMBED = 1001
SITE_TXT_LAYER = 1024


class TxtLayout:
  """Names / shapes / offsets of transformers' BertModel parameters inside one flat buffer: [small | big]."""

  def __init__(self, c):
    d, ff, L = c["hidden_size"], c["intermediate_size"], c["num_hidden_layers"]
    small, big = [], []
    small.append(("embeddings.position_embeddings.weight", (c["max_position_embeddings"], d)))
    small.append(("embeddings.token_type_embeddings.weight", (c["type_vocab_size"], d)))
    small.append(("embeddings.LayerNorm.weight", (d,)))
    small.append(("embeddings.LayerNorm.bias", (d,)))
    for l in range(L):
      p = "encoder.layer.%d." % l
      for n in ("query", "key", "value"):                      # contiguous -> [3d]
        small.append((p + "attention.self.%s.bias" % n, (d,)))
      small.append((p + "attention.output.dense.bias", (d,)))
      small.append((p + "attention.output.LayerNorm.weight", (d,)))
      small.append((p + "attention.output.LayerNorm.bias", (d,)))
      small.append((p + "intermediate.dense.bias", (ff,)))
      small.append((p + "output.dense.bias", (d,)))
      small.append((p + "output.LayerNorm.weight", (d,)))
      small.append((p + "output.LayerNorm.bias", (d,)))
    small.append(("pooler.dense.bias", (d,)))
    big.append(("embeddings.word_embeddings.weight", (c["vocab_size"], d)))
    for l in range(L):
      p = "encoder.layer.%d." % l
      for n in ("query", "key", "value"):                      # contiguous -> [3d, d]
        big.append((p + "attention.self.%s.weight" % n, (d, d)))
      big.append((p + "attention.output.dense.weight", (d, d)))
      big.append((p + "intermediate.dense.weight", (ff, d)))
      big.append((p + "output.dense.weight", (d, ff)))
    big.append(("pooler.dense.weight", (d, d)))
    self.segments = collections.OrderedDict()
    off = 0
    for name, shape in small:
      self.segments[name] = Segment(name, shape, off, True, False)
      off = _align(off + self.segments[name].numel)
    self.small_numel = off
    for name, shape in big:
      self.segments[name] = Segment(name, shape, off, False, False)
      off = _align(off + self.segments[name].numel)
    self.numel = off

  def off(self, name):
    return self.segments[name].offset

  def view(self, flat, name):
    s = self.segments[name]
    return flat[s.offset:s.offset + s.numel].view(s.shape)


class _Node(nn.Module):
  pass


class _TxtFn(torch.autograd.Function):
  """embeddings + L encoder layers; backward hand-written (engine16.layers_backward + txtops)."""

  @staticmethod
  def forward(ctx, anchor, net, ids, mask, training, seed):
    lib = _lib.load()
    st = _lib.stream_ptr()
    L = net.layout
    flat, w16 = net.flat, net.w16
    dt, f16 = w16.dt, w16.flat16
    R, W = ids.shape
    d = net.d
    p_hid = net.p_hidden if training else 0.0
    p_att = net.p_attn if training else 0.0
    rows = R * W
    e = "embeddings."
    h = _empty((rows, d), flat)
    h16 = torch.empty((rows, d), device=flat.device, dtype=_lib.torch_dtype(dt))
    sv = Saved()
    sv.mean0, sv.rstd0 = _empty((rows,), flat), _empty((rows,), flat)
    _lib.check(lib.mmt_txt_embed_ln_fwd(
        _lib.ptr(ids), _lib.ptr(flat, L.off(e + "word_embeddings.weight")), _lib.ptr(flat, L.off(e + "position_embeddings.weight")),
        _lib.ptr(flat, L.off(e + "token_type_embeddings.weight")), _lib.ptr(flat, L.off(e + "LayerNorm.weight")),
        _lib.ptr(flat, L.off(e + "LayerNorm.bias")), rows, W, net.vocab, d, net.eps, p_hid, seed, net.seed_ctr,
        SITE_TXT_EMBED, _lib.ptr(h), _lib.ptr(h16), _lib.ptr(sv.mean0), _lib.ptr(sv.rstd0), dt, st), "mmt_txt_embed_ln_fwd")
    h, sv.layers = engine16.layers_forward(net.spec, flat, f16, dt, h, h16, mask, R, W, p_hid, p_att, seed, net.seed_ctr)
    sv.R, sv.W, sv.seed, sv.p_hid, sv.p_att, sv.ids, sv.mask = R, W, seed, p_hid, p_att, ids, mask
    ctx.net, ctx.sv = net, sv
    return h.view(R, W, d)

  @staticmethod
  def backward(ctx, dh):
    net, sv = ctx.net, ctx.sv
    lib = _lib.load()
    st = _lib.stream_ptr()
    L = net.layout
    flat, w16 = net.flat, net.w16
    dt, f16 = w16.dt, w16.flat16
    R, W, d = sv.R, sv.W, net.d
    rows = R * W
    accumulate = any(p.grad is not None for p in net._hot_params())
    gflat = torch.empty_like(flat) if accumulate else net._grad_flat()
    gflat[:L.small_numel].zero_()                      # vectors / tables accumulated with atomics
    wseg = L.segments["embeddings.word_embeddings.weight"]
    train_word = net._param("embeddings.word_embeddings.weight").requires_grad
    if train_word:
      gflat[wseg.offset:wseg.offset + wseg.numel].zero_()
    sg = net.scale16
    dh_ = engine16.layers_backward(net.spec, flat, f16, gflat, dt, sv.layers, sv.mask, R, W,
                                   dh.contiguous().view(rows, d), sv.p_hid, sv.p_att, sv.seed, net.seed_ctr, sg, net)
    e = "embeddings."
    train_emb = net._param(e + "position_embeddings.weight").requires_grad
    _lib.check(lib.mmt_txt_embed_ln_bwd(
        _lib.ptr(dh_), _lib.ptr(sv.ids), _lib.ptr(flat, L.off(e + "word_embeddings.weight")),
        _lib.ptr(flat, L.off(e + "position_embeddings.weight")), _lib.ptr(flat, L.off(e + "token_type_embeddings.weight")),
        _lib.ptr(sv.mean0), _lib.ptr(sv.rstd0), _lib.ptr(flat, L.off(e + "LayerNorm.weight")), rows, W, net.vocab, d,
        sv.p_hid, sv.seed, net.seed_ctr, SITE_TXT_EMBED,
        _lib.ptr(gflat, wseg.offset) if train_word else None,
        _lib.ptr(gflat, L.off(e + "position_embeddings.weight")) if train_emb else None,
        _lib.ptr(gflat, L.off(e + "token_type_embeddings.weight")),
        _lib.ptr(gflat, L.off(e + "LayerNorm.weight")), _lib.ptr(gflat, L.off(e + "LayerNorm.bias")), st),
               "mmt_txt_embed_ln_bwd")
    net._publish_grads(gflat, accumulate)
    ctx.sv = None
    return None, None, None, None, None, None


class TxtBert(nn.Module):
  """transformers.BertModel's encoder on sm_100a kernels (see the module docstring)."""

  DEFAULTS = dict(vocab_size=28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                  intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
                  hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, hidden_act="gelu", initializer_range=0.02)

  def __init__(self, config=None, precision=None, **kw):
    super().__init__()
    c = dict(self.DEFAULTS)
    if config is not None:
      c.update({k: getattr(config, k) for k in self.DEFAULTS if hasattr(config, k)} if not isinstance(config, dict) else config)
    c.update(kw)
    if c["hidden_act"] != "gelu":
      raise NotImplementedError("TxtBert: only hidden_act='gelu' (erf)")
    if c["hidden_size"] % 128 != 0 or c["hidden_size"] // c["num_attention_heads"] not in (64, 128):
      raise NotImplementedError("TxtBert: hidden_size must be a multiple of 128 and the head dim 64 or 128")
    self.cfg_dict = c
    self.config = types.SimpleNamespace(**c)
    self.d, self.ff, self.H, self.L = c["hidden_size"], c["intermediate_size"], c["num_attention_heads"], c["num_hidden_layers"]
    self.vocab, self.eps = c["vocab_size"], float(c["layer_norm_eps"])
    self.p_hidden, self.p_attn = float(c["hidden_dropout_prob"]), float(c["attention_probs_dropout_prob"])
    self.precision = _lib.PREC_F16 if precision is None else precision
    if not _lib.is16(self.precision):
      raise NotImplementedError("TxtBert runs on the 16-bit operand path (PREC_F16 / PREC_BF16)")
    self.layout = TxtLayout(c)
    self.spec = engine16.EncSpec(self.layout, "encoder.layer.%d.", "LayerNorm", self.d, self.ff, self.H, self.L, self.eps,
                                 SITE_TXT_LAYER)
    flat = torch.zeros(self.layout.numel)
    std = c["initializer_range"]
    for name in self.layout.segments:
      v = self.layout.view(flat, name)
      if name.endswith("LayerNorm.weight"):
        v.fill_(1.0)
      elif name.endswith(".weight"):
        v.normal_(0.0, std)                                 # transformers BertPreTrainedModel._init_weights
    self._names = list(self.layout.segments)
    for name in self._names:
      mod, leaf = self._leaf(name)
      mod.register_parameter(leaf, nn.Parameter(self.layout.view(flat, name)))
    object.__setattr__(self, "flat", flat)
    self._gflat = None
    self._step = 0
    self.seed_ctr = None
    self.w16 = None
    self.scale16_override = None
    self._sync_device()

  # the engine16 helpers expect these on a "cfg"-like object
  @property
  def scale16(self):
    if self.scale16_override is not None:
      return float(self.scale16_override)
    return 65536.0 if self.precision == _lib.PREC_F16 else 1.0

  @classmethod
  def from_hf(cls, hf_model, precision=None):
    """Build from a live transformers BertModel (same geometry, weights copied by state_dict)."""
    net = cls(hf_model.config, precision=precision)
    sd = {k: v for k, v in hf_model.state_dict().items() if k in net.layout.segments}
    missing = set(net.layout.segments) - set(sd)
    if missing:
      raise RuntimeError("TxtBert.from_hf: the source model lacks %s" % sorted(missing)[:3])
    net.load_state_dict(sd, strict=True)
    return net

  def _leaf(self, dotted):
    parts = dotted.split(".")
    mod = self
    for p in parts[:-1]:
      if p not in mod._modules:
        mod.add_module(p, _Node())
      mod = mod._modules[p]
    return mod, parts[-1]

  def _param(self, name):
    mod, leaf = self._leaf(name)
    return mod._parameters[leaf]

  def _hot_params(self):
    cache = self.__dict__.get("_hot_cache")
    if cache is None or cache[0] is not self._param(self._names[0]):
      cache = [self._param(n) for n in self._names]
      self.__dict__["_hot_cache"] = cache
      self.__dict__["_grad_views"] = None
    return cache

  def _apply(self, fn, *a, **kw):
    out = super()._apply(fn, *a, **kw)
    self._sync_device()
    return out

  def _sync_device(self):
    first = self._param(self._names[0])
    flat = torch.zeros(self.layout.numel, device=first.device, dtype=torch.float32)
    with torch.no_grad():
      for n in self._names:
        p = self._param(n)
        v = self.layout.view(flat, n)
        v.copy_(p.data.to(torch.float32))
        p.data = v
        p.grad = None
    object.__setattr__(self, "flat", flat)
    self._gflat = None
    self.w16 = None

  def _grad_flat(self):
    if self._gflat is None or self._gflat.device != self.flat.device:
      self._gflat = torch.zeros_like(self.flat)
    return self._gflat

  def _publish_grads(self, gflat, accumulate):
    params = self._hot_params()
    views = self.__dict__.get("_grad_views")
    if views is None or views[0] != gflat.data_ptr():
      views = (gflat.data_ptr(), [None if n.startswith("pooler.") else self.layout.view(gflat, n) for n in self._names])
      if gflat is self._gflat:
        self.__dict__["_grad_views"] = views
    for p, v in zip(params, views[1]):
      if v is None or not p.requires_grad:
        continue
      if accumulate and p.grad is not None:
        p.grad.add_(v)
      else:
        p.grad = v if not accumulate else v.clone()

  def _prepare16(self):
    w = self.w16
    if w is None or w.flat16.device != self.flat.device or w.dt != _lib.dt_of(self.precision):
      w = self.w16 = _FlatCopy16(self.flat, _lib.dt_of(self.precision))
    sig = 0
    for p in self._hot_params():
      sig += p._version
    w.refresh(self.flat, sig)

  def forward(self, input_ids, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None, **kw):
    dev = self.flat.device
    if dev.type != "cuda":
      raise RuntimeError("mmt_b200.TxtBert needs a CUDA device (no CPU fallback); call .to('cuda')")
    if head_mask is not None:
      raise NotImplementedError("TxtBert: head_mask is not supported (the reference passes None)")
    R, W = input_ids.shape
    ids = input_ids.to(dev, torch.int32).contiguous()
    if attention_mask is None:
      mask = torch.ones((R, W), device=dev, dtype=torch.float32)
    else:
      mask = attention_mask.to(dev, torch.float32).contiguous()
    self._prepare16()
    self._step += 1
    seed = (torch.initial_seed() * 1000033 + self._step) & 0x7FFFFFFFFFFFFFFF
    anchor = next((p for p in self._hot_params() if p.requires_grad), None)
    if anchor is None or not torch.is_grad_enabled():
      with torch.no_grad():
        h = _TxtFn.apply(None, self, ids, mask, self.training, seed)
    else:
      h = _TxtFn.apply(anchor, self, ids, mask, self.training, seed)
    return (h,)


class _FlatCopy16:
  """16-bit copy of a flat fp32 parameter buffer (see engine16.Weights16)."""

  def __init__(self, flat, dt):
    self.dt = dt
    self.flat16 = torch.zeros(flat.numel(), device=flat.device, dtype=_lib.torch_dtype(dt))
    self.sig = None

  def refresh(self, flat, sig, force=False):
    if not force and self.sig is not None and self.sig == sig:
      return
    _lib.cast16(self.dt, flat, 1, flat.numel(), flat.numel(), self.flat16, flat.numel(), flat.numel())
    self.sig = sig

  def refresh_padded(self, *a):
    pass
