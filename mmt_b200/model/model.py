"""Drop-in `model.model` surface of gabeur/mmt backed by the sm_100a kernels.

Mirrors the reference's plugin interface for the hot path (reference model/model.py):
  * `CENet(**arch_args, expert_dims=, tokenizer=)`            model.py:48-73   (same kwargs)
  * `CENet.forward(token_ids, features, features_t, features_ind, features_avgpool,
                   features_maxpool, query_masks, out='conf', device=None, debug=None)`
                                                               model.py:312-322 (same returns)
  * `sharded_cross_view_inner_product(vid_embds, text_embds, vid_weights, text_weights,
                                      subspaces, merge_caption_similiarities='avg')`  model.py:789-794
Parameter names / shapes are the reference's, so its checkpoints load with
`load_state_dict`.  Only the branch every published config selects is implemented
(vid_cont=bert, vid_inp=both, pos_enc=tint, out_tok=mxp, vid_wgh=none, txt_wgh=emb, txt_pro=gbn,
txt_agg=bert*, keep_missing_modalities=true); other values raise NotImplementedError.

Deliberately dropped reference side effects: forward() does not overwrite the caller's `features`
/ `features_t` dicts (model.py:437, 518-520) and does not store `self.device`.

There is no CPU fallback: tensors are moved to the module's CUDA device and the CUDA library must
be present (mmt_b200/_lib.py raises otherwise).
"""
import collections
import math
import re

import torch
import torch.nn as nn

from .. import _lib, engine
from ..params import Layout

_SUPPORTED = dict(vid_cont="bert", vid_inp="both", pos_enc="tint", out_tok="mxp", vid_wgh="none",
                  txt_wgh="emb", txt_pro="gbn")


def _build_txt_bert(txt_bert_params):
  """bert-base-cased text encoder (third-party, outside the named hot path), built exactly as the
  reference does (model.py:161: `TxtBertModel.from_pretrained('bert-base-cased', **txt_bert_params)`): a missing
  or corrupt checkpoint FAILS.  Only with MMT_ALLOW_RANDOM_TXT_BERT=1 (benches / tests on machines without the
  weights) a random-init model of the same geometry is substituted, loudly (vocab 28996 is what the reference
  hard-codes, model.py:205)."""
  import os
  import warnings
  from transformers import BertConfig, BertModel
  kw = dict(txt_bert_params or {})
  try:
    return BertModel.from_pretrained("bert-base-cased", **kw)
  except Exception as e:
    if os.environ.get("MMT_ALLOW_RANDOM_TXT_BERT", "0") != "1":
      raise RuntimeError("mmt_b200.CENet: could not load the pretrained 'bert-base-cased' text encoder (%s). "
                         "Training / evaluating on random text-encoder weights is never done silently; set "
                         "MMT_ALLOW_RANDOM_TXT_BERT=1 to allow a random-init model of the same geometry, or pass "
                         "txt_bert=<module>." % (e,)) from e
    warnings.warn("mmt_b200.CENet: 'bert-base-cased' weights unavailable; using a RANDOM-INIT text encoder "
                  "(MMT_ALLOW_RANDOM_TXT_BERT=1)")
    return BertModel(BertConfig(vocab_size=28996, **kw))


class EncodeFn(torch.autograd.Function):
  """video encoder + text head; backward is hand-written (engine.encode_backward)."""

  @staticmethod
  def forward(ctx, anchor, text, net, feats, maxp, ft, ind, training, seed):
    vid, txt, tw, sv = engine.encode_forward(net.cfg, net.flat, net.buf_flat, text, feats, maxp,
                                             ft, ind, training, seed)
    ctx.net, ctx.sv = net, sv
    return vid, txt, tw

  @staticmethod
  def backward(ctx, dvid, dtxt, dtw):
    net, sv = ctx.net, ctx.sv
    accumulate = any(p.grad is not None for p in net._hot_params())
    gflat = torch.empty_like(net.flat) if accumulate else net._grad_flat()
    dtext = engine.encode_backward(net.cfg, net.flat, gflat, sv, dvid.contiguous(),
                                   dtxt.contiguous(), dtw.contiguous(),
                                   need_dtext=ctx.needs_input_grad[1])
    net._publish_grads(gflat, accumulate)
    ctx.sv = None
    return None, dtext, None, None, None, None, None, None, None


class SimsFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, vid, txt, vw, tw, caps, merge_avg, bwd_precision=engine.PREC_FP32, scale16=1.0):
    vid, txt, vw, tw = vid.contiguous(), txt.contiguous(), vw.contiguous(), tw.contiguous()
    training = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])     # a gradient will be asked for
    sims, dots = engine.sims_forward(vid, txt, vw, tw, caps, merge_avg, bwd_precision if training else None)
    ctx.save_for_backward(vid, txt, vw, tw, dots)
    ctx.caps, ctx.merge_avg, ctx.bwd_precision, ctx.scale16 = caps, merge_avg, bwd_precision, scale16
    return sims

  @staticmethod
  def backward(ctx, dsims):
    vid, txt, vw, tw, dots = ctx.saved_tensors
    dvid, dtxt, dtw = engine.sims_backward(dsims.contiguous(), dots, vid, txt, vw, tw, ctx.caps,
                                           ctx.merge_avg, ctx.bwd_precision, ctx.scale16)
    return dvid, dtxt, None, dtw, None, None, None, None


def sharded_cross_view_inner_product(vid_embds, text_embds, vid_weights, text_weights, subspaces,
                                     merge_caption_similiarities="avg"):
  """Reference model/model.py:789-837.  vid_embds[mod] [b,d], text_embds[mod] [b,caps,d]."""
  if merge_caption_similiarities not in ("avg", "indep"):
    raise ValueError("unrecognised merge mode: {}".format(merge_caption_similiarities))
  first = vid_embds[subspaces[0]]
  in_dev = first.device
  dev = in_dev if in_dev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
  b = first.size(0)
  caps = text_embds[subspaces[0]].size(1)
  vid = torch.stack([vid_embds[m].to(dev, torch.float32) for m in subspaces], 1)            # [b,M,d]
  txt = torch.stack([text_embds[m].to(dev, torch.float32).reshape(b * caps, -1) for m in subspaces], 1)
  vw = vid_weights.to(dev, torch.float32).reshape(b, -1)
  tw = text_weights.to(dev, torch.float32).reshape(b * caps, -1)
  sims = SimsFn.apply(vid, txt, vw, tw, caps, merge_caption_similiarities == "avg")
  return sims.to(in_dev)


class _Node(nn.Module):
  """Bare container so that parameters get the reference's dotted state_dict names."""
  pass


class CENet(nn.Module):
  """Whole cross-modal architecture (reference model/model.py:45), hot path on sm_100a kernels."""

  def __init__(self, l2renorm, expert_dims, tokenizer, keep_missing_modalities, test_caption_mode,
               freeze_weights=False, mimic_ce_dims=False, concat_experts=False,
               concat_mix_experts=False, use_experts="origfeat", txt_inp=None, txt_agg=None,
               txt_pro=None, txt_wgh=None, vid_inp=None, vid_cont=None, vid_wgh=None, pos_enc=None,
               out_tok=None, use_mask="nomask", same_dim=512, vid_bert_params=None,
               txt_bert_params=None, agg_dims=None, normalize_experts=True, txt_bert=None):
    super().__init__()
    got = dict(vid_cont=vid_cont, vid_inp=vid_inp, pos_enc=pos_enc, out_tok=out_tok,
               vid_wgh=vid_wgh, txt_wgh=txt_wgh, txt_pro=txt_pro)
    for k, v in _SUPPORTED.items():
      if got[k] != v:
        raise NotImplementedError("mmt_b200.CENet: %s=%r is not on the published hot path "
                                  "(only %r)" % (k, got[k], v))
    if not keep_missing_modalities or not normalize_experts or txt_agg is None or \
        txt_agg[:4] != "bert" or mimic_ce_dims or concat_experts or concat_mix_experts:
      raise NotImplementedError("mmt_b200.CENet: unsupported option combination")
    self.modalities = list(expert_dims.keys())
    self.expert_dims = expert_dims
    self.test_caption_mode = test_caption_mode
    self.keep_missing_modalities = keep_missing_modalities
    self.l2renorm = l2renorm
    self.same_dim = same_dim
    self.txt_inp, self.txt_agg, self.txt_pro, self.txt_wgh = txt_inp, txt_agg, txt_pro, txt_wgh
    self.vid_inp, self.vid_cont, self.vid_wgh = vid_inp, vid_cont, vid_wgh
    self.pos_enc, self.out_tok = pos_enc, out_tok
    self.vid_bert_params = vid_bert_params
    self.normalize_experts = normalize_experts
    if vid_bert_params["hidden_size"] != same_dim:
      raise ValueError("vid_bert hidden_size must equal same_dim")
    if vid_bert_params.get("hidden_act", "gelu") != "gelu":
      raise NotImplementedError("only hidden_act='gelu' (erf) is implemented")

    # ---- text encoder (model.py:133-193) ----
    z = re.match(r"bert([a-z]{3})(\d*)(\D*)", txt_agg)
    assert z
    state, freeze_until = z.groups()[0], z.groups()[1]
    self.post_agg = z.groups()[2] if (z.groups()[2] and z.groups()[2] != "cls") else "cls"
    if txt_bert_params is None:
      dout = vid_bert_params["hidden_dropout_prob"]
      txt_bert_params = {"hidden_dropout_prob": dout, "attention_probs_dropout_prob": dout}
    if txt_bert is None:
      txt_bert = _build_txt_bert(txt_bert_params)
      import os
      if os.environ.get("MMT_TXT_BERT", "native") == "native":
        # the same weights on this repo's kernels (mmt_b200/model/txt_bert.py); MMT_TXT_BERT=hf keeps transformers' module
        from .txt_bert import TxtBert
        txt_bert = TxtBert.from_hf(txt_bert)
    self.txt_bert = txt_bert
    if state == "frz":
      for name, param in self.txt_bert.named_parameters():
        parts = name.split(".")
        if parts[0] != "encoder":
          continue
        if freeze_until:
          if len(parts) > 2 and parts[2].isdigit() and int(parts[2]) < int(freeze_until):
            param.requires_grad = False
        else:
          param.requires_grad = False
    if txt_inp == "bertfrz" and hasattr(self.txt_bert, "embeddings"):
      for param in self.txt_bert.embeddings.parameters():
        param.requires_grad = False
    text_dim = self.txt_bert.config.hidden_size

    # ---- hot-path parameters: one flat buffer, reference names as views ----
    self.layout = Layout(expert_dims, vid_bert_params, text_dim, same_dim)
    type_idx = [expert_dims[m]["idx"] for m in self.modalities]
    self.cfg = engine.Config(self.layout, vid_bert_params, type_idx,
                             txt_bert_params["hidden_dropout_prob"])
    flat = torch.zeros(self.layout.numel)
    self._init_flat(flat, vid_bert_params.get("initializer_range", 0.02))
    self._register_views(flat)
    buf = torch.zeros(self.layout.buffers_numel)
    buf[self.cfg.M * same_dim:] = 1.0                       # running_var = 1
    self._register_buffer_views(buf)
    self._step = 0
    self._gflat = None
    self.dp_group = None
    self._dp = False
    self._sync_device()

  # ------------------------------------------------------------------ parameter plumbing
  def _init_flat(self, flat, std):
    L = self.layout
    for name, seg in L.segments.items():
      v = L.view(flat, name)
      if name.startswith("vid_bert."):
        if "layer_norm.weight" in name:
          v.fill_(1.0)
        elif name.endswith(".weight"):
          v.normal_(0.0, std)                               # bert.py:361-369
      elif "batch_norm.weight" in name:
        v.fill_(1.0)
      elif name.endswith("batch_norm.bias"):
        pass
      elif name.endswith(".weight"):                        # nn.Linear default init
        nn.init.kaiming_uniform_(v, a=math.sqrt(5))
      elif name.endswith(".bias"):
        fan_in = L.segments[name[:-4] + "weight"].shape[1]
        bound = 1.0 / math.sqrt(fan_in)
        v.uniform_(-bound, bound)

  def _leaf(self, dotted):
    parts = dotted.split(".")
    mod = self
    for p in parts[:-1]:
      if p not in mod._modules:
        mod.add_module(p, _Node())
      mod = mod._modules[p]
    return mod, parts[-1]

  def _register_views(self, flat):
    self._names = []
    for name in self.layout.segments:
      mod, leaf = self._leaf(name)
      mod.register_parameter(leaf, nn.Parameter(self.layout.view(flat, name)))
      self._names.append(name)
    object.__setattr__(self, "flat", flat)

  def _register_buffer_views(self, buf):
    for name, seg in self.layout.buffers.items():
      mod, leaf = self._leaf(name)
      mod.register_buffer(leaf, buf[seg.offset:seg.offset + seg.numel])
    nbt = torch.zeros(len(self.modalities), dtype=torch.long)
    for i, m in enumerate(self.modalities):
      mod, leaf = self._leaf("text_GU.%s.cg.batch_norm.num_batches_tracked" % m)
      mod.register_buffer(leaf, nbt[i])
    object.__setattr__(self, "buf_flat", buf)
    object.__setattr__(self, "nbt_flat", nbt)

  def _param(self, name):
    mod, leaf = self._leaf(name)
    return mod._parameters[leaf]

  def _hot_params(self):
    """The Parameters that live in the flat buffer, in layout order.  The list is cached (the objects
    survive .to(): _sync_device only re-points their .data) and re-validated by identity."""
    cache = self.__dict__.get("_hot_cache")
    if cache is None or len(cache) != len(self._names) or cache[0] is not self._param(self._names[0]):
      cache = [self._param(n) for n in self._names]
      self.__dict__["_hot_cache"] = cache
      self.__dict__["_grad_views"] = None
    return cache

  def _apply(self, fn, *a, **kw):
    out = super()._apply(fn, *a, **kw)
    self._sync_device()
    return out

  def _sync_device(self):
    """After .to()/.cuda() every Parameter owns fresh storage: re-pack them into one flat buffer
    and re-point the Parameters (and BN buffers) at views of it."""
    first = self._param(self._names[0])
    dev = first.device
    flat = torch.empty(self.layout.numel, device=dev, dtype=torch.float32)
    flat.zero_()
    with torch.no_grad():
      for n in self._names:
        p = self._param(n)
        v = self.layout.view(flat, n)
        v.copy_(p.data.to(torch.float32))
        p.data = v
        p.grad = None
      buf = torch.empty(self.layout.buffers_numel, device=dev, dtype=torch.float32)
      for n, seg in self.layout.buffers.items():
        mod, leaf = self._leaf(n)
        v = buf[seg.offset:seg.offset + seg.numel]
        v.copy_(mod._buffers[leaf].to(torch.float32))
        mod._buffers[leaf] = v
      nbt = torch.empty(len(self.modalities), device=dev, dtype=torch.long)
      for i, m in enumerate(self.modalities):
        mod, leaf = self._leaf("text_GU.%s.cg.batch_norm.num_batches_tracked" % m)
        nbt[i] = mod._buffers[leaf].to(dev)
        mod._buffers[leaf] = nbt[i]
      object.__setattr__(self, "nbt_flat", nbt)
    object.__setattr__(self, "flat", flat)
    object.__setattr__(self, "buf_flat", buf)
    self._gflat = None
    self.cfg.w16 = None
    self.cfg.type_idx_dev = torch.tensor(self.cfg.type_idx, dtype=torch.int32, device=dev)

  def _prepare16(self):
    """16-bit operand modes: make sure the 16-bit weight copy exists and is current.  FusedAdam refreshes
    it inside its kernel; after any other change of the parameters (another optimizer, load_state_dict,
    .to()) torch has bumped the flat buffer's version counter and one cast pass runs here."""
    cfg = self.cfg
    if not _lib.is16(cfg.precision):
      return
    from ..engine16 import Weights16
    w = cfg.w16
    if w is None or w.flat16.device != self.flat.device or w.dt != _lib.dt_of(cfg.precision):
      w = cfg.w16 = Weights16(cfg, self.flat)
    sig = 0
    for p in self._hot_params():
      sig += p._version
    w.refresh(cfg, self.flat, sig)

  def enable_data_parallel(self, group=None):
    """Shard the train step by batch over the ranks of `group` (mmt_b200/parallel.py): in
    training mode forward() then takes the LOCAL batch and returns the GLOBAL confusion matrix /
    embeddings; backward all-reduces the flat gradient."""
    import torch.distributed as dist
    if not dist.is_initialized():
      raise RuntimeError("enable_data_parallel needs an initialised torch.distributed group")
    self.dp_group = group
    self._dp = dist.get_world_size(group) > 1

  @property
  def w16(self):
    """The 16-bit weight copy of the hot-path parameters (None in the fp32 / tf32 modes)."""
    return self.cfg.w16 if _lib.is16(self.cfg.precision) else None

  def allreduce_outside_grads(self, group=None):
    """Data-parallel step: all-reduce (SUM) the gradients of the trainable parameters that do not live in the flat
    buffer -- the text encoder -- as ONE flattened NCCL call.  The loss is the global-batch mean and each rank
    back-propagated only its own rows of d loss / d text, so the sum is the single-device gradient."""
    import torch.distributed as dist
    from ..parallel import allreduce_grads
    hot = set(id(p) for p in self._hot_params())
    tb = self.txt_bert
    if hasattr(tb, "_gflat") and tb._gflat is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
      # a flat-buffer text encoder (TxtBert): its parameters' gradients are views of ONE buffer -> one in-place call
      if any(p.grad is not None for p in tb._hot_params()):
        dist.all_reduce(tb._gflat, op=dist.ReduceOp.SUM, group=group)
      hot |= set(id(p) for p in tb._hot_params())
    allreduce_grads([p for p in self.parameters() if id(p) not in hot], group)

  def _grad_flat(self):
    if self._gflat is None or self._gflat.device != self.flat.device:
      self._gflat = torch.zeros_like(self.flat)
    return self._gflat

  def _publish_grads(self, gflat, accumulate):
    """Expose the flat gradient as per-parameter .grad views (reference names).  The pooler's
    parameters get no gradient, as in the reference (its output is discarded, model.py:583-584)."""
    params = self._hot_params()
    views = self.__dict__.get("_grad_views")
    if views is None or views[0] != gflat.data_ptr():        # per-parameter views of this gradient buffer
      views = (gflat.data_ptr(), [None if n.startswith("vid_bert.pooler.") else self.layout.view(gflat, n)
                                  for n in self._names])
      if gflat is self._gflat:
        self.__dict__["_grad_views"] = views
    for p, v in zip(params, views[1]):
      if v is None or not p.requires_grad:
        continue
      if accumulate and p.grad is not None:
        p.grad.add_(v)
      else:
        p.grad = v if not accumulate else v.clone()

  # ------------------------------------------------------------------ forward
  def _text_features(self, token_ids, dev):
    """model.py:350-387: token_ids [b,caps,W,2] -> txt_bert -> [b*caps, text_dim]."""
    b, caps, w, fd = token_ids.size()
    tok = token_ids.view(b * caps, w, fd).to(dev)
    input_ids = tok[:, :, 0].to(torch.long)
    attention_mask = tok[:, :, 1].to(torch.long)
    token_type_ids = torch.zeros_like(input_ids)
    position_ids = torch.arange(w, device=dev, dtype=torch.long).unsqueeze(0).expand(b * caps, w)
    out = self.txt_bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids,
                        position_ids=position_ids, head_mask=None)
    last = out[0]
    if self.post_agg == "cls":
      return last[:, 0]
    if self.post_agg == "mxp":
      return torch.max(last[:, 1:], 1)[0]
    return torch.mean(last[:, 1:], 1)

  def forward(self, token_ids, features, features_t, features_ind, features_avgpool,
              features_maxpool, query_masks, out="conf", device=None, debug=None):
    dev = self.flat.device
    if dev.type != "cuda":
      raise RuntimeError("mmt_b200.CENet needs a CUDA device (no CPU fallback); call .to('cuda')")
    mods = self.modalities
    b, caps = token_ids.size(0), token_ids.size(1)
    if self.training and b * caps == 1 and not self._dp:
      # torch.nn.BatchNorm1d (text_GU.*.cg.batch_norm) refuses a single row in training mode
      raise ValueError("Expected more than 1 value per channel when training, got input size "
                       "torch.Size([1, %d])" % self.same_dim)
    text = self._text_features(token_ids, dev).to(torch.float32).contiguous()

    def prep(x):
      return x.to(dev, torch.float32).contiguous()

    feats = [prep(features[m]) for m in mods]
    maxp = [prep(features_maxpool[m]) for m in mods]
    ft = torch.stack([prep(features_t[m]) for m in mods], 0)
    ind = torch.stack([prep(features_ind[m]) for m in mods], 0)

    self._prepare16()
    self._step += 1
    seed = (torch.initial_seed() * 1000003 + self._step) & 0x7FFFFFFFFFFFFFFF
    anchor = next((p for p in self._hot_params() if p.requires_grad), None)
    if self._dp and self.training:
      from ..parallel import DPEncodeFn
      vid, txt, tw = DPEncodeFn.apply(anchor, text, self, feats, maxp, ft, ind, True, seed,
                                      self.dp_group)
      b = vid.shape[0]                                   # global batch from here on
    else:
      vid, txt, tw = EncodeFn.apply(anchor, text, self, feats, maxp, ft, ind, self.training, seed)
    if self.training:
      self.nbt_flat += 1            # BatchNorm1d.num_batches_tracked of all M text GEUs

    M = len(mods)
    # model.py:593-607: vid_wgh='none' -> ones, L1-normalised (no availability masking)
    vid_weights = torch.full((b, M), 1.0 / M, device=dev, dtype=torch.float32)
    text_weights = tw.view(b, caps, M)
    merge = "avg" if self.training else self.test_caption_mode
    self.merge_caption_similarities = merge
    if out == "conf":
      conf = SimsFn.apply(vid, txt, vid_weights, tw, caps, merge == "avg", self.cfg.precision, self.cfg.scale16)
      return {"modalities": mods, "cross_view_conf_matrix": conf}
    return {
        "vid_embds": vid,                                             # [b, M, d]
        "text_embds": txt.view(b, caps, M, -1).permute(0, 2, 1, 3),   # [b, M, caps, d]
        "vid_weights": vid_weights,
        "text_weights": text_weights,
    }
