"""Flat parameter storage for the MMT hot path.

All hot-path parameters live in ONE fp32 buffer (and their gradients in a second one of the same
layout), so that (a) the kernels see concatenated operands where the math wants them -- Q|K|V
weights as one [3d, d] matrix, the M text GatedEmbeddingUnit `fc` weights as one [M*d, text_dim]
matrix, BatchNorm vectors as [M*d] -- (b) the data-parallel gradient exchange is a single NCCL
all-reduce, and (c) Adam is one kernel over the whole buffer.  The nn.Module exposes views of this
buffer under the reference's parameter names so reference checkpoints load.

Layout: [small | big].  "small" (biases, LayerNorm / BatchNorm affine, embedding tables, moe
weights) is the region whose gradients are accumulated with atomics and must be zeroed each step;
"big" matrices are written by the weight-gradient GEMMs.
"""
import collections

import torch


class Segment:
  __slots__ = ("name", "shape", "offset", "numel", "small", "head")

  def __init__(self, name, shape, offset, small, head):
    self.name, self.shape, self.offset = name, tuple(shape), offset
    n = 1
    for s in shape:
      n *= s
    self.numel, self.small, self.head = n, small, head


def _align(n, a=8):   # 8 elements: 16-byte aligned in the fp32 buffers AND in their 16-bit copy
  return (n + a - 1) // a * a


class Layout:
  """Names/shapes/offsets of every hot-path parameter (reference names, Appendix B)."""

  def __init__(self, expert_dims, vid_bert_params, text_dim, same_dim):
    self.mods = list(expert_dims.keys())
    self.expert_dims = expert_dims
    self.d = d = same_dim
    self.ff = ff = vid_bert_params["intermediate_size"]
    self.L = L = vid_bert_params["num_hidden_layers"]
    self.text_dim = td = text_dim
    self.max_pos = vid_bert_params["max_position_embeddings"]
    self.type_vocab = vid_bert_params["type_vocab_size"]
    M = len(self.mods)
    small, big = [], []
    # ---- small region (order matters only for the concatenated groups) ----
    for m in self.mods:
      small.append(("video_dim_reduce.%s.fc.bias" % m, (d,), False))
    small.append(("vid_bert.embeddings.position_embeddings.weight", (self.max_pos, d), False))
    small.append(("vid_bert.embeddings.token_type_embeddings.weight", (self.type_vocab, d), False))
    small.append(("vid_bert.embeddings.layer_norm.weight", (d,), False))
    small.append(("vid_bert.embeddings.layer_norm.bias", (d,), False))
    for l in range(L):
      p = "vid_bert.encoder.layer.%d." % l
      for n in ("query", "key", "value"):                      # contiguous -> [3d]
        small.append((p + "attention.self.%s.bias" % n, (d,), False))
      small.append((p + "attention.output.dense.bias", (d,), False))
      small.append((p + "attention.output.layer_norm.weight", (d,), False))
      small.append((p + "attention.output.layer_norm.bias", (d,), False))
      small.append((p + "intermediate.dense.bias", (ff,), False))
      small.append((p + "output.dense.bias", (d,), False))
      small.append((p + "output.layer_norm.weight", (d,), False))
      small.append((p + "output.layer_norm.bias", (d,), False))
    small.append(("vid_bert.pooler.dense.bias", (d,), False))
    for grp in ("fc.bias", "cg.fc.bias", "cg.batch_norm.weight", "cg.batch_norm.bias"):
      for m in self.mods:                                      # contiguous -> [M*d]
        small.append(("text_GU.%s.%s" % (m, grp), (d,), True))
    for m in self.mods:                                        # contiguous -> [M, td]
      small.append(("moe_fc_txt.%s.weight" % m, (1, td), True))
    for m in self.mods:                                        # contiguous -> [M]
      small.append(("moe_fc_txt.%s.bias" % m, (1,), True))
    # ---- big region ----
    for m in self.mods:
      big.append(("video_dim_reduce.%s.fc.weight" % m, (d, expert_dims[m]["dim"]), False))
    for l in range(L):
      p = "vid_bert.encoder.layer.%d." % l
      for n in ("query", "key", "value"):                      # contiguous -> [3d, d]
        big.append((p + "attention.self.%s.weight" % n, (d, d), False))
      big.append((p + "attention.output.dense.weight", (d, d), False))
      big.append((p + "intermediate.dense.weight", (ff, d), False))
      big.append((p + "output.dense.weight", (d, ff), False))
    big.append(("vid_bert.pooler.dense.weight", (d, d), False))
    for m in self.mods:                                        # contiguous -> [M*d, td]
      big.append(("text_GU.%s.fc.weight" % m, (d, td), True))
    for m in self.mods:                                        # contiguous -> [M, d, d]
      big.append(("text_GU.%s.cg.fc.weight" % m, (d, d), True))

    self.segments = collections.OrderedDict()
    off = 0
    contiguous_groups = ("moe_fc_txt",)     # groups that must stay densely packed (no padding)
    for name, shape, head in small:
      self.segments[name] = Segment(name, shape, off, True, head)
      off += self.segments[name].numel
      if not name.startswith(contiguous_groups):
        off = _align(off)
    off = _align(off)
    self.small_numel = off
    for name, shape, head in big:
      self.segments[name] = Segment(name, shape, off, False, head)
      off = _align(off + self.segments[name].numel)
    self.numel = off
    # buffers (BatchNorm running statistics), separate flat buffer: [M*d] mean | [M*d] var
    self.buffers = collections.OrderedDict()
    boff = 0
    for grp in ("running_mean", "running_var"):
      for m in self.mods:
        n = "text_GU.%s.cg.batch_norm.%s" % (m, grp)
        self.buffers[n] = Segment(n, (d,), boff, True, True)
        boff += d
    self.buffers_numel = boff

  def off(self, name):
    return self.segments[name].offset

  def no_grad_ranges(self):
    """Weight matrices that never receive a gradient (the pooler, reference model.py:583-584)."""
    seg = self.segments["vid_bert.pooler.dense.weight"]
    return [(seg.offset, _align(seg.numel))]

  def layer_big_range(self, l):
    """(offset, numel) of encoder layer l's weight matrices (Wq|Wk|Wv, Wo, W1, W2: one contiguous block)."""
    p = "vid_bert.encoder.layer.%d." % l
    first, last = self.segments[p + "attention.self.query.weight"], self.segments[p + "output.dense.weight"]
    return first.offset, last.offset + last.numel - first.offset

  def view(self, flat, name):
    s = self.segments[name]
    return flat[s.offset:s.offset + s.numel].view(s.shape)
