"""Data-parallel train step: one process per GPU, NCCL over NVLink/NVSwitch.

The reference only has a (non-functional for >1 GPU) nn.DataParallel wrapper
(base/base_trainer.py:49-50, trainer/trainer.py:182-199: gather embeddings, one global
similarity + loss).  Here rank r owns samples [r*B/W, (r+1)*B/W); there are exactly two exchange
steps per train step:

  forward : ONE all-gather of the per-rank head inputs -- video expert embeddings [B/W, M, d] and
            text features [B/W, text_dim] -- after which every rank evaluates the (tiny) text head,
            similarity matrix and loss on the GLOBAL batch, so BatchNorm statistics and the loss
            are exactly those of the single-device reference at batch B (no SyncBN collective);
  backward: all-reduce(SUM) of the flat gradient buffer, issued in a few contiguous pieces as
            they become final -- the head's runs right after the head backward, each encoder
            layer's weight-gradient block when that layer's backward has been enqueued, the rest
            (biases, LayerNorm, embeddings, ReduceDim) at the end -- so that NCCL runs on its own
            stream underneath the remaining backward kernels.  Encoder gradients are per-rank
            partial sums of the global-mean loss; head gradients are identical on every rank and
            are pre-scaled by 1/W (exact for W a power of two) so the same SUM leaves them intact.
"""
import torch
import torch.distributed as dist

from . import engine


def _all_gather_rows(x, group):
  w = dist.get_world_size(group)
  out = torch.empty((w * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
  dist.all_gather_into_tensor(out, x.contiguous(), group=group)
  return out


def allreduce_grads(params, group=None):
  """All-reduce (SUM) the .grad of `params` as ONE flattened collective (parameters outside the flat buffer)."""
  grads = [p.grad for p in params if p.requires_grad and p.grad is not None]
  if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
    return
  flat = torch.cat([g.reshape(-1) for g in grads])
  dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  off = 0
  for g in grads:
    n = g.numel()
    g.copy_(flat[off:off + n].view_as(g))
    off += n


def head_segments(layout):
  """[(offset, numel)] contiguous runs of the flat buffer that belong to the text head."""
  runs = []
  for seg in layout.segments.values():
    if not seg.head:
      continue
    if runs and runs[-1][0] + runs[-1][1] >= seg.offset:
      runs[-1] = (runs[-1][0], seg.offset + seg.numel - runs[-1][0])
    else:
      runs.append((seg.offset, seg.numel))
  return runs


class GradReducer:
  """Asynchronous all-reduce of disjoint pieces of one flat gradient buffer.  `reduce(off, n)` may be
  called as soon as gflat[off:off+n] is final on the current stream; `finish()` reduces whatever was
  not covered and makes the current stream wait for every piece."""

  def __init__(self, gflat, group):
    self.gflat, self.group, self.works, self.done = gflat, group, [], []

  def reduce(self, off, n):
    if n <= 0:
      return
    self.works.append(dist.all_reduce(self.gflat[off:off + n], op=dist.ReduceOp.SUM, group=self.group,
                                      async_op=True))
    self.done.append((off, n))

  def skip(self, off, n):
    """Exclude a piece from the exchange altogether (parameters that never receive a gradient)."""
    if n > 0:
      self.done.append((off, n))

  def finish(self):
    pos = 0
    for off, n in sorted(self.done) + [(self.gflat.numel(), 0)]:
      if off < pos:
        raise RuntimeError("GradReducer: overlapping pieces at offset %d" % off)
      self.reduce(pos, off - pos)
      pos = off + n
    for w in self.works:
      w.wait()
    self.works = []


class DPEncodeFn(torch.autograd.Function):
  """EncodeFn with the embedding all-gather inside (forward) and the gradient all-reduce
  (backward).  Returns GLOBAL-batch (vid, txt, tw)."""

  @staticmethod
  def forward(ctx, anchor, text, net, feats, maxp, ft, ind, training, seed, group):
    rank = dist.get_rank(group)
    vid_l, sv_v = engine.video_forward(net.cfg, net.flat, feats, maxp, ft, ind, training,
                                       seed + 7919 * (rank + 1))        # per-rank dropout masks
    vid = _all_gather_rows(vid_l, group)
    text_g = _all_gather_rows(text, group)
    # identical seed on every rank: the head runs redundantly on the global batch
    txt, tw, sv_h = engine.head_forward(net.cfg, net.flat, net.buf_flat, text_g, training, seed)
    ctx.net, ctx.sv, ctx.group = net, (sv_v, sv_h), group
    ctx.local_rows = (vid_l.shape[0], text.shape[0])
    return vid, txt, tw

  @staticmethod
  def backward(ctx, dvid, dtxt, dtw):
    net, (sv_v, sv_h), group = ctx.net, ctx.sv, ctx.group
    w, rank = dist.get_world_size(group), dist.get_rank(group)
    bl, rl = ctx.local_rows
    accumulate = any(p.grad is not None for p in net._hot_params())
    gflat = torch.empty_like(net.flat) if accumulate else net._grad_flat()
    engine.zero_small_grads(net.cfg, gflat)
    dtext_g = engine.head_backward(net.cfg, net.flat, gflat, sv_h, dtxt.contiguous(),
                                   dtw.contiguous(), need_dtext=ctx.needs_input_grad[1])
    red = GradReducer(gflat, group)
    small_end = getattr(net.layout, "small_numel", 0)
    for off, n in head_segments(net.layout):
      gflat[off:off + n].mul_(1.0 / w)
      if off >= small_end:            # the head's weight matrices go out now; its few small vectors stay
        red.reduce(off, n)            # with the rest of the small region (one piece at the end)
    for off, n in getattr(net.layout, "no_grad_ranges", lambda: [])():
      red.skip(off, n)                # e.g. the unused pooler: nothing to exchange
    layer_range = getattr(net.layout, "layer_big_range", None)
    engine.video_backward(net.cfg, net.flat, gflat, sv_v, dvid[rank * bl:(rank + 1) * bl].contiguous(),
                          on_layer_done=(lambda l: red.reduce(*layer_range(l))) if layer_range else None)
    red.finish()
    net._publish_grads(gflat, accumulate)
    ctx.sv = None
    dtext = dtext_g[rank * rl:(rank + 1) * rl] if dtext_g is not None else None
    if dtext is not None:
      # every trainable parameter OUTSIDE the flat buffer (the text encoder: `txt_agg=bertftn` trains it) receives
      # a per-rank partial gradient from this rank's slice of d loss / d text: sum them over the group once the
      # whole backward pass has run (the same end-of-backward callback torch's DistributedDataParallel uses)
      torch.autograd.Variable._execution_engine.queue_callback(lambda: net.allreduce_outside_grads(group))
    return None, dtext, None, None, None, None, None, None, None, None
