"""world_size-2 gloo test (CPU) of the data-parallel exchange logic in mmt_b200/parallel.py:
all-gather of the head inputs, redundant global head, local slice of d(vid), 1/W pre-scaling of the
head gradients and the single flat all-reduce must reproduce the single-process full-batch
gradients.  The CUDA kernels are replaced by a small torch stand-in with the same call contract
(video part per-sample, head part coupled across the batch like BatchNorm)."""
import os
import socket
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

D_IN, D, TD, M = 6, 4, 5, 3
B_LOCAL, WORLD = 3, 2


class _Seg:
  def __init__(self, offset, numel, head):
    self.offset, self.numel, self.head = offset, numel, head


def _layout():
  segs = {"wv": _Seg(0, D_IN * M * D, False), "wt": _Seg(D_IN * M * D, TD * M * D, True),
          "wm": _Seg(D_IN * M * D + TD * M * D, TD * M, True)}
  # "layer 0" = the video weights, so the early per-layer all-reduce path is exercised
  L = types.SimpleNamespace(segments=segs, numel=D_IN * M * D + TD * M * D + TD * M,
                            layer_big_range=lambda l: (0, D_IN * M * D))
  return L


def _views(flat):
  L = _layout()
  s = L.segments
  wv = flat[s["wv"].offset:s["wv"].offset + s["wv"].numel].view(D_IN, M * D)
  wt = flat[s["wt"].offset:s["wt"].offset + s["wt"].numel].view(TD, M * D)
  wm = flat[s["wm"].offset:s["wm"].offset + s["wm"].numel].view(TD, M)
  return wv, wt, wm


def _video(flat, x):
  wv, _, _ = _views(flat)
  return torch.tanh(x @ wv).view(-1, M, D)


def _head(flat, text):
  _, wt, wm = _views(flat)
  c = text - text.mean(0, keepdim=True)                 # couples the batch like BatchNorm
  return (c @ wt).view(-1, M, D), torch.softmax(text @ wm, -1)


def _loss(vid, txt, tw):
  sims = torch.einsum("imd,jmd->ij", txt * tw[:, :, None], vid)
  return (sims ** 2).mean() + sims.diag().sum()


class _StubEngine:
  """Same call contract as mmt_b200.engine, torch math instead of kernels."""

  @staticmethod
  def video_forward(cfg, flat, feats, maxp, ft, ind, training, seed):
    return _video(flat, feats).detach(), feats

  @staticmethod
  def head_forward(cfg, flat, bufs, text, training, seed):
    txt, tw = _head(flat, text)
    return txt.detach(), tw.detach(), text

  @staticmethod
  def zero_small_grads(cfg, gflat):
    gflat.zero_()

  @staticmethod
  def head_backward(cfg, flat, gflat, sv, dtxt, dtw, need_dtext=True):
    with torch.enable_grad():
      f = flat.detach().clone().requires_grad_(True)
      t = sv.detach().clone().requires_grad_(True)
      txt, tw = _head(f, t)
      gf, gt = torch.autograd.grad([txt, tw], [f, t], [dtxt, dtw])
    gflat += gf
    return gt

  @staticmethod
  def video_backward(cfg, flat, gflat, sv, dvid, on_layer_done=None):
    with torch.enable_grad():
      f = flat.detach().clone().requires_grad_(True)
      gf, = torch.autograd.grad(_video(f, sv), f, dvid)
    nv = D_IN * M * D                      # only the video weights: the head runs are already in flight
    gflat[:nv] += gf[:nv]
    if on_layer_done is not None:
      on_layer_done(0)


def _worker(rank, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=WORLD)
  from mmt_b200 import parallel
  parallel.engine = _StubEngine
  g = torch.Generator().manual_seed(0)
  L = _layout()
  flat = torch.randn(L.numel, generator=g) * 0.3
  x_all = torch.randn(WORLD * B_LOCAL, D_IN, generator=g)
  t_all = torch.randn(WORLD * B_LOCAL, TD, generator=g)
  gflat = torch.zeros_like(flat)
  published = {}
  # a trainable "text encoder" OUTSIDE the flat buffer: its gradient is a per-rank partial sum that the
  # end-of-backward callback must all-reduce (every published config trains txt_bert)
  wenc = (torch.eye(TD) + 0.1 * torch.randn(TD, TD, generator=g)).requires_grad_(True)
  net = types.SimpleNamespace(
      cfg=None, flat=flat, buf_flat=None, layout=L, _hot_params=lambda: [],
      _grad_flat=lambda: gflat, _publish_grads=lambda gf, acc: published.update(g=gf.clone()),
      allreduce_outside_grads=lambda group: parallel.allreduce_grads([wenc], group))
  sl = slice(rank * B_LOCAL, (rank + 1) * B_LOCAL)
  raw = t_all[sl].clone().requires_grad_(True)
  text = raw @ wenc
  anchor = torch.zeros(1, requires_grad=True)
  vid, txt, tw = parallel.DPEncodeFn.apply(anchor, text, net, x_all[sl], None, None, None, True, 5,
                                           None)
  assert vid.shape[0] == WORLD * B_LOCAL and txt.shape[0] == WORLD * B_LOCAL
  loss = _loss(vid, txt, tw)
  loss.backward()
  # single-process reference on the full batch
  f = flat.clone().requires_grad_(True)
  tr = t_all.clone().requires_grad_(True)
  wr = wenc.detach().clone().requires_grad_(True)
  txt_r, tw_r = _head(f, tr @ wr)
  loss_r = _loss(_video(f, x_all), txt_r, tw_r)
  loss_r.backward()
  ok = (torch.allclose(published["g"], f.grad, rtol=1e-5, atol=1e-6) and
        torch.allclose(raw.grad, tr.grad[sl], rtol=1e-5, atol=1e-6) and
        torch.allclose(wenc.grad, wr.grad, rtol=1e-5, atol=1e-6) and
        abs(float(loss) - float(loss_r)) < 1e-6)
  q.put((rank, bool(ok), float((published["g"] - f.grad).abs().max())))
  dist.destroy_process_group()


def test_data_parallel_exchange_matches_full_batch_gradients():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
  for p in procs:
    p.start()
  res = [q.get(timeout=60) for _ in range(WORLD)]
  for p in procs:
    p.join(timeout=60)
  assert all(ok for _, ok, _ in res), res


def test_head_segments_are_contiguous_runs():
  from mmt_b200.parallel import head_segments
  from mmt_b200.params import Layout
  from oracle import mmt_oracle as O
  ed = O.compute_dims(["s3d", "vggish", "ocr"])
  vb = {"hidden_size": 128, "num_hidden_layers": 1, "num_attention_heads": 4, "intermediate_size": 256,
        "max_position_embeddings": 32, "type_vocab_size": 19}
  L = Layout(ed, vb, 96, 128)
  runs = head_segments(L)
  assert len(runs) == 2                               # one run in the small region, one in the big
  covered = sum(n for _, n in runs)
  assert covered >= sum(s.numel for s in L.segments.values() if s.head)
  for s in L.segments.values():
    inside = any(off <= s.offset and s.offset + s.numel <= off + n for off, n in runs)
    assert inside == s.head, s.name
