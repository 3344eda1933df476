"""2-GPU data-parallel parity (NCCL): rank-local batches through CENet.enable_data_parallel() must
reproduce the single-device result at the global batch (conf matrix, loss, every gradient, BN
running statistics).  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, precision="fp32", trainable_text=False):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  sys.path.insert(0, root)
  sys.path.insert(0, os.path.join(root, "tests"))
  import torch.distributed as dist
  import mmt_test_helpers as H
  from mmt_b200.model.loss import MaxMarginRankingLoss
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  torch.manual_seed(0)
  Bg = 8
  ed, vb, P, batch, cfg = H.make_case(["s3d", "vggish", "ocr"], Bg, 9, layers=2)
  crit = MaxMarginRankingLoss(0.05, True)
  # single-device reference at the global batch
  def txt():
    return H.TxtEmb() if trainable_text else None

  ref = H.build_cuda_net(ed, vb, P, batch, device=dev, precision=precision, txt_bert=txt()).train()
  conf_ref = ref(**H.batch_kwargs(batch, dev))["cross_view_conf_matrix"]
  loss_ref = crit(conf_ref)
  loss_ref.backward()
  # sharded run
  bl = Bg // world
  sl = slice(rank * bl, (rank + 1) * bl)
  local = {k: ({m: v[sl] for m, v in batch[k].items()} if isinstance(batch[k], dict) else batch[k][sl])
           for k in batch}
  net = H.build_cuda_net(ed, vb, P, local, device=dev, precision=precision, txt_bert=txt()).train()
  net.enable_data_parallel()
  conf = net(**H.batch_kwargs(local, dev))["cross_view_conf_matrix"]
  loss = crit(conf)
  loss.backward()
  torch.cuda.synchronize()
  errs = {"conf": H.rel_err(conf, conf_ref), "loss": abs(float(loss) - float(loss_ref))}
  gmax = max(float(p.grad.abs().max()) for p in ref._hot_params() if p.grad is not None)
  worst = 0.0
  for n in ref._names:
    g, gr = net._param(n).grad, ref._param(n).grad
    if gr is None:
      continue
    worst = max(worst, float((g - gr).abs().max()) / max(float(gr.abs().max()), 1e-3 * gmax))
  errs["grad"] = worst
  errs["bn"] = H.rel_err(net.buf_flat, ref.buf_flat)
  errs["txt_grad"] = 0.0
  if trainable_text:      # parameters outside the flat buffer: per-rank partial gradients must have been summed
    errs["txt_grad"] = H.rel_err(net.txt_bert.emb.weight.grad, ref.txt_bert.emb.weight.grad)
  q.put((rank, errs))
  dist.destroy_process_group()


@pytest.mark.parametrize("precision,trainable_text", [("fp32", False), ("f16", False), ("f16", True)])
def test_two_gpu_data_parallel_matches_single_device_global_batch(precision, trainable_text):
  if torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs")
  import torch.multiprocessing as mp
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q, precision, trainable_text)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=300) for _ in range(2)]
  for p in procs:
    p.join(timeout=60)
  print("2-GPU data parallel vs single device (%s, trainable text %s): %s" % (precision, trainable_text, res))
  for rank, e in res:
    assert e["conf"] < 1e-5 and e["loss"] < 1e-6 and e["grad"] < 2e-4 and e["bn"] < 1e-5 and e["txt_grad"] < 1e-4, (rank, e)
