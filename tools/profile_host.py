"""Host-side cost of one eager train step (python + ctypes + torch allocator), GPU running asynchronously."""
import cProfile, io, os, pstats, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mmt_test_helpers as H
from mmt_b200.model.loss import MaxMarginRankingLoss
from mmt_b200.optim import FusedAdam
mods = ["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"]
ed, vb, P, batch, cfg = H.make_case(mods, 64, 30, layers=4, dropout=0.1)
net = H.build_cuda_net(ed, vb, P, batch, dropout=0.1, precision="tf32").train()
crit, opt = MaxMarginRankingLoss(0.05, True), FusedAdam(net, lr=5e-5)
kw = H.batch_kwargs(batch, "cuda")
def step():
  opt.zero_grad()
  out = net(**kw)
  loss = crit(out["cross_view_conf_matrix"])
  loss.backward()
  opt.step()
  return loss
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms/step, with drain %.2f ms/step" % ((t1 - t0) * 50, (t2 - t0) * 50))
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
