"""Times the fused attention forward and backward (mmt_attention16_fwd / _bwd: delta + main + dq-finish kernels) of one
encoder layer at the benchmark shape, CUDA events, rotating buffers (4 sets > L2).

  python tools/bench_attention16.py            # forward, backward
  MMT_ATT_BWD_DEBUG=1|2|4 ...                  # backward with a phase switched off (timing experiments, wrong results)
  MMT_ATT_BWD_WARPS=16 ...                     # 16 softmax warps per CTA
"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmt_b200 import _lib
lib = _lib.load()
ptr, check = _lib.ptr, _lib.check
B, H, S, dh = 64, 4, 218, 128
d = H * dh
dev = "cuda"
NSET = 4
g = torch.Generator().manual_seed(0)
sets = []
for i in range(NSET):
  qkv16 = (torch.randn(B * S, 3 * d, generator=g) * 0.5).to(dev).half()
  dctx16 = (torch.randn(B * S, d, generator=g) * 0.05).to(dev).half()
  sets.append(dict(qkv16=qkv16, dctx16=dctx16, ctx16=torch.empty(B * S, d, device=dev, dtype=torch.half),
                   lse=torch.empty(B, H, S, device=dev), dqkv16=torch.empty(B * S, 3 * d, device=dev, dtype=torch.half),
                   dq32=torch.zeros(B * S, d, device=dev), delta=torch.empty(B, H, S, device=dev),
                   dbias=torch.zeros(3 * d, device=dev)))
mask = torch.ones(B, S, device=dev)
scale = 1 / math.sqrt(dh)
dt = _lib.DT_F16
st = _lib.stream_ptr()
def fwd(s):
  check(lib.mmt_attention16_fwd(ptr(s["qkv16"]), ptr(mask), B, H, S, dh, scale, 0.1, 1, None, 7, ptr(s["ctx16"]), ptr(s["lse"]), dt, st), "fwd")
def bwd(s):
  check(lib.mmt_attention16_bwd(ptr(s["qkv16"]), ptr(s["ctx16"]), ptr(s["dctx16"]), ptr(s["lse"]), ptr(mask), B, H, S, dh, scale, 0.1, 1, None, 7,
                                1.0, ptr(s["dqkv16"]), ptr(s["dq32"]), ptr(s["delta"]), ptr(s["dbias"]), dt, st), "bwd")
for name, fn in (("attention16_fwd", fwd), ("attention16_bwd (3 kernels)", bwd)):
  for s in sets: fn(s)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(40): fn(sets[i % NSET])
  e1.record(); torch.cuda.synchronize()
  print("%-28s %7.1f us   (debug=%s warps=%s)" % (name, e0.elapsed_time(e1) * 25, os.environ.get("MMT_ATT_BWD_DEBUG", "0"),
                                                 os.environ.get("MMT_ATT_BWD_WARPS", "8")))
