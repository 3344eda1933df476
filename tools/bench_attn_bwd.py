"""Times the attention-backward sequence of one encoder layer (saved P / Pd): dP = dctx V^T, dV = Pd^T dctx,
softmax backward, dQ = dS K, dK = dS^T Q -- the batched tcgen05 GEMMs + the row kernel, CUDA events."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmt_b200 import _lib
lib = _lib.load()
B, H, S, dh = 64, 4, 218, 128
d, Sp = H * dh, 220
dev = "cuda"
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B * S, 3 * d, generator=g).to(dev)
dctx = torch.randn(B * S, d, generator=g).to(dev)
P = torch.softmax(torch.randn(B, H, S, Sp, generator=g), -1).to(dev)
Pd = P.clone()
dP = torch.empty(B, H, S, Sp, device=dev)
dqkv = torch.empty(B * S, 3 * d, device=dev)
bsP, bsQ = (H * S * Sp, S * Sp), (S * 3 * d, dh)
T = _lib.PREC_TF32
def g_dP(): _lib.gemm(S, S, dh, dctx, d, 1, qkv, 3 * d, 1, dP, Sp, b_off=2 * d, batch=B * H, batch_inner=H, a_bs=(S * d, dh), b_bs=bsQ, c_bs=bsP, precision=T)
def g_dV(): _lib.gemm(S, dh, S, Pd, 1, Sp, dctx, 1, d, dqkv, 3 * d, c_off=2 * d, batch=B * H, batch_inner=H, a_bs=bsP, b_bs=(S * d, dh), c_bs=bsQ, precision=T)
def sm_bwd(): _lib.check(lib.mmt_softmax_mask_bwd(_lib.ptr(dP), _lib.ptr(P), B, H, S, Sp, 1 / math.sqrt(dh), 0.1, 1, 2, _lib.stream_ptr()), "sm")
def g_dQ(): _lib.gemm(S, dh, S, dP, Sp, 1, qkv, 1, 3 * d, dqkv, 3 * d, b_off=d, batch=B * H, batch_inner=H, a_bs=bsP, b_bs=bsQ, c_bs=bsQ, precision=T)
def g_dK(): _lib.gemm(S, dh, S, dP, 1, Sp, qkv, 1, 3 * d, dqkv, 3 * d, c_off=d, batch=B * H, batch_inner=H, a_bs=bsP, b_bs=bsQ, c_bs=bsQ, precision=T)
for name, fn in (("dP", g_dP), ("dV", g_dV), ("softmax_bwd", sm_bwd), ("dQ", g_dQ), ("dK", g_dK)):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10): fn()
  e1.record(); torch.cuda.synchronize()
  print("%-12s %7.1f us" % (name, e0.elapsed_time(e1) * 100))
