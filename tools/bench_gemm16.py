"""Times mmt_gemm16 on the train step's GEMM shapes with CUDA events (rotating operand buffers, so operands do
not sit in L2 from the previous launch).  Development aid; prints us and TFLOP/s per shape / variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmt_b200 import _lib

dev = torch.device("cuda")
_lib.load()
DT = 0
tdt = torch.float16


def bench(name, M, N, K, a_mn=0, b_mn=0, reps=5, iters=20, dbg=0, **kw):
  A = [torch.randn((K, M) if a_mn else (M, K), device=dev).to(tdt) for _ in range(reps)]
  B = [(torch.randn((K, N) if b_mn else (N, K), device=dev) * 0.05).to(tdt) for _ in range(reps)]
  out32 = [torch.empty(M, N, device=dev) for _ in range(reps)] if kw.get("c32") else None
  pitch = kw.get("pitch", N)               # row pitch of the 16-bit outputs (elements); "inter": aux in the same rows
  out16 = [torch.empty(M, pitch, device=dev, dtype=tdt) for _ in range(reps)] if kw.get("c16") else None
  if kw.get("epi") and kw.get("inter"):
    aux = None
  else:
    aux = [torch.randn(M, pitch, device=dev).to(tdt) for _ in range(reps)] if kw.get("epi") else None
  add = [torch.randn(M, N, device=dev) for _ in range(reps)] if kw.get("add") else None
  bias = torch.randn(N, device=dev)
  cs = torch.zeros(N, device=dev) if kw.get("colsum") else None

  def launch(i):
    import ctypes
    d = _lib.GemmDesc16()
    d.M, d.N, d.K, d.dtype = M, N, K, DT
    d.A, d.a_ld, d.a_mn = A[i].data_ptr(), (M if a_mn else K), a_mn
    d.B, d.b_ld, d.b_mn = B[i].data_ptr(), (N if b_mn else K), b_mn
    if out32: d.C32, d.c32_ld = out32[i].data_ptr(), N
    if out16: d.C16, d.c16_ld, d.out16_scale = out16[i].data_ptr(), pitch, 1.0
    if not kw.get("split"): d.bias = bias.data_ptr()
    if add: d.add, d.add_ld = add[i].data_ptr(), N
    if aux: d.aux16, d.aux_ld = aux[i].data_ptr(), pitch
    elif kw.get("epi"): d.aux16, d.aux_ld = out16[i].data_ptr() + 2 * N, pitch
    d.epilogue = kw.get("epi", 0)
    d.alpha = 1.0
    d.p_drop, d.seed, d.site = kw.get("p", 0.0), 5, 3
    d.batch = d.batch_inner = 1
    if cs is not None: d.colsum, d.colsum_scale = cs.data_ptr(), 1.0
    d.flags = dbg | (1 if kw.get("split") else 0)
    _lib.check(_lib.load().mmt_gemm16(ctypes.byref(d), _lib.stream_ptr()), "gemm16")

  for i in range(3):
    launch(i % reps)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(iters):
    launch(i % reps)
  e1.record()
  torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 1e3 / iters
  print("%-34s M=%5d N=%4d K=%4d  %7.1f us  %7.1f TFLOP/s" % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6), flush=True)


BS = 13952
if __name__ == "__main__":
  which = sys.argv[1] if len(sys.argv) > 1 else "all"
  if which in ("all", "shapes"):
    bench("QKV fwd (C16)", BS, 1536, 512, c16=1)
    bench("O-proj (drop+add, C32)", BS, 512, 512, c32=1, add=1, p=0.1)
    bench("FFN-up GELU (aux16+C16)", BS, 3072, 512, c16=1, epi=1)
    bench("FFN-down (drop+add, C32)", BS, 512, 3072, c32=1, add=1, p=0.1)
    bench("du DGELU (C16+colsum)", BS, 3072, 512, b_mn=1, c16=1, epi=2, colsum=1)
    bench("da dgrad (C32)", BS, 512, 3072, b_mn=1, c32=1)
    bench("dctx dgrad (C16)", BS, 512, 512, b_mn=1, c16=1)
    bench("dh dgrad (add, C32)", BS, 512, 1536, b_mn=1, c32=1, add=1)
    bench("dW2 wgrad split-K", 512, 3072, BS, a_mn=1, b_mn=1, c32=1, split=1)
    bench("dW1 wgrad split-K", 3072, 512, BS, a_mn=1, b_mn=1, c32=1, split=1)
    bench("dWqkv wgrad split-K", 1536, 512, BS, a_mn=1, b_mn=1, c32=1, split=1)
    bench("dWo wgrad split-K", 512, 512, BS, a_mn=1, b_mn=1, c32=1, split=1)
  if which == "pitch":
    bench("plain C16 pitch N", BS, 3072, 512, c16=1)
    bench("plain C16 pitch 2N", BS, 3072, 512, c16=1, pitch=6144)
    bench("GELU separate tensors", BS, 3072, 512, c16=1, epi=1)
    bench("GELU separate, pitch 2N", BS, 3072, 512, c16=1, epi=1, pitch=6144)
    bench("GELU interleaved rows [f|g]", BS, 3072, 512, c16=1, epi=1, pitch=6144, inter=1)
    bench("plain C32", BS, 3072, 512, c32=1)
    bench("N=1536 plain C16", BS, 1536, 512, c16=1)
    bench("N=1536 plain C16 pitch 2N", BS, 1536, 512, c16=1, pitch=3072)
  if which == "ffnup":
    bench("FFN-up GELU (aux16+C16)", BS, 3072, 512, c16=1, epi=1, iters=3)
    bench("QKV fwd (C16)", BS, 1536, 512, c16=1, iters=3)
  if which in ("all", "exp"):
    for K in (64, 256, 512, 1024, 2048):
      bench("N=3072 plain C16, K sweep", BS, 3072, K, c16=1)
    bench("N=3072 K=512 C16 no stores", BS, 3072, 512, c16=1, dbg=256)
    bench("N=3072 K=512 C16 no epilogue", BS, 3072, 512, c16=1, dbg=512)
    bench("N=3072 K=512 C32", BS, 3072, 512, c32=1)
    bench("N=3072 K=512 GELU", BS, 3072, 512, c16=1, epi=1)
    bench("N=3072 K=512 GELU no stores", BS, 3072, 512, c16=1, epi=1, dbg=256)
    bench("N=512 K=512 C16", BS, 512, 512, c16=1)
    bench("N=512 K=512 C16 no epilogue", BS, 512, 512, c16=1, dbg=512)
    bench("N=512 K=3072 C32 no epilogue", BS, 512, 3072, c32=1, dbg=512)
