"""Times mmt_gemm (tf32 path) on the encoder's GEMM shapes with CUDA events; prints us and TFLOP/s.
Usage: python tools/bench_gemm.py [M]   (env switches MMT_PAIR_BN / MMT_PAIR_SPLIT2 select tile policies)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmt_b200 import _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 13952
dev = "cuda"
g = torch.Generator().manual_seed(0)
# (name, N, K, epilogue, has_add, b_transposed)
CASES = [("qkv_fwd", 1536, 512, _lib.EPI_NONE, False), ("oproj_fwd", 512, 512, _lib.EPI_NONE, True),
         ("ffn_up_gelu", 3072, 512, _lib.EPI_GELU, False), ("ffn_down", 512, 3072, _lib.EPI_NONE, True),
         ("dgelu_dgrad", 3072, 512, _lib.EPI_DGELU, False), ("ffn_up_dgrad", 512, 3072, _lib.EPI_NONE, False),
         ("qkv_dgrad", 512, 1536, _lib.EPI_NONE, False)]
for name, N, K, epi, has_add in CASES:
  A = torch.randn(M, K, generator=g).to(dev)
  W = torch.randn(N, K, generator=g).to(dev) * 0.05
  C = torch.empty(M, N, device=dev)
  bias = torch.randn(N, generator=g).to(dev)
  add = torch.randn(M, N, generator=g).to(dev) if has_add else None
  aux = torch.randn(M, N, generator=g).to(dev) if epi != _lib.EPI_NONE else None
  def run():
    _lib.gemm(M, N, K, A, K, 1, W, K, 1, C, N, bias=bias, add=add, aux=aux, epilogue=epi,
              precision=_lib.PREC_TF32)
  for _ in range(3): run()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  iters = 20
  e0.record()
  for _ in range(iters): run()
  e1.record(); torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 1e3 / iters
  ref = (A @ W.t() + bias + (add if add is not None else 0)) if epi == _lib.EPI_NONE else None
  err = float((C - ref).abs().max() / ref.abs().max()) if ref is not None else float("nan")
  print("%-14s M=%d N=%4d K=%4d  %7.1f us  %6.1f TFLOP/s  err %.1e" % (name, M, N, K, us, 2.0 * M * N * K / us * 1e-6, err))
  if os.environ.get("MMT_BENCH_BF16", "1") != "0":           # experimental 16-bit operand mode, same shape
    Ab, Wb = A.to(torch.bfloat16), W.to(torch.bfloat16)
    def run16():
      _lib.gemm(M, N, K, Ab, K, 1, Wb, K, 1, C, N, bias=bias, add=add, aux=aux, epilogue=epi, precision=_lib.PREC_BF16)
    for _ in range(3): run16()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters): run16()
    e1.record(); torch.cuda.synchronize()
    us16 = e0.elapsed_time(e1) * 1e3 / iters
    ref16 = (Ab.double() @ Wb.double().t() + bias.double() + (add.double() if add is not None else 0)) if epi == _lib.EPI_NONE else None
    err16 = float((C.double() - ref16).abs().max() / ref16.abs().max()) if ref16 is not None else float("nan")
    print("%-14s   bf16 operands     %7.1f us  %6.1f TFLOP/s  err %.1e (vs fp64 product of the rounded operands)" %
          ("", us16, 2.0 * M * N * K / us16 * 1e-6, err16))
