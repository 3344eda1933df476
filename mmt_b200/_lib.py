"""ctypes binding of libmmt_b200.so (the C ABI declared in include/mmt_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, this module
raises.  PyTorch is only the owner of device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmt_b200.so")

c_f = ctypes.c_float
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_u32 = ctypes.c_uint32
c_u64 = ctypes.c_uint64
c_p = ctypes.c_void_p

EPI_NONE, EPI_GELU, EPI_DGELU = 0, 1, 2
PREC_FP32, PREC_TF32, PREC_BF16, PREC_F16 = 0, 1, 2, 3     # engine-level precision of the train step
GEMM_SPLIT_K = 1
DT_F16, DT_BF16 = 0, 1
MAX_EXPERTS = 16


def is16(precision):
  return precision in (PREC_F16, PREC_BF16)


def dt_of(precision):
  return DT_BF16 if precision == PREC_BF16 else DT_F16


def torch_dtype(dt):
  return torch.bfloat16 if dt == DT_BF16 else torch.float16


class GemmDesc(ctypes.Structure):
  _fields_ = [
      ("M", c_i32), ("N", c_i32), ("K", c_i32),
      ("A", c_p), ("a_ms", c_i64), ("a_ks", c_i64), ("a_kb", c_i32), ("a_kbs", c_i64),
      ("B", c_p), ("b_ns", c_i64), ("b_ks", c_i64),
      ("C", c_p), ("c_ms", c_i64), ("c_mb", c_i32), ("c_mbs", c_i64),
      ("bias", c_p), ("add", c_p), ("aux", c_p),
      ("epilogue", c_i32), ("alpha", c_f),
      ("batch", c_i32), ("batch_inner", c_i32),
      ("a_bs0", c_i64), ("a_bs1", c_i64), ("b_bs0", c_i64), ("b_bs1", c_i64),
      ("c_bs0", c_i64), ("c_bs1", c_i64),
      ("bias_bs", c_i64),
      ("precision", c_i32),
      ("colsum", c_p), ("colsum_bs", c_i64),
      ("flags", c_i32),
  ]


class GemmDesc16(ctypes.Structure):
  _fields_ = [
      ("M", c_i32), ("N", c_i32), ("K", c_i32), ("dtype", c_i32),
      ("A", c_p), ("a_ld", c_i64), ("a_mn", c_i32),
      ("B", c_p), ("b_ld", c_i64), ("b_mn", c_i32),
      ("C32", c_p), ("c32_ld", c_i64),
      ("C16", c_p), ("c16_ld", c_i64), ("out16_scale", c_f),
      ("bias", c_p),
      ("add", c_p), ("add_ld", c_i64),
      ("aux16", c_p), ("aux_ld", c_i64),
      ("epilogue", c_i32), ("alpha", c_f),
      ("p_drop", c_f), ("site", c_u32), ("seed", c_u64), ("seed_ctr", c_p),
      ("batch", c_i32), ("batch_inner", c_i32),
      ("a_bs0", c_i64), ("a_bs1", c_i64), ("b_bs0", c_i64), ("b_bs1", c_i64),
      ("c_bs0", c_i64), ("c_bs1", c_i64), ("bias_bs", c_i64),
      ("colsum", c_p), ("colsum_scale", c_f), ("colsum_bs", c_i64),
      ("flags", c_i32),
  ]


class PackDesc(ctypes.Structure):
  _fields_ = [
      ("feats", c_p * MAX_EXPERTS), ("maxp", c_p * MAX_EXPERTS), ("out", c_p * MAX_EXPERTS),
      ("in_", c_i32 * MAX_EXPERTS), ("ld", c_i32 * MAX_EXPERTS),
      ("n", c_i32), ("B", c_i32), ("T", c_i32), ("dtype", c_i32),
  ]


# name -> (restype, argtypes); must list every symbol include/mmt_b200.h declares
SIGNATURES = {
    "mmt_version": (c_i32, []),
    "mmt_last_error": (c_i32, [ctypes.c_char_p, ctypes.c_size_t]),
    "mmt_launch_count": (c_i64, []),
    "mmt_set_step_counter": (c_i32, [c_p]),
    "mmt_gemm": (c_i32, [ctypes.POINTER(GemmDesc), c_p]),
    "mmt_colsum": (c_i32, [c_p, c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_i32, c_p]),
    "mmt_embed_ln_fwd": (c_i32, [c_p] * 8 + [c_i32] * 5 + [c_f, c_f, c_u64, c_u32] + [c_p] * 7 + [c_p]),
    "mmt_embed_ln_bwd": (c_i32, [c_p] * 10 + [c_i32] * 4 + [c_f, c_u64, c_u32] + [c_p] * 5 + [c_p]),
    "mmt_res_ln_fwd": (c_i32, [c_p] * 4 + [c_i64, c_i32, c_f, c_f, c_u64, c_u32] + [c_p] * 3 + [c_p]),
    "mmt_res_ln_bwd": (c_i32, [c_p] * 6 + [c_i64, c_i32, c_f, c_u64, c_u32] + [c_p] * 5 + [c_p]),
    "mmt_softmax_mask_fwd": (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_f, c_f, c_u64, c_u32, c_p, c_p, c_p]),
    "mmt_softmax_mask_bwd": (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_f, c_f, c_u64, c_u32, c_p]),
    "mmt_attention_fwd": (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_f, c_f, c_u64, c_u32, c_p, c_p, c_p, c_p,
                                  c_i32, c_p]),
    "mmt_cast_bf16": (c_i32, [c_p, c_p, c_i64, c_p]),
    "mmt_retrieval_ranks": (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p]),
    "mmt_readout_norm_fwd": (c_i32, [c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_p, c_p]),
    "mmt_readout_norm_bwd": (c_i32, [c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_p]),
    "mmt_geu_gate_fwd": (c_i32, [c_p] * 6 + [c_i32] * 4 + [c_f, c_f] + [c_p] * 6 + [c_p]),
    "mmt_geu_gate_bwd": (c_i32, [c_p] * 11 + [c_i32] * 4 + [c_p] * 4 + [c_p]),
    "mmt_dropout": (c_i32, [c_p, c_p, c_i64, c_i32, c_f, c_u64, c_u32, c_p]),
    "mmt_moe_softmax_fwd": (c_i32, [c_p, c_i32, c_i32, c_i32, c_p, c_p]),
    "mmt_moe_softmax_bwd": (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p]),
    "mmt_sims_combine_fwd": (c_i32, [c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_p]),
    "mmt_sims_combine_bwd": (c_i32, [c_p] * 4 + [c_i32] * 5 + [c_p, c_p, c_p]),
    "mmt_max_margin_fwd_bwd": (c_i32, [c_p, c_i32, c_f, c_i32, c_p, c_p, c_p, c_p]),
    "mmt_adam_step": (c_i32, [c_p] * 4 + [c_i64] + [c_f] * 5 + [c_i32, c_f, c_p]),
    # ---- 16-bit operand path ----
    "mmt_gemm16": (c_i32, [ctypes.POINTER(GemmDesc16), c_p]),
    "mmt_cast16": (c_i32, [c_p, c_i64, c_i32, c_i64, c_p, c_p, c_i32, c_i64, c_f, c_f, c_u64, c_p, c_u32, c_i32, c_p]),
    "mmt_pack_inputs16": (c_i32, [ctypes.POINTER(PackDesc), c_p]),
    "mmt_embed_ln16_fwd": (c_i32, [c_p] * 8 + [c_i32] * 5 + [c_f, c_f, c_u64, c_p, c_u32] + [c_p] * 8 + [c_i32, c_p]),
    "mmt_embed_ln16_bwd": (c_i32, [c_p] * 10 + [c_i32] * 4 + [c_f, c_u64, c_p, c_u32, c_p, c_p, c_f] + [c_p] * 4 +
                           [c_i32, c_p]),
    "mmt_ln16_fwd": (c_i32, [c_p] * 3 + [c_i64, c_i32, c_f] + [c_p] * 4 + [c_i32, c_p]),
    "mmt_ln16_bwd": (c_i32, [c_p] * 6 + [c_i64, c_i32, c_f, c_u64, c_p, c_u32, c_p, c_p, c_f] + [c_p] * 3 + [c_i32, c_p]),
    "mmt_attention16_fwd": (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_f, c_f, c_u64, c_p, c_u32, c_p, c_p, c_i32,
                                    c_p]),
    "mmt_attention16_bwd": (c_i32, [c_p] * 5 + [c_i32] * 4 + [c_f, c_f, c_u64, c_p, c_u32, c_f] + [c_p] * 4 +
                            [c_i32, c_p]),
    "mmt_adam16_step": (c_i32, [c_p] * 5 + [c_i64] + [c_f] * 5 + [c_i32, c_p, c_f, c_i32, c_p]),
    # ---- text encoder ----
    "mmt_txt_embed_ln_fwd": (c_i32, [c_p] * 6 + [c_i64, c_i32, c_i32, c_i32, c_f, c_f, c_u64, c_p, c_u32] + [c_p] * 4 +
                             [c_i32, c_p]),
    "mmt_txt_embed_ln_bwd": (c_i32, [c_p] * 8 + [c_i64, c_i32, c_i32, c_i32, c_f, c_u64, c_p, c_u32] + [c_p] * 5 + [c_p]),
    "mmt_txt_attention_fwd": (c_i32, [c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_f, c_f, c_u64, c_p, c_u32, c_p, c_i32, c_p]),
    "mmt_txt_attention_bwd": (c_i32, [c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_f, c_f, c_u64, c_p, c_u32, c_p, c_i32,
                                      c_p]),
    "mmt_colsum16": (c_i32, [c_p, c_i64, c_i32, c_i64, c_f, c_p, c_i32, c_p]),
}

_lib = None
_note = None      # optional callable(*tensors): test hook, see gemm16()


def load():
  """Loads libmmt_b200.so (building is __graft_entry__.build()'s job).  Raises if missing."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.isfile(LIB_PATH):
    raise RuntimeError(
        "mmt_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; "
        "g.build()'` (or `make -C mmt_b200/csrc`). There is no CPU fallback." % LIB_PATH)
  lib = ctypes.CDLL(LIB_PATH)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)      # AttributeError here = header / library mismatch
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


def last_error():
  buf = ctypes.create_string_buffer(512)
  load().mmt_last_error(buf, 512)
  return buf.value.decode(errors="replace")


def check(rc, what):
  if rc != 0:
    raise RuntimeError("mmt_b200 %s failed (code %d): %s" % (what, rc, last_error()))


def stream_ptr():
  """Raw cudaStream_t of torch's current stream on the current device (the C-level accessor: the python
  Stream object costs ~13 us to build, and a train step asks ~100 times)."""
  return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def ptr(t, offset=0):
  """Device pointer of a torch tensor (+ element offset); None -> NULL."""
  if t is None:
    return None
  return t.data_ptr() + offset * t.element_size()


def require_cuda(*tensors):
  for t in tensors:
    if t is not None and not t.is_cuda:
      raise RuntimeError("mmt_b200: expected CUDA tensors (no CPU fallback exists)")


def launch_count():
  return int(load().mmt_launch_count())


def gemm(M, N, K, A, a_ms, a_ks, B, b_ns, b_ks, C, c_ms, *, a_off=0, b_off=0, c_off=0, bias=None,
         bias_off=0, add=None, add_off=0, aux=None, aux_off=0, epilogue=EPI_NONE, alpha=1.0,
         a_kb=0, a_kbs=0, c_mb=0, c_mbs=0, batch=1, batch_inner=1, a_bs=(0, 0), b_bs=(0, 0),
         c_bs=(0, 0), bias_bs=0, precision=PREC_FP32, split_k=False, colsum=None, colsum_off=0,
         colsum_bs=0):
  d = GemmDesc()
  d.M, d.N, d.K = M, N, K
  d.A, d.a_ms, d.a_ks, d.a_kb, d.a_kbs = ptr(A, a_off), a_ms, a_ks, a_kb, a_kbs
  d.B, d.b_ns, d.b_ks = ptr(B, b_off), b_ns, b_ks
  d.C, d.c_ms, d.c_mb, d.c_mbs = ptr(C, c_off), c_ms, c_mb, c_mbs
  d.bias = ptr(bias, bias_off)
  d.add = ptr(add, add_off)
  d.aux = ptr(aux, aux_off)
  d.epilogue, d.alpha = epilogue, alpha
  d.batch, d.batch_inner = batch, batch_inner
  d.a_bs0, d.a_bs1 = a_bs
  d.b_bs0, d.b_bs1 = b_bs
  d.c_bs0, d.c_bs1 = c_bs
  d.bias_bs = bias_bs
  d.precision = precision
  d.flags = GEMM_SPLIT_K if split_k else 0
  d.colsum = ptr(colsum, colsum_off)
  d.colsum_bs = colsum_bs
  check(load().mmt_gemm(ctypes.byref(d), stream_ptr()), "mmt_gemm")


def gemm16(dt, M, N, K, A, a_ld, a_mn, B, b_ld, b_mn, *, a_off=0, b_off=0, C32=None, c32_off=0, c32_ld=0, C16=None,
           c16_off=0, c16_ld=0, out16_scale=1.0, bias=None, bias_off=0, add=None, add_off=0, add_ld=0, aux16=None,
           aux_off=0, aux_ld=0, epilogue=EPI_NONE, alpha=1.0, p_drop=0.0, seed=0, seed_ctr=None, site=0, batch=1,
           batch_inner=1, a_bs=(0, 0), b_bs=(0, 0), c_bs=(0, 0), bias_bs=0, colsum=None, colsum_off=0,
           colsum_scale=1.0, colsum_bs=0, split_k=False):
  """C = epilogue(alpha * A B^T) on 16-bit operands (mmt_gemm16, include/mmt_b200.h).  Offsets in elements.
  ~75 calls per train step: the descriptor is filled by ONE positional constructor call (field order of
  GemmDesc16._fields_) and pointers are computed inline."""
  if _note is not None:                     # test hook (tests/test_abi_and_host.py records the tensors a call touches)
    _note(A, B, C32, C16, bias, add, aux16, colsum)
  d = GemmDesc16(
      M, N, K, dt,
      A.data_ptr() + 2 * a_off, a_ld, a_mn,
      B.data_ptr() + 2 * b_off, b_ld, b_mn,
      None if C32 is None else C32.data_ptr() + 4 * c32_off, c32_ld,
      None if C16 is None else C16.data_ptr() + 2 * c16_off, c16_ld, out16_scale,
      None if bias is None else bias.data_ptr() + 4 * bias_off,
      None if add is None else add.data_ptr() + 4 * add_off, add_ld,
      None if aux16 is None else aux16.data_ptr() + 2 * aux_off, aux_ld,
      epilogue, alpha, p_drop, site, seed, seed_ctr, batch, batch_inner,
      a_bs[0], a_bs[1], b_bs[0], b_bs[1], c_bs[0], c_bs[1], bias_bs,
      None if colsum is None else colsum.data_ptr() + 4 * colsum_off, colsum_scale, colsum_bs,
      GEMM_SPLIT_K if split_k else 0)
  rc = load().mmt_gemm16(ctypes.byref(d), torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
                         if A.is_cuda else stream_ptr())
  if rc != 0:
    check(rc, "mmt_gemm16")


def cast16(dt, src, rows, cols, in_ld, out, out_cols, out_ld, scale=1.0, p_drop=0.0, seed=0, seed_ctr=None, site=0,
           src_off=0, out_off=0, out_lo=None):
  check(load().mmt_cast16(ptr(src, src_off), rows, cols, in_ld, ptr(out, out_off), ptr(out_lo), out_cols, out_ld, scale,
                          p_drop, seed, seed_ctr, site, dt, stream_ptr()), "mmt_cast16")
