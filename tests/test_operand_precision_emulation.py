"""Which operand precisions can hold the 1e-3 parity bar?  CPU emulation on the oracle: every Linear /
matmul operand of the hot path is rounded to the candidate format before an fp32-accumulated product
(what a tensor-core MMA with fp32 accumulation computes), the confusion matrix is compared with the
unrounded fp32 oracle.  This pins the numerical rationale of DESIGN.md: tf32 TRUNCATION (what
`kind::tf32` does to fp32 operands) needs the bias compensation; round-to-nearest 10-bit-mantissa
operands -- tf32-RN or fp16, the latter at half the bytes -- sit at ~6e-4; bf16 is ~5x outside."""
import torch
import torch.nn.functional as F

import mmt_test_helpers as H
from oracle import mmt_oracle as O


def _round(x, mode):
  if mode == "bf16":
    return x.to(torch.bfloat16).to(torch.float32)
  if mode == "fp16":
    return x.to(torch.float16).to(torch.float32)
  if mode == "tf32_trunc":
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
  if mode == "tf32_rn":
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
  raise ValueError(mode)


def _conf_error(mode, P, batch, cfg, ref):
  lin0, mm0 = F.linear, torch.matmul
  O.F.linear = lambda x, w, b=None: lin0(_round(x, mode), _round(w, mode), b)
  O.torch.matmul = lambda a, b: mm0(_round(a, mode), _round(b, mode))
  try:
    out = O.cenet_forward(P, batch, cfg, training=False, out="conf", text_feat=batch["text_feat"])
  finally:
    O.F.linear, O.torch.matmul = lin0, mm0
  return H.rel_err(out["cross_view_conf_matrix"], ref)


def test_operand_precision_classes_against_the_parity_bar():
  mods = ["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"]
  ed, vb, P, batch, cfg = H.make_case(mods, 8, 30, layers=4)
  with torch.no_grad():
    ref = O.cenet_forward(P, batch, cfg, training=False, out="conf", text_feat=batch["text_feat"])["cross_view_conf_matrix"]
    err = {m: _conf_error(m, P, batch, cfg, ref) for m in ("tf32_trunc", "tf32_rn", "fp16", "bf16")}
  print("emulated operand rounding, conf max-rel:", {k: "%.2e" % v for k, v in err.items()})
  assert err["tf32_trunc"] > 1e-3            # why the tensor-core epilogues compensate the truncation bias
  assert err["tf32_rn"] < 1e-3 and err["fp16"] < 1e-3
  assert err["bf16"] > 2e-3                  # 8-bit mantissa: its own tolerance class
