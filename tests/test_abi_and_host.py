"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol the header
declares with matching arity, and the host-side kernel sequencing (mmt_b200/engine.py) addresses
only memory it owns (dry run against a recording stub of the library -- no compute)."""
import ctypes
import os
import re

import pytest
import torch

from mmt_b200 import _lib, engine, engine16
from oracle import mmt_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
  src = open(os.path.join(ROOT, "include", "mmt_b200.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  out = {}
  for m in re.finditer(r"\b(?:int|int64_t)\s+(mmt_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
    args = m.group(2).strip()
    n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    out[m.group(1)] = n
  return out


def test_header_and_binding_agree():
  fns = _header_functions()
  assert len(fns) >= 20
  assert set(fns) == set(_lib.SIGNATURES), set(fns) ^ set(_lib.SIGNATURES)
  for name, n in fns.items():
    assert len(_lib.SIGNATURES[name][1]) == n, (name, n, len(_lib.SIGNATURES[name][1]))


def test_library_loads_and_exports_every_symbol():
  if not os.path.isfile(_lib.LIB_PATH):
    import __graft_entry__ as g
    g.build()
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for name in _header_functions():
    assert hasattr(lib, name), name
  lib.mmt_version.restype = ctypes.c_int32
  assert lib.mmt_version() >= 100


def test_gemm_desc_matches_header_struct():
  src = open(os.path.join(ROOT, "include", "mmt_b200.h")).read()
  body = re.search(r"typedef struct mmt_gemm_desc \{(.*?)\} mmt_gemm_desc;", src, flags=re.S).group(1)
  body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
  names = []
  for decl in body.split(";"):
    decl = decl.strip()
    if not decl:
      continue
    decl = re.sub(r"^(const\s+)?(float\*|int32_t|int64_t|float)\s*", "", decl)
    names += [n.strip().lstrip("*") for n in decl.split(",")]
  assert names == [f[0] for f in _lib.GemmDesc._fields_]


class _Recorder:
  """Stands in for libmmt_b200.so: checks that every pointer/extent a call receives lies inside a
  tensor the host code allocated."""

  def __init__(self):
    self.ranges = []
    self.calls = []

  def note(self, t):
    if t is not None:
      st = t.untyped_storage()
      self.ranges.append((st.data_ptr(), st.data_ptr() + st.nbytes()))

  def inside(self, p, nbytes):
    return any(lo <= p and p + nbytes <= hi for lo, hi in self.ranges)

  def mmt_gemm(self, dref, stream):
    d = dref._obj
    self.calls.append("mmt_gemm")
    M, N, K = d.M, d.N, d.K
    for z in range(d.batch):
      z0, z1 = z // d.batch_inner, z % d.batch_inner
      a = d.A + 4 * (z0 * d.a_bs0 + z1 * d.a_bs1)
      b = d.B + 4 * (z0 * d.b_bs0 + z1 * d.b_bs1)
      c = d.C + 4 * (z0 * d.c_bs0 + z1 * d.c_bs1)

      def a_off(m, k):
        if d.a_kb > 0:
          return m * d.a_ms + (k // d.a_kb) * d.a_kbs + (k % d.a_kb) * d.a_ks
        return m * d.a_ms + k * d.a_ks

      def c_off(m, n):
        if d.c_mb > 0:
          return (m // d.c_mb) * d.c_mbs + (m % d.c_mb) * d.c_ms + n
        return m * d.c_ms + n

      for (m, k) in ((0, 0), (M - 1, K - 1), (M - 1, 0), (0, K - 1)):
        assert self.inside(a + 4 * a_off(m, k), 4), ("A", M, N, K, m, k)
      for (n, k) in ((0, 0), (N - 1, K - 1)):
        assert self.inside(b + 4 * (n * d.b_ns + k * d.b_ks), 4), ("B", M, N, K)
      for (m, n) in ((0, 0), (M - 1, N - 1)):
        assert self.inside(c + 4 * c_off(m, n), 4), ("C", M, N, K)
        if d.aux:
          assert self.inside(d.aux + 4 * (z0 * d.c_bs0 + z1 * d.c_bs1 + c_off(m, n)), 4)
        if d.add:
          assert self.inside(d.add + 4 * (z0 * d.c_bs0 + z1 * d.c_bs1 + c_off(m, n)), 4)
      if d.bias:
        assert self.inside(d.bias + 4 * (z * d.bias_bs + N - 1), 4)
    return 0

  def mmt_gemm16(self, dref, stream):
    d = dref._obj
    self.calls.append("mmt_gemm16")
    M, N, K = d.M, d.N, d.K
    assert d.a_ld % 8 == 0 and d.b_ld % 8 == 0 and d.A % 16 == 0 and d.B % 16 == 0, "TMA alignment"
    assert d.C32 or d.C16
    for z in range(d.batch):
      z0, z1 = z // d.batch_inner, z % d.batch_inner
      a = d.A + 2 * (z0 * d.a_bs0 + z1 * d.a_bs1)
      b = d.B + 2 * (z0 * d.b_bs0 + z1 * d.b_bs1)
      zc = z0 * d.c_bs0 + z1 * d.c_bs1
      for (m, k) in ((0, 0), (M - 1, K - 1), (M - 1, 0), (0, K - 1)):
        off = k * d.a_ld + m if d.a_mn else m * d.a_ld + k
        assert self.inside(a + 2 * off, 2), ("A", M, N, K, m, k)
      for (n, k) in ((0, 0), (N - 1, K - 1), (N - 1, 0), (0, K - 1)):
        off = k * d.b_ld + n if d.b_mn else n * d.b_ld + k
        assert self.inside(b + 2 * off, 2), ("B", M, N, K, n, k)
      for (m, n) in ((0, 0), (M - 1, N - 1)):
        if d.C32:
          assert self.inside(d.C32 + 4 * (zc + m * d.c32_ld + n), 4), ("C32", M, N, K)
        if d.C16:
          assert self.inside(d.C16 + 2 * (zc + m * d.c16_ld + n), 2), ("C16", M, N, K)
        if d.add:
          assert self.inside(d.add + 4 * (zc + m * d.add_ld + n), 4)
        if d.aux16:
          assert self.inside(d.aux16 + 2 * (zc + m * d.aux_ld + n), 2)
      if d.bias:
        assert self.inside(d.bias + 4 * (z * d.bias_bs + N - 1), 4)
      if d.colsum:
        assert self.inside(d.colsum + 4 * (z1 * d.colsum_bs + N - 1), 4)
    return 0

  def __getattr__(self, name):
    if not name.startswith("mmt_"):
      raise AttributeError(name)

    def f(*args):
      self.calls.append(name)
      for a in args:
        if isinstance(a, int) and a > (1 << 32):      # looks like a pointer
          assert self.inside(a, 4), (name, hex(a))
      return 0
    return f


@pytest.fixture
def recorder(monkeypatch):
  rec = _Recorder()
  real_ptr = _lib.ptr

  def ptr(t, offset=0):
    rec.note(t)
    return real_ptr(t, offset)

  monkeypatch.setattr(_lib, "_lib", rec)
  monkeypatch.setattr(_lib, "_note", lambda *ts: [rec.note(t) for t in ts])
  monkeypatch.setattr(_lib, "ptr", ptr)
  monkeypatch.setattr(engine, "ptr", ptr)
  monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
  monkeypatch.setattr(engine, "stream_ptr", lambda: 0)
  monkeypatch.setattr(engine16, "ptr", ptr)
  monkeypatch.setattr(engine16, "stream_ptr", lambda: 0)
  return rec


def _mini_net(heads=4):
  import types
  from mmt_b200.model.model import CENet
  ed = O.compute_dims(["s3d", "vggish", "ocr"])
  vb = {"hidden_size": 128, "num_hidden_layers": 2, "num_attention_heads": heads,
        "intermediate_size": 256, "hidden_act": "gelu", "hidden_dropout_prob": 0.1,
        "attention_probs_dropout_prob": 0.1, "max_position_embeddings": 32, "type_vocab_size": 19,
        "initializer_range": 0.02, "layer_norm_eps": 1e-12}

  class Stub(torch.nn.Module):
    def __init__(self):
      super().__init__()
      self.config = types.SimpleNamespace(hidden_size=96)

  net = CENet(l2renorm=False, expert_dims=ed, tokenizer=None, keep_missing_modalities=True,
              test_caption_mode="indep", txt_inp="bertftn", txt_agg="bertftn", txt_wgh="emb",
              vid_wgh="none", vid_cont="bert", vid_inp="both", pos_enc="tint", out_tok="mxp",
              same_dim=128, vid_bert_params=vb, txt_pro="gbn",
              txt_bert_params={"hidden_dropout_prob": 0.1, "attention_probs_dropout_prob": 0.1},
              txt_bert=Stub())
  return net, ed


@pytest.mark.parametrize("heads", [1, 2])
def test_engine16_dry_run_addresses_only_owned_memory(recorder, heads):
  """The 16-bit operand sequencing (engine16.py): every pointer / extent handed to the library lies inside a
  tensor the host code allocated, TMA alignment rules hold for every GEMM operand.  heads=1: dh = 128, the
  tcgen05 attention kernels; heads=2: dh = 64, the short-sequence attention of the text encoder."""
  net, ed = _mini_net(heads)
  net.cfg.precision = _lib.PREC_F16
  net._prepare16()
  B, T = 5, 7
  batch = O.synth_batch(ed, B, T, text_dim=96)
  mods = list(ed.keys())
  feats = [batch["features"][m] for m in mods]
  maxp = [batch["features_maxpool"][m] for m in mods]
  ft = torch.stack([batch["features_t"][m] for m in mods], 0)
  ind = torch.stack([batch["features_ind"][m] for m in mods], 0)
  for training in (True, False):
    vid, txt, tw, sv = engine.encode_forward(net.cfg, net.flat, net.buf_flat, batch["text_feat"],
                                             feats, maxp, ft, ind, training, 123)
    assert vid.shape == (B, 3, 128) and txt.shape == (B, 3, 128) and tw.shape == (B, 3)
    g = torch.zeros_like(net.flat)
    dtext = engine.encode_backward(net.cfg, net.flat, g, sv, torch.zeros_like(vid),
                                   torch.zeros_like(txt), torch.zeros_like(tw))
    assert dtext.shape == (B, 96)
  sims, dots = engine.sims_forward(vid, txt, torch.ones(B, 3) / 3, tw, 1, True)
  engine.sims_backward(torch.zeros_like(sims), dots, vid, txt, torch.ones(B, 3) / 3, tw, 1, True,
                       precision=_lib.PREC_F16, scale16=net.cfg.scale16)
  assert recorder.calls.count("mmt_gemm16") > 60
  for name in ("mmt_pack_inputs16", "mmt_embed_ln16_fwd", "mmt_embed_ln16_bwd", "mmt_ln16_fwd", "mmt_ln16_bwd",
               "mmt_attention16_fwd" if heads == 1 else "mmt_txt_attention_fwd",
               "mmt_attention16_bwd" if heads == 1 else "mmt_txt_attention_bwd", "mmt_cast16", "mmt_readout_norm_fwd",
               "mmt_readout_norm_bwd", "mmt_geu_gate_fwd", "mmt_geu_gate_bwd", "mmt_moe_softmax_fwd",
               "mmt_moe_softmax_bwd", "mmt_sims_combine_fwd", "mmt_sims_combine_bwd", "mmt_colsum"):
    assert name in recorder.calls, name


def test_engine_dry_run_addresses_only_owned_memory(recorder):
  net, ed = _mini_net()
  net.cfg.precision = _lib.PREC_TF32
  B, T = 5, 7
  batch = O.synth_batch(ed, B, T, text_dim=96)
  mods = list(ed.keys())
  feats = [batch["features"][m] for m in mods]
  maxp = [batch["features_maxpool"][m] for m in mods]
  ft = torch.stack([batch["features_t"][m] for m in mods], 0)
  ind = torch.stack([batch["features_ind"][m] for m in mods], 0)
  for training in (True, False):
    vid, txt, tw, sv = engine.encode_forward(net.cfg, net.flat, net.buf_flat, batch["text_feat"],
                                             feats, maxp, ft, ind, training, 123)
    assert vid.shape == (B, 3, 128) and txt.shape == (B, 3, 128) and tw.shape == (B, 3)
    g = torch.zeros_like(net.flat)
    dtext = engine.encode_backward(net.cfg, net.flat, g, sv, torch.zeros_like(vid),
                                   torch.zeros_like(txt), torch.zeros_like(tw))
    assert dtext.shape == (B, 96)
  sims, dots = engine.sims_forward(vid, txt, torch.ones(B, 3) / 3, tw, 1, True)
  engine.sims_backward(torch.zeros_like(sims), dots, vid, txt, torch.ones(B, 3) / 3, tw, 1, True)
  engine.max_margin(sims, 0.05, True)
  assert recorder.calls.count("mmt_gemm") > 60
  for name in ("mmt_embed_ln_fwd", "mmt_embed_ln_bwd", "mmt_res_ln_fwd", "mmt_res_ln_bwd",
               "mmt_softmax_mask_fwd", "mmt_softmax_mask_bwd", "mmt_readout_norm_fwd",
               "mmt_readout_norm_bwd", "mmt_geu_gate_fwd", "mmt_geu_gate_bwd", "mmt_dropout",
               "mmt_moe_softmax_fwd", "mmt_moe_softmax_bwd", "mmt_sims_combine_fwd",
               "mmt_sims_combine_bwd", "mmt_max_margin_fwd_bwd", "mmt_colsum"):
    assert name in recorder.calls, name


def test_cenet_keeps_reference_state_dict_names_and_flat_views():
  net, ed = _mini_net()
  P = O.init_params(ed, net.vid_bert_params, text_dim=96, same_dim=128)
  sd = net.state_dict()
  assert set(sd.keys()) == set(P.keys())
  for k in P:
    assert tuple(sd[k].shape) == tuple(P[k].shape), k
  net.load_state_dict(P)
  L = net.layout
  # Q|K|V weights are one contiguous [3d, d] block of the flat buffer
  q = L.off("vid_bert.encoder.layer.0.attention.self.query.weight")
  qkv = net.flat[q:q + 3 * 128 * 128].view(3 * 128, 128)
  assert torch.equal(qkv[128:256], P["vid_bert.encoder.layer.0.attention.self.key.weight"])
  # parameters are views: an in-place optimiser update lands in the flat buffer
  p = net._param("text_GU.ocr.fc.weight")
  with torch.no_grad():
    p.add_(1.0)
  o = L.off("text_GU.ocr.fc.weight")
  assert torch.equal(net.flat[o:o + p.numel()].view_as(p), p)
  # text GEU fc weights of all experts form one [M*d, text_dim] matrix
  o0 = L.off("text_GU.%s.fc.weight" % list(ed.keys())[0])
  assert L.off("text_GU.%s.fc.weight" % list(ed.keys())[1]) == o0 + 128 * 96


def test_unsupported_branches_raise():
  from mmt_b200.model.model import CENet
  ed = O.compute_dims(["s3d"])
  with pytest.raises(NotImplementedError):
    CENet(l2renorm=False, expert_dims=ed, tokenizer=None, keep_missing_modalities=True,
          test_caption_mode="indep", txt_agg="bertftn", vid_cont="coll", vid_inp="both",
          pos_enc="tint", out_tok="mxp", vid_wgh="none", txt_wgh="emb", txt_pro="gbn",
          vid_bert_params={"hidden_size": 512})


def test_product_path_fails_loudly_without_cuda():
  net, ed = _mini_net()
  batch = O.synth_batch(ed, 2, 3, text_dim=96)
  with pytest.raises(RuntimeError, match="CUDA"):
    net(batch["token_ids"], batch["features"], batch["features_t"], batch["features_ind"],
        batch["features_avgpool"], batch["features_maxpool"], batch["query_masks"])


def test_grad_reducer_covers_the_buffer_exactly_once():
  """GradReducer (data-parallel gradient exchange in pieces): early pieces + finish() must cover the flat
  buffer exactly once, and overlapping pieces are a bug that is reported, not summed twice."""
  import torch
  from mmt_b200 import parallel

  class _Work:
    def wait(self):
      pass

  calls = []

  class _Dist:
    class ReduceOp:
      SUM = "sum"

    @staticmethod
    def all_reduce(t, op=None, group=None, async_op=False):
      calls.append((t.storage_offset(), t.numel()))
      t.mul_(2.0)                                   # stand-in for SUM over two identical ranks
      return _Work()

  real = parallel.dist
  parallel.dist = _Dist
  try:
    g = torch.arange(100, dtype=torch.float32)
    red = parallel.GradReducer(g, None)
    red.reduce(60, 20)
    red.reduce(10, 30)
    red.reduce(0, 0)                                # empty pieces are ignored
    red.finish()
    assert sorted(calls) == [(0, 10), (10, 30), (40, 20), (60, 20), (80, 20)]
    assert torch.equal(g, 2 * torch.arange(100, dtype=torch.float32))
    red = parallel.GradReducer(g, None)
    red.reduce(0, 50)
    red.reduce(40, 20)
    with pytest.raises(RuntimeError):
      red.finish()
  finally:
    parallel.dist = real


def test_data_parallel_gradient_pieces_partition_the_real_layout():
  """The pieces the data-parallel backward exchanges (head matrices early, one block per encoder layer, the
  rest at the end; the pooler skipped) cover the real flat layout exactly once."""
  import mmt_test_helpers as H
  from mmt_b200 import parallel
  from mmt_b200.params import Layout
  ed, vb, P, batch, cfg = H.make_case(["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"], 2, 30)
  L = Layout(ed, vb, 768, 512)
  pieces = [(o, n) for o, n in parallel.head_segments(L) if o >= L.small_numel]
  pieces += [L.layer_big_range(l) for l in range(L.L)]
  skipped = L.no_grad_ranges()
  covered = sorted(pieces + skipped)
  pos, rest = 0, []
  for o, n in covered + [(L.numel, 0)]:
    assert o >= pos, "overlapping pieces"
    if o > pos:
      rest.append((pos, o - pos))
    pos = o + n
  assert len(rest) == 1 and rest[0][0] == 0            # ONE remainder piece: small region + ReduceDim weights
  assert rest[0][1] == L.layer_big_range(0)[0]
  # every trainable parameter lies inside exactly one exchanged piece
  for name, seg in L.segments.items():
    if name.startswith("vid_bert.pooler.dense.weight"):
      continue
    inside = [1 for o, n in pieces + rest if o <= seg.offset and seg.offset + seg.numel <= o + n]
    assert len(inside) == 1, name


def test_bench_reference_arm_prints_the_contract_line():
  """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints one JSON line with the
  contract's keys; smallest workload so the test stays in seconds."""
  import json
  import subprocess
  import sys
  r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                      "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-2000:]
  line = json.loads(r.stdout.strip().splitlines()[-1])
  for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
    assert k in line, k
  assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
  assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="needs the reference checkout")
def test_integration_shim_package_overrides_three_modules_and_keeps_the_rest():
  """integration/model ahead of the reference on sys.path: model.model / model.loss / model.metric are
  the B200 surface, any other submodule (here model.bert) is the reference's own file."""
  import subprocess
  import sys
  code = (
      "import model, model.model, model.loss, model.metric, model.bert\n"
      "import mmt_b200.model.model as mm, mmt_b200.model.loss as ml, mmt_b200.model.metric as mt\n"
      "assert model.model.CENet is mm.CENet\n"
      "assert model.model.sharded_cross_view_inner_product is mm.sharded_cross_view_inner_product\n"
      "assert model.loss.MaxMarginRankingLoss is ml.MaxMarginRankingLoss\n"
      "assert model.metric.t2v_metrics is mt.t2v_metrics and model.metric.v2t_metrics is mt.v2t_metrics\n"
      "assert model.bert.__file__.startswith('/root/reference/model/'), model.bert.__file__\n"
      "assert hasattr(model.bert, 'BertModel')\n"
      "print('shim ok')\n")
  env = dict(os.environ, MMT_REFERENCE_ROOT="/root/reference",
             PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "integration"), ROOT]))
  r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300, cwd="/tmp")
  assert r.returncode == 0 and "shim ok" in r.stdout, r.stderr[-2000:]
