"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded
inputs, plus the committed golden fixtures.  Tolerance: 1e-3 relative fp32 (BASELINE.json
north_star); the fp32 CUDA-core mode is held to 2e-4."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import mmt_oracle as O
import mmt_test_helpers as H

pytestmark = pytest.mark.gpu

TOL = 1e-3
TOL_FP32 = 2e-4


@pytest.fixture(scope="module")
def dev():
  assert torch.cuda.is_available(), "gpu tests need a CUDA device"
  from mmt_b200 import _lib
  _lib.load()          # fails loudly if libmmt_b200.so is missing
  return torch.device("cuda")


# ------------------------------------------------------------------------------- GEMM
def _ref_gemm(A, B, bias=None, add=None):
  r = A.double() @ B.double().t()
  if bias is not None:
    r = r + bias.double()
  if add is not None:
    r = r + add.double()
  return r


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 96, 300), (13, 7, 768), (512, 512, 1000),
                                   (1, 1, 1), (300, 1536, 512)])
def test_gemm_fp32_layouts(dev, M, N, K):
  from mmt_b200 import _lib
  g = torch.Generator().manual_seed(M * 7 + N)
  A = torch.randn(M, K, generator=g).to(dev)
  B = torch.randn(N, K, generator=g).to(dev)
  bias = torch.randn(N, generator=g).to(dev)
  add = torch.randn(M, N, generator=g).to(dev)
  ref = _ref_gemm(A, B, bias, add)
  for a_t in (False, True):
    for b_t in (False, True):
      Am = A.t().contiguous() if a_t else A          # a_t: stored [K, M]
      Bm = B.t().contiguous() if b_t else B
      C = torch.empty(M, N, device=dev)
      _lib.gemm(M, N, K, Am, 1 if a_t else K, M if a_t else 1, Bm, 1 if b_t else K,
                N if b_t else 1, C, N, bias=bias, add=add)
      assert H.rel_err(C, ref) < 2e-5, (a_t, b_t)


def test_gemm_epilogues_batch_remap_splitk(dev):
  from mmt_b200 import _lib
  g = torch.Generator().manual_seed(5)
  M, N, K = 150, 260, 96
  A = torch.randn(M, K, generator=g).to(dev)
  B = torch.randn(N, K, generator=g).to(dev)
  bias = torch.randn(N, generator=g).to(dev)
  u_ref = _ref_gemm(A, B, bias)
  f = torch.empty(M, N, device=dev)
  u = torch.empty(M, N, device=dev)
  _lib.gemm(M, N, K, A, K, 1, B, K, 1, f, N, bias=bias, epilogue=_lib.EPI_GELU, aux=u)
  assert H.rel_err(u, u_ref) < 2e-5
  assert H.rel_err(f, O.gelu(u_ref)) < 2e-5
  # dgelu epilogue: C = (A B^T) * gelu'(aux)
  dg = torch.empty(M, N, device=dev)
  _lib.gemm(M, N, K, A, K, 1, B, K, 1, dg, N, epilogue=_lib.EPI_DGELU, aux=u)
  ur = u_ref.clone().requires_grad_(True)
  O.gelu(ur).sum().backward()
  assert H.rel_err(dg, _ref_gemm(A, B) * ur.grad) < 2e-5
  # batched, two-level batch index, strided heads (attention scores layout)
  Bt, Hh, S, dh = 3, 4, 37, 128
  qkv = torch.randn(Bt * S, 3 * Hh * dh, generator=g).to(dev)
  Sp = (S + 3) // 4 * 4
  P = torch.zeros(Bt, Hh, S, Sp, device=dev)
  d = Hh * dh
  _lib.gemm(S, S, dh, qkv, 3 * d, 1, qkv, 3 * d, 1, P, Sp, b_off=d, batch=Bt * Hh, batch_inner=Hh,
            a_bs=(S * 3 * d, dh), b_bs=(S * 3 * d, dh), c_bs=(Hh * S * Sp, S * Sp))
  q = qkv[:, :d].view(Bt, S, Hh, dh).permute(0, 2, 1, 3).double()
  k = qkv[:, d:2 * d].view(Bt, S, Hh, dh).permute(0, 2, 1, 3).double()
  assert H.rel_err(P[..., :S], q @ k.transpose(-1, -2)) < 2e-5
  # output row remap (ReduceDim writes into token slots) and 2-level K index (its wgrad)
  Bb, T, Sx, dd, din = 4, 5, 13, 128, 70
  x = torch.randn(Bb * T, din, generator=g).to(dev)
  W = torch.randn(dd, din, generator=g).to(dev)
  proj = torch.zeros(Bb * Sx, dd, device=dev)
  _lib.gemm(Bb * T, dd, din, x, din, 1, W, din, 1, proj, dd, c_off=2 * dd, c_mb=T, c_mbs=Sx * dd)
  ref = _ref_gemm(x, W).view(Bb, T, dd)
  assert H.rel_err(proj.view(Bb, Sx, dd)[:, 2:2 + T], ref) < 2e-5
  assert float(proj.view(Bb, Sx, dd)[:, :2].abs().max()) == 0.0
  dW = torch.empty(dd, din, device=dev)
  _lib.gemm(dd, din, Bb * T, proj, 1, dd, x, 1, din, dW, din, a_off=2 * dd, a_kb=T, a_kbs=Sx * dd)
  assert H.rel_err(dW, ref.reshape(Bb * T, dd).t() @ x.double()) < 2e-5
  # split-K (weight-gradient shape)
  Kl = 6000
  dY = torch.randn(Kl, 256, generator=g).to(dev)
  X = torch.randn(Kl, 384, generator=g).to(dev)
  dW = torch.empty(256, 384, device=dev)
  _lib.gemm(256, 384, Kl, dY, 1, 256, X, 1, 384, dW, 384, split_k=True)
  assert H.rel_err(dW, dY.double().t() @ X.double()) < 2e-5


@pytest.mark.parametrize("M,N,K,batch", [(64, 64, 512, 7), (130, 200, 256, 3), (256, 256, 128, 2), (8, 9, 132, 1)])
def test_gemm_fp32_warp_block_dot_path(dev, M, N, K, batch):
  """Small outputs with both operands contiguous along k (the similarity dot products, also at the
  data-parallel global batch): one warp per 8 x 8 block, strided batches, ragged edges, bias."""
  from mmt_b200 import _lib
  g = torch.Generator().manual_seed(M + N)
  ld = K + 4
  A = torch.randn(M, batch, ld, generator=g).to(dev)           # rows strided by batch * ld, batch stride ld
  B = torch.randn(N, batch, ld, generator=g).to(dev)
  bias = torch.randn(batch, N, generator=g).to(dev)
  C = torch.full((batch, M, N), float("nan"), device=dev)
  _lib.gemm(M, N, K, A, batch * ld, 1, B, batch * ld, 1, C, N, batch=batch, a_bs=(ld, 0), b_bs=(ld, 0),
            c_bs=(M * N, 0), bias=bias, bias_bs=N, alpha=0.5)
  ref = 0.5 * torch.einsum("mbk,nbk->bmn", A[..., :K].double(), B[..., :K].double()) + bias.double()[:, None, :]
  assert torch.isfinite(C).all()
  assert H.rel_err(C, ref) < 2e-6


def test_gemm_rejects_bad_arguments(dev):
  from mmt_b200 import _lib
  A = torch.zeros(4, 4, device=dev)
  with pytest.raises(RuntimeError, match="aux"):
    _lib.gemm(4, 4, 4, A, 4, 1, A, 4, 1, A, 4, epilogue=_lib.EPI_GELU)
  with pytest.raises(RuntimeError, match="shape"):
    _lib.gemm(4, 4, -1, A, 4, 1, A, 4, 1, A, 4)


# ------------------------------------------------------------------------------- row kernels
def test_res_ln_fwd_bwd_matches_torch(dev):
  from mmt_b200 import _lib
  lib = _lib.load()
  rows, d = 333, 512
  g = torch.Generator().manual_seed(1)
  t = torch.randn(rows, d, generator=g)
  r = torch.randn(rows, d, generator=g)
  gamma = 1 + 0.1 * torch.randn(d, generator=g)
  beta = 0.1 * torch.randn(d, generator=g)
  dy = torch.randn(rows, d, generator=g)
  dy2 = torch.randn(rows, d, generator=g)
  tr, rr, gr, br = (x.double().requires_grad_(True) for x in (t, r, gamma, beta))
  y_ref = torch.nn.functional.layer_norm(tr + rr, (d,), gr, br, 1e-12)
  y_ref.backward((dy + dy2).double())
  td, rd, gd, bd, dyd, dy2d = (x.to(dev) for x in (t, r, gamma, beta, dy, dy2))   # keep references alive
  y = torch.empty(rows, d, device=dev)
  mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
  _lib.check(lib.mmt_res_ln_fwd(_lib.ptr(td), _lib.ptr(rd), _lib.ptr(gd),
                                _lib.ptr(bd), rows, d, 1e-12, 0.0, 0, 0, _lib.ptr(y),
                                _lib.ptr(mean), _lib.ptr(rstd), _lib.stream_ptr()), "res_ln_fwd")
  e_y, e_z = H.rel_err(y, y_ref), H.rel_err(td, t.double() + r.double())
  assert e_y < 1e-5, e_y
  assert e_z < 1e-6, e_z                                         # z written in place
  dz = torch.empty(rows, d, device=dev)
  dgam, dbet, dbias = (torch.zeros(d, device=dev) for _ in range(3))
  _lib.check(lib.mmt_res_ln_bwd(_lib.ptr(dyd), _lib.ptr(dy2d), _lib.ptr(td),
                                _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gd), rows, d,
                                0.0, 0, 0, _lib.ptr(dz), None, _lib.ptr(dgam), _lib.ptr(dbet),
                                _lib.ptr(dbias), _lib.stream_ptr()), "res_ln_bwd")
  errs = [H.rel_err(dz, tr.grad), H.rel_err(dgam, gr.grad), H.rel_err(dbet, br.grad),
          H.rel_err(dbias, tr.grad.sum(0))]
  assert max(errs) < 2e-5, errs


def test_softmax_mask_fwd_bwd(dev):
  from mmt_b200 import _lib
  lib = _lib.load()
  B, Hh, S = 3, 4, 45
  Sp = 48
  g = torch.Generator().manual_seed(2)
  sc = torch.randn(B, Hh, S, Sp, generator=g)
  mask = (torch.rand(B, S, generator=g) > 0.3).float()
  mask[:, 0] = 1
  scale = 1 / math.sqrt(128)
  x = (sc[..., :S].double() * scale + (1 - mask.double())[:, None, None, :] * -10000.0).requires_grad_(True)
  p_ref = torch.softmax(x, -1)
  dP = torch.randn(B, Hh, S, Sp, generator=g)
  p_ref.backward(dP[..., :S].double())
  P = sc.to(dev)
  maskd = mask.to(dev)
  _lib.check(lib.mmt_softmax_mask_fwd(_lib.ptr(P), _lib.ptr(maskd), B, Hh, S, Sp, scale, 0.0,
                                      0, 0, _lib.ptr(P), None, _lib.stream_ptr()), "softmax_fwd")
  assert H.rel_err(P[..., :S], p_ref) < 1e-5
  assert float(P[..., S:].abs().max()) == 0.0
  dPd = dP.to(dev)
  _lib.check(lib.mmt_softmax_mask_bwd(_lib.ptr(dPd), _lib.ptr(P), B, Hh, S, Sp, scale, 0.0, 0, 0,
                                      _lib.stream_ptr()), "softmax_bwd")
  assert H.rel_err(dPd[..., :S], x.grad * scale) < 2e-5     # gradient w.r.t. the raw QK^T


def test_dropout_is_deterministic_unbiased_and_regenerated(dev):
  from mmt_b200 import _lib
  lib = _lib.load()
  rows, n, p = 4096, 512, 0.1
  x = torch.ones(rows, n, device=dev)
  o1, o2, o3 = (torch.empty_like(x) for _ in range(3))
  st = _lib.stream_ptr()
  _lib.check(lib.mmt_dropout(_lib.ptr(x), _lib.ptr(o1), rows, n, p, 11, 3, st), "dropout")
  _lib.check(lib.mmt_dropout(_lib.ptr(x), _lib.ptr(o2), rows, n, p, 11, 3, st), "dropout")
  _lib.check(lib.mmt_dropout(_lib.ptr(x), _lib.ptr(o3), rows, n, p, 12, 3, st), "dropout")
  assert torch.equal(o1, o2)                                  # same (seed, site) -> same mask
  assert not torch.equal(o1, o3)
  keep = float((o1 > 0).float().mean())
  assert abs(keep - (1 - p)) < 3e-3                            # keep-rate
  assert abs(float(o1.mean()) - 1.0) < 5e-3                    # inverted-dropout scaling
  vals = torch.unique(o1)
  assert len(vals) == 2 and abs(float(vals.max()) - 1 / (1 - p)) < 1e-6


# ------------------------------------------------------------------------------- head / loss
def test_sims_and_loss_match_golden_and_ranking_is_exact(dev, golden_dir):
  from mmt_b200.model.model import sharded_cross_view_inner_product
  from mmt_b200.model.loss import MaxMarginRankingLoss
  z = np.load(os.path.join(golden_dir, "sims_loss.npz"))
  mods = ["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"]
  vid = {m: torch.from_numpy(z["vid/" + m]).to(dev) for m in mods}
  txt = {m: torch.from_numpy(z["txt/" + m]).to(dev) for m in mods}
  vw, tw = torch.from_numpy(z["vw"]).to(dev), torch.from_numpy(z["tw"]).to(dev)
  for merge in ("avg", "indep"):
    s = sharded_cross_view_inner_product(vid, txt, vw, tw, mods, merge)
    ref = z["sims_" + merge]
    assert s.shape == ref.shape
    np.testing.assert_allclose(s.cpu().numpy(), ref, rtol=0, atol=2e-6)
    a = np.argsort(-s.cpu().numpy(), axis=1, kind="stable")
    b = np.argsort(-ref, axis=1, kind="stable")
    assert (a == b).all()                  # ranking indices bit-exact at the similarity boundary
  # CPU inputs are accepted (eval path, trainer.py:396-403) and come back on the CPU
  s_cpu = sharded_cross_view_inner_product({m: v.cpu() for m, v in vid.items()},
                                           {m: v.cpu() for m, v in txt.items()}, vw.cpu(),
                                           tw.cpu(), mods, "indep")
  assert s_cpu.device.type == "cpu"
  for margin, fix in ((0.05, True), (0.2, True), (0.05, False)):
    x = torch.from_numpy(z["sims_avg"]).to(dev).requires_grad_(True)
    l = MaxMarginRankingLoss(margin=margin, fix_norm=fix)(x)
    np.testing.assert_allclose(float(l), float(z["loss_m%g_fix%d" % (margin, fix)]), rtol=2e-6)
    l.backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), z["dloss_m%g_fix%d" % (margin, fix)],
                               rtol=1e-5, atol=1e-9)
  with pytest.raises(ValueError):
    sharded_cross_view_inner_product(vid, txt, vw, tw, mods, "bogus")


@pytest.mark.parametrize("n", [1, 2, 63, 64, 1000, 2048])
def test_max_margin_sizes(dev, n):
  from mmt_b200 import engine
  g = torch.Generator().manual_seed(n)
  x = (torch.rand(n, n, generator=g) * 2 - 1)
  xr = x.double().requires_grad_(True)
  ref = O.max_margin_ranking_loss(xr, 0.05, True)
  loss, dx = engine.max_margin(x.to(dev), 0.05, True)
  if n == 1:
    assert torch.isnan(loss)               # empty mean, like the reference (loss.py:63-65)
    return
  ref.backward()
  assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
  assert H.rel_err(dx, xr.grad) < 1e-5
  loss_f, none = engine.max_margin(x.to(dev), 0.05, True, want_grad=False)   # streaming forward-only kernel
  assert none is None and abs(float(loss_f) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))


def test_sims_backward_matches_oracle_autograd(dev):
  from mmt_b200.model.model import SimsFn
  g = torch.Generator().manual_seed(9)
  Nv, caps, M, d = 10, 2, 5, 256
  vid = torch.nn.functional.normalize(torch.randn(Nv, M, d, generator=g), dim=-1)
  txt = torch.nn.functional.normalize(torch.randn(Nv * caps, M, d, generator=g), dim=-1)
  vw = torch.nn.functional.normalize(torch.rand(Nv, M, generator=g), p=1, dim=-1)
  tw = torch.softmax(torch.randn(Nv * caps, M, generator=g), -1)
  w = torch.randn(Nv, Nv, generator=g)
  vr, tr, twr = (x.double().requires_grad_(True) for x in (vid, txt, tw))
  s_ref = O.sharded_cross_view_inner_product(
      {m: vr[:, m] for m in range(M)}, {m: tr[:, m].view(Nv, caps, d) for m in range(M)},
      vw.double(), twr.view(Nv, caps, M), list(range(M)), "avg")
  (s_ref * w.double()).sum().backward()
  vc, tc, twc = (x.to(dev).requires_grad_(True) for x in (vid, txt, tw))
  s = SimsFn.apply(vc, tc, vw.to(dev), twc, caps, True)
  (s * w.to(dev)).sum().backward()
  assert H.rel_err(s, s_ref) < 1e-5
  assert H.rel_err(vc.grad, vr.grad) < 2e-5
  assert H.rel_err(tc.grad, tr.grad) < 2e-5
  assert H.rel_err(twc.grad, twr.grad) < 2e-5


def test_adam_matches_torch(dev):
  from mmt_b200 import _lib
  lib = _lib.load()
  n = 10007
  g = torch.Generator().manual_seed(4)
  p0 = torch.randn(n, generator=g)
  pr = p0.clone().requires_grad_(True)
  opt = torch.optim.Adam([pr], lr=5e-5, weight_decay=0.01)
  p, m, v = p0.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
  for step in range(1, 4):
    gr = torch.randn(n, generator=g)
    pr.grad = gr.clone()
    opt.step()
    grd = gr.to(dev)
    _lib.check(lib.mmt_adam_step(_lib.ptr(p), _lib.ptr(grd), _lib.ptr(m), _lib.ptr(v), n, 5e-5,
                                 0.9, 0.999, 1e-8, 0.01, step, 1.0, _lib.stream_ptr()), "adam")
  assert float((p.cpu() - pr.detach()).abs().max()) < 1e-6


# ------------------------------------------------------------------------------- end to end
def _run_both(modalities, B, T, layers, caps=1, training=True, **kw):
  from mmt_b200.model.loss import MaxMarginRankingLoss
  ed, vb, P, batch, cfg = H.make_case(modalities, B, T, layers=layers, caps=caps, **kw)
  # oracle (CPU, fp32, autograd backward)
  Pr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
        for k, v in P.items()}
  new_stats = {}
  out_ref = O.cenet_forward(Pr, batch, cfg, training=training, out="conf", new_stats=new_stats,
                            text_feat=batch["text_feat"], return_intermediates=True)
  net = H.build_cuda_net(ed, vb, P, batch)
  net.train(training)
  out = net(**H.batch_kwargs(batch, "cuda"), out="conf", device=torch.device("cuda"))
  return ed, P, Pr, batch, cfg, out_ref, new_stats, net, out


@pytest.mark.parametrize("modalities,B,T,layers", [
    (["s3d", "vggish"], 8, 14, 4),                                  # C1 (BASELINE configs[0])
    (["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"], 6, 30, 2),   # C2 geometry, small B
])
def test_train_step_parity_with_oracle(dev, modalities, B, T, layers):
  from mmt_b200.model.loss import MaxMarginRankingLoss
  ed, P, Pr, batch, cfg, out_ref, new_stats, net, out = _run_both(modalities, B, T, layers)
  conf_ref = out_ref["cross_view_conf_matrix"]
  conf = out["cross_view_conf_matrix"]
  assert out["modalities"] == list(ed.keys())
  assert H.rel_err(conf, conf_ref) < TOL_FP32 and H.rel_l2(conf, conf_ref) < TOL_FP32
  loss_ref = O.max_margin_ranking_loss(conf_ref, 0.05, True)
  loss = MaxMarginRankingLoss(margin=0.05, fix_norm=True)(conf)
  assert abs(float(loss) - float(loss_ref)) < TOL_FP32 * abs(float(loss_ref))
  loss_ref.backward()
  loss.backward()
  gmax = max(float(p.grad.abs().max()) for p in Pr.values() if getattr(p, "grad", None) is not None)
  checked = 0
  for name, pr in Pr.items():
    if not (pr.is_floating_point() and pr.requires_grad):
      continue
    p = net._param(name)
    if pr.grad is None:
      assert p.grad is None, name                      # pooler: no gradient on either side
      continue
    assert p.grad is not None, name
    scale = max(float(pr.grad.abs().max()), 1e-3 * gmax)
    err = float((p.grad.cpu().double() - pr.grad.double()).abs().max()) / scale
    assert err < TOL_FP32, (name, err)
    checked += 1
  assert checked >= 20 + 16 * layers
  # BatchNorm running statistics (model.py:745-750 -> torch BatchNorm1d update)
  for name, v in new_stats.items():
    mod, leaf = net._leaf(name)
    assert H.rel_err(mod._buffers[leaf], v) < 1e-5, name
  assert int(net.nbt_flat[0]) == 1


def test_eval_embds_two_captions(dev):
  ed, vb, P, batch, cfg = H.make_case(["s3d", "vggish", "ocr"], 5, 9, layers=2, caps=2)
  ref = O.cenet_forward(P, batch, cfg, training=False, out="embds", text_feat=batch["text_feat"])
  refc = O.cenet_forward(P, batch, cfg, training=False, out="conf", text_feat=batch["text_feat"])
  net = H.build_cuda_net(ed, vb, P, batch).eval()
  with torch.no_grad():
    e = net(**H.batch_kwargs(batch, "cuda"), out="embds")
    c = net(**H.batch_kwargs(batch, "cuda"), out="conf")
  for k in ("vid_embds", "text_embds", "vid_weights", "text_weights"):
    assert tuple(e[k].shape) == tuple(ref[k].shape), k
    assert H.rel_err(e[k], ref[k]) < TOL_FP32, k
  assert tuple(c["cross_view_conf_matrix"].shape) == (10, 5)
  assert H.rel_err(c["cross_view_conf_matrix"], refc["cross_view_conf_matrix"]) < TOL_FP32
  # t2v ranks equal the oracle's (model/metric.py semantics)
  r1 = O.retrieval_ranks(c["cross_view_conf_matrix"].cpu().numpy())
  r2 = O.retrieval_ranks(refc["cross_view_conf_matrix"].numpy())
  assert (r1 == r2).all()


def test_missing_experts_and_ragged_inputs(dev):
  """All-padding experts (k=0 valid frames), position clamp, B not a multiple of anything."""
  ed, vb, P, batch, cfg = H.make_case(["ocr", "speech", "s3d"], 7, 11, layers=1, seed=77)
  batch["features_ind"]["ocr"][:] = 0            # expert entirely missing for every video
  batch["features"]["ocr"][:] = 0
  batch["features_maxpool"]["ocr"][:] = 0
  batch["features_t"]["ocr"][:] = 1
  batch["features_t"]["s3d"][0, 0] = 500.0       # clamp to max_pos - 1 (model.py:516)
  ref = O.cenet_forward(P, batch, cfg, training=True, out="conf", text_feat=batch["text_feat"])
  net = H.build_cuda_net(ed, vb, P, batch).train()
  out = net(**H.batch_kwargs(batch, "cuda"))
  assert H.rel_err(out["cross_view_conf_matrix"], ref["cross_view_conf_matrix"]) < TOL_FP32
  assert torch.isfinite(out["cross_view_conf_matrix"]).all()


def test_dropout_training_is_finite_and_close_in_expectation(dev):
  ed, vb, P, batch, cfg = H.make_case(["s3d", "vggish"], 8, 14, layers=2, dropout=0.1)
  net = H.build_cuda_net(ed, vb, P, batch, dropout=0.1).train()
  from mmt_b200.model.loss import MaxMarginRankingLoss
  crit = MaxMarginRankingLoss(0.05, True)
  outs = []
  for _ in range(3):
    net.zero_grad()
    out = net(**H.batch_kwargs(batch, "cuda"))["cross_view_conf_matrix"]
    crit(out).backward()
    assert torch.isfinite(out).all()
    assert all(torch.isfinite(p.grad).all() for p in net._hot_params() if p.grad is not None)
    outs.append(out.detach())
  assert not torch.equal(outs[0], outs[1])       # a fresh mask every step
  net.eval()
  with torch.no_grad():
    e1 = net(**H.batch_kwargs(batch, "cuda"), out="embds")["vid_embds"]
    e2 = net(**H.batch_kwargs(batch, "cuda"), out="embds")["vid_embds"]
  assert torch.equal(e1, e2)                     # eval: no dropout, deterministic


def test_gradient_accumulation_and_optimizer_view_semantics(dev):
  from mmt_b200.model.loss import MaxMarginRankingLoss
  ed, vb, P, batch, cfg = H.make_case(["s3d", "vggish"], 4, 6, layers=1)
  net = H.build_cuda_net(ed, vb, P, batch).train()
  crit = MaxMarginRankingLoss(0.05, True)
  crit(net(**H.batch_kwargs(batch, "cuda"))["cross_view_conf_matrix"]).backward()
  g1 = {n: net._param(n).grad.clone() for n in net._names if net._param(n).grad is not None}
  gmax = max(float(g.abs().max()) for g in g1.values())
  crit(net(**H.batch_kwargs(batch, "cuda"))["cross_view_conf_matrix"]).backward()   # no zero_grad
  for n, g in g1.items():
    # key.bias gradients are analytically zero (rounding noise): floor the scale
    err = float((net._param(n).grad - 2 * g).abs().max()) / max(float(g.abs().max()), 1e-3 * gmax)
    assert err < 1e-4, (n, err)
  opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=1e-3)
  before = net.flat.clone()
  opt.step()
  assert not torch.equal(before, net.flat)       # the optimiser wrote through the views


# ------------------------------------------------------------------------------- tcgen05 TF32 path
def _tf32_trunc(x):
  return (x.contiguous().view(torch.int32) & -8192).view(torch.float32)      # keep 10 mantissa bits


TF32_COMP = 1.0 + 2.0 * 0.7213475 / 2048.0     # truncation-bias compensation (mmt_b200/csrc/gemm_tc.cu)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 512), (200, 100, 300),
                                   (13952, 512, 512), (1000, 3072, 512)])
def test_gemm_tf32_all_operand_layouts(dev, M, N, K):
  """tcgen05 kind::tf32 tiles, TMA-fed, for K-major and MN-major A / B (forward, dgrad, wgrad
  layouts).  Checked against fp64 on tf32-truncated operands (tight) and on the raw operands
  (tf32 error bound)."""
  from mmt_b200 import _lib
  g = torch.Generator().manual_seed(M + N + K)
  Mp, Np, Kp = (M + 3) // 4 * 4, (N + 3) // 4 * 4, (K + 3) // 4 * 4      # 16-byte aligned strides
  A = torch.randn(M, K, generator=g).to(dev)
  B = torch.randn(N, K, generator=g).to(dev)
  bias = torch.randn(N, generator=g).to(dev)
  ref = _ref_gemm(A, B, bias)
  ref_t = _ref_gemm(_tf32_trunc(A), _tf32_trunc(B)) * TF32_COMP + bias.double()
  for a_mn in (False, True):
    for b_mn in (False, True):
      if a_mn:
        Am = torch.zeros(K, Mp, device=dev); Am[:, :M] = A.t(); a_ms, a_ks = 1, Mp
      else:
        Am = torch.zeros(M, Kp, device=dev); Am[:, :K] = A; a_ms, a_ks = Kp, 1
      if b_mn:
        Bm = torch.zeros(K, Np, device=dev); Bm[:, :N] = B.t(); b_ns, b_ks = 1, Np
      else:
        Bm = torch.zeros(N, Kp, device=dev); Bm[:, :K] = B; b_ns, b_ks = Kp, 1
      C = torch.full((M, Np), float("nan"), device=dev)
      _lib.gemm(M, N, K, Am, a_ms, a_ks, Bm, b_ns, b_ks, C, Np, bias=bias, precision=_lib.PREC_TF32)
      torch.cuda.synchronize()
      e_t, e = H.rel_err(C[:, :N], ref_t), H.rel_err(C[:, :N], ref)
      print("tf32 gemm %dx%dx%d a_mn=%d b_mn=%d: err vs truncated %.2e, vs exact %.2e" % (M, N, K, a_mn, b_mn, e_t, e))
      assert e < 3e-3, (a_mn, b_mn, e)
      assert e_t < 1e-3 or e < 1e-3, (a_mn, b_mn, e_t)


def test_gemm_tf32_epilogues_and_splitk(dev):
  from mmt_b200 import _lib
  g = torch.Generator().manual_seed(21)
  M, N, K = 384, 256, 128
  A = torch.randn(M, K, generator=g).to(dev)
  B = torch.randn(N, K, generator=g).to(dev)
  bias = torch.randn(N, generator=g).to(dev)
  add = torch.randn(M, N, generator=g).to(dev)
  At, Bt = _tf32_trunc(A), _tf32_trunc(B)
  u_ref = _ref_gemm(At, Bt) * TF32_COMP + bias.double() + add.double()
  f, u = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
  _lib.gemm(M, N, K, A, K, 1, B, K, 1, f, N, bias=bias, add=add, epilogue=_lib.EPI_GELU, aux=u,
            precision=_lib.PREC_TF32)
  assert H.rel_err(u, u_ref) < 1e-4 and H.rel_err(f, O.gelu(u_ref)) < 1e-4
  dg = torch.empty(M, N, device=dev)
  _lib.gemm(M, N, K, A, K, 1, B, K, 1, dg, N, epilogue=_lib.EPI_DGELU, aux=u, precision=_lib.PREC_TF32)
  ur = u_ref.clone().requires_grad_(True)
  O.gelu(ur).sum().backward()
  assert H.rel_err(dg, _ref_gemm(At, Bt) * TF32_COMP * ur.grad) < 1e-4
  # row remap (ReduceDim -> token slots)
  Bb, T, Sx, dd, din = 8, 30, 63, 512, 300
  x = torch.randn(Bb * T, din, generator=g).to(dev)
  W = torch.randn(dd, din, generator=g).to(dev)
  proj = torch.zeros(Bb * Sx, dd, device=dev)
  _lib.gemm(Bb * T, dd, din, x, din, 1, W, din, 1, proj, dd, c_off=2 * dd, c_mb=T, c_mbs=Sx * dd,
            precision=_lib.PREC_TF32)
  ref = (_ref_gemm(_tf32_trunc(x), _tf32_trunc(W)) * TF32_COMP).view(Bb, T, dd)
  assert H.rel_err(proj.view(Bb, Sx, dd)[:, 2:2 + T], ref) < 1e-4
  assert float(proj.view(Bb, Sx, dd)[:, :2].abs().max()) == 0.0
  # split-K weight gradient: dW [512, 3072] = dY^T X over 13952 rows (both operands MN-major)
  Kl = 13952
  dY = torch.randn(Kl, 512, generator=g).to(dev)
  X = torch.randn(Kl, 3072, generator=g).to(dev)
  dW = torch.empty(512, 3072, device=dev)
  _lib.gemm(512, 3072, Kl, dY, 1, 512, X, 1, 3072, dW, 3072, precision=_lib.PREC_TF32, split_k=True)
  assert H.rel_err(dW, TF32_COMP * (_tf32_trunc(dY).double().t() @ _tf32_trunc(X).double())) < 1e-4
  dW2 = torch.empty(512, 512, device=dev)
  _lib.gemm(512, 512, Kl, dY, 1, 512, X, 1, 3072, dW2, 512, precision=_lib.PREC_TF32, split_k=True)
  assert H.rel_err(dW2, TF32_COMP * (_tf32_trunc(dY).double().t() @ _tf32_trunc(X[:, :512]).double())) < 1e-4


@pytest.mark.parametrize("Bt", [3, 40])
def test_gemm_tf32_batched_attention_layouts(dev, Bt):
  """The six attention matmuls (scores, context and their four backward products) as batched
  tcgen05 GEMMs over (b, h) with head-strided operands (rank-4 TMA maps).  Bt=3 takes the tiled
  kernel, Bt=40 the persistent one (work items = (b, h, tile))."""
  from mmt_b200 import _lib
  g = torch.Generator().manual_seed(33)
  Hh, S, dh = 4, 218, 128
  d = Hh * dh
  Sp = (S + 3) // 4 * 4
  qkv = torch.randn(Bt * S, 3 * d, generator=g).to(dev)
  q = qkv[:, :d].view(Bt, S, Hh, dh).permute(0, 2, 1, 3)
  k = qkv[:, d:2 * d].view(Bt, S, Hh, dh).permute(0, 2, 1, 3)
  v = qkv[:, 2 * d:].view(Bt, S, Hh, dh).permute(0, 2, 1, 3)
  tq, tk, tv = (_tf32_trunc(x.contiguous()).double() for x in (q, k, v))
  bsP, bsQ = (Hh * S * Sp, S * Sp), (S * 3 * d, dh)
  kw = dict(batch=Bt * Hh, batch_inner=Hh, precision=_lib.PREC_TF32)
  P = torch.zeros(Bt, Hh, S, Sp, device=dev)
  _lib.gemm(S, S, dh, qkv, 3 * d, 1, qkv, 3 * d, 1, P, Sp, b_off=d, a_bs=bsQ, b_bs=bsQ, c_bs=bsP, **kw)
  assert H.rel_err(P[..., :S], TF32_COMP * (tq @ tk.transpose(-1, -2))) < 1e-4
  Pn = torch.softmax(torch.randn(Bt, Hh, S, Sp, generator=g), -1).to(dev)
  Pn[..., S:] = 0
  tP = _tf32_trunc(Pn).double()[..., :S]
  ctx = torch.zeros(Bt * S, d, device=dev)
  _lib.gemm(S, dh, S, Pn, Sp, 1, qkv, 1, 3 * d, ctx, d, b_off=2 * d, a_bs=bsP, b_bs=bsQ, c_bs=(S * d, dh), **kw)
  ref = (TF32_COMP * (tP @ tv)).permute(0, 2, 1, 3).reshape(Bt * S, d)
  assert H.rel_err(ctx, ref) < 1e-4
  dctx = torch.randn(Bt * S, d, generator=g).to(dev)
  tdc = _tf32_trunc(dctx).double().view(Bt, S, Hh, dh).permute(0, 2, 1, 3)
  dP = torch.zeros(Bt, Hh, S, Sp, device=dev)
  _lib.gemm(S, S, dh, dctx, d, 1, qkv, 3 * d, 1, dP, Sp, b_off=2 * d, a_bs=(S * d, dh), b_bs=bsQ, c_bs=bsP, **kw)
  assert H.rel_err(dP[..., :S], TF32_COMP * (tdc @ tv.transpose(-1, -2))) < 1e-4
  dqkv = torch.zeros(Bt * S, 3 * d, device=dev)
  _lib.gemm(S, dh, S, Pn, 1, Sp, dctx, 1, d, dqkv, 3 * d, c_off=2 * d, a_bs=bsP, b_bs=(S * d, dh), c_bs=bsQ, **kw)
  _lib.gemm(S, dh, S, Pn, Sp, 1, qkv, 1, 3 * d, dqkv, 3 * d, b_off=d, a_bs=bsP, b_bs=bsQ, c_bs=bsQ, **kw)
  _lib.gemm(S, dh, S, Pn, 1, Sp, qkv, 1, 3 * d, dqkv, 3 * d, c_off=d, a_bs=bsP, b_bs=bsQ, c_bs=bsQ, **kw)

  def heads(x):
    return x.view(Bt, S, Hh, dh).permute(0, 2, 1, 3)

  assert H.rel_err(heads(dqkv[:, 2 * d:]), TF32_COMP * (tP.transpose(-1, -2) @ tdc)) < 1e-4
  assert H.rel_err(heads(dqkv[:, :d]), TF32_COMP * (tP @ tk)) < 1e-4
  assert H.rel_err(heads(dqkv[:, d:2 * d]), TF32_COMP * (tP.transpose(-1, -2) @ tq)) < 1e-4


@pytest.mark.parametrize("attn", ["fp32", "tf32"])
def test_train_step_parity_tf32_tensor_core_path(dev, attn):
  """The performance configuration: all encoder linear layers (and, for attn='tf32', the
  attention matmuls) on the tcgen05 tf32 path.  BASELINE.json tolerance: 1e-3 relative fp32 on
  the outputs."""
  from mmt_b200 import _lib
  from mmt_b200.model.loss import MaxMarginRankingLoss
  ed, vb, P, batch, cfg = H.make_case(["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"], 8,
                                      30, layers=4)
  Pr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
        for k, v in P.items()}
  ref = O.cenet_forward(Pr, batch, cfg, training=True, out="conf", text_feat=batch["text_feat"])
  loss_ref = O.max_margin_ranking_loss(ref["cross_view_conf_matrix"], 0.05, True)
  loss_ref.backward()
  net = H.build_cuda_net(ed, vb, P, batch).train()
  net.cfg.precision = _lib.PREC_TF32
  net.cfg.attn_precision = _lib.PREC_TF32 if attn == "tf32" else _lib.PREC_FP32
  out = net(**H.batch_kwargs(batch, "cuda"))["cross_view_conf_matrix"]
  loss = MaxMarginRankingLoss(0.05, True)(out)
  loss.backward()
  e_conf, e_l2 = H.rel_err(out, ref["cross_view_conf_matrix"]), H.rel_l2(out, ref["cross_view_conf_matrix"])
  e_loss = abs(float(loss) - float(loss_ref)) / abs(float(loss_ref))
  gmax = max(float(p.grad.abs().max()) for p in Pr.values() if getattr(p, "grad", None) is not None)
  worst = ("", 0.0)
  for name, pr in Pr.items():
    if not (pr.is_floating_point() and pr.requires_grad) or pr.grad is None:
      continue
    scale = max(float(pr.grad.abs().max()), 1e-2 * gmax)
    err = float((net._param(name).grad.cpu().double() - pr.grad.double()).abs().max()) / scale
    if err > worst[1]:
      worst = (name, err)
  print("tf32 path (attention %s): conf max-rel %.2e rel-L2 %.2e, loss rel %.2e, worst grad %s %.2e" %
        (attn, e_conf, e_l2, e_loss, worst[0], worst[1]))
  assert e_conf < TOL and e_l2 < TOL and e_loss < TOL
  assert worst[1] < 1e-2, worst


# ------------------------------------------------------------------------------- fused attention
def _attention_ref(qkv, mask, Bt, Hh, S, dh):
  d = Hh * dh
  q, k, v = (qkv[:, i * d:(i + 1) * d].double().view(Bt, S, Hh, dh).permute(0, 2, 1, 3) for i in range(3))
  sc = q @ k.transpose(-1, -2) / math.sqrt(dh) + (1.0 - mask.double())[:, None, None, :] * -10000.0
  p = torch.softmax(sc, -1)
  return (p @ v).permute(0, 2, 1, 3).reshape(Bt * S, d), torch.logsumexp(sc, -1)


@pytest.mark.parametrize("Bt,S", [(3, 218), (2, 442), (2, 31), (1, 224), (2, 225)])
def test_fused_attention_forward_matches_reference(dev, Bt, S):
  """tcgen05 flash-style attention: one key block (S <= 224) and the online-softmax multi-block
  path (S = 225, 442 -- BASELINE config C3), ragged last query tile, masked keys."""
  from mmt_b200 import _lib
  lib = _lib.load()
  Hh, dh = 4, 128
  d = Hh * dh
  g = torch.Generator().manual_seed(S)
  qkv = torch.randn(Bt * S, 3 * d, generator=g)
  mask = (torch.rand(Bt, S, generator=g) > 0.3).float()
  mask[:, 0] = 1
  ref, lse_ref = _attention_ref(qkv, mask, Bt, Hh, S, dh)
  qd, md = qkv.to(dev), mask.to(dev)
  ctx = torch.full((Bt * S, d), float("nan"), device=dev)
  lse = torch.empty(Bt, Hh, S, device=dev)
  _lib.check(lib.mmt_attention_fwd(_lib.ptr(qd), _lib.ptr(md), Bt, Hh, S, dh, 1 / math.sqrt(dh), 0.0, 0, 0,
                                   _lib.ptr(ctx), _lib.ptr(lse), None, None, 0, _lib.stream_ptr()), "attention_fwd")
  torch.cuda.synchronize()
  e_ctx, e_lse = H.rel_err(ctx, ref), float((lse.cpu().double() - lse_ref).abs().max())
  print("fused attention S=%d: ctx rel err %.2e, lse abs err %.2e" % (S, e_ctx, e_lse))
  assert torch.isfinite(ctx).all()
  assert e_ctx < 2e-3 and e_lse < 5e-3


def test_fused_attention_dropout_matches_unfused_path(dev):
  """Same (seed, site) -> the fused kernel drops exactly the probabilities the materialised path
  drops (the backward pass relies on it when it recomputes P)."""
  from mmt_b200 import _lib, engine
  lib = _lib.load()
  Bt, Hh, S, dh = 2, 4, 218, 128
  d = Hh * dh
  Sp = (S + 3) // 4 * 4
  g = torch.Generator().manual_seed(3)
  qkv = torch.randn(Bt * S, 3 * d, generator=g).to(dev)
  mask = torch.ones(Bt, S, device=dev)
  scale, p, seed, site = 1 / math.sqrt(dh), 0.1, 1234567, 20
  ctx = torch.empty(Bt * S, d, device=dev)
  _lib.check(lib.mmt_attention_fwd(_lib.ptr(qkv), _lib.ptr(mask), Bt, Hh, S, dh, scale, p, seed, site,
                                   _lib.ptr(ctx), None, None, None, 0, _lib.stream_ptr()), "attention_fwd")
  P = torch.empty(Bt, Hh, S, Sp, device=dev)
  _lib.gemm(S, S, dh, qkv, 3 * d, 1, qkv, 3 * d, 1, P, Sp, b_off=d, batch=Bt * Hh, batch_inner=Hh,
            a_bs=(S * 3 * d, dh), b_bs=(S * 3 * d, dh), c_bs=(Hh * S * Sp, S * Sp))
  Pd = torch.empty_like(P)
  _lib.check(lib.mmt_softmax_mask_fwd(_lib.ptr(P), _lib.ptr(mask), Bt, Hh, S, Sp, scale, p, seed, site,
                                      _lib.ptr(P), _lib.ptr(Pd), _lib.stream_ptr()), "softmax")
  ctx2 = torch.empty(Bt * S, d, device=dev)
  _lib.gemm(S, dh, S, Pd, Sp, 1, qkv, 1, 3 * d, ctx2, d, b_off=2 * d, batch=Bt * Hh, batch_inner=Hh,
            a_bs=(Hh * S * Sp, S * Sp), b_bs=(S * 3 * d, dh), c_bs=(S * d, dh))
  assert H.rel_err(ctx, ctx2) < 2e-3
  # training mode of the fused kernel: P and dropout(P) streamed out for the backward pass must be the
  # materialised path's matrices (same dropout decisions, entry for entry), ctx and lse unchanged
  for S2, p2 in ((218, 0.1), (218, 0.0), (31, 0.1), (224, 0.1)):
    Sp2 = (S2 + 3) // 4 * 4
    g2 = torch.Generator().manual_seed(S2)
    qkv2 = torch.randn(Bt * S2, 3 * d, generator=g2).to(dev)
    mask2 = (torch.rand(Bt, S2, generator=g2) > 0.3).float()
    mask2[:, 0] = 1
    mask2 = mask2.to(dev)
    Pm = torch.empty(Bt, Hh, S2, Sp2, device=dev)
    _lib.gemm(S2, S2, dh, qkv2, 3 * d, 1, qkv2, 3 * d, 1, Pm, Sp2, b_off=d, batch=Bt * Hh, batch_inner=Hh,
              a_bs=(S2 * 3 * d, dh), b_bs=(S2 * 3 * d, dh), c_bs=(Hh * S2 * Sp2, S2 * Sp2),
              precision=_lib.PREC_TF32)
    Pdm = torch.empty_like(Pm) if p2 > 0 else Pm
    _lib.check(lib.mmt_softmax_mask_fwd(_lib.ptr(Pm), _lib.ptr(mask2), Bt, Hh, S2, Sp2, scale, p2, seed, site,
                                        _lib.ptr(Pm), _lib.ptr(Pdm) if p2 > 0 else None, _lib.stream_ptr()),
               "softmax")
    ctx_a = torch.empty(Bt * S2, d, device=dev)
    ctx_b = torch.empty(Bt * S2, d, device=dev)
    lse_a, lse_b = torch.empty(Bt, Hh, S2, device=dev), torch.empty(Bt, Hh, S2, device=dev)
    Pf = torch.full((Bt, Hh, S2, Sp2), float("nan"), device=dev)
    Pdf = torch.full((Bt, Hh, S2, Sp2), float("nan"), device=dev) if p2 > 0 else Pf
    _lib.check(lib.mmt_attention_fwd(_lib.ptr(qkv2), _lib.ptr(mask2), Bt, Hh, S2, dh, scale, p2, seed, site,
                                     _lib.ptr(ctx_a), _lib.ptr(lse_a), None, None, 0, _lib.stream_ptr()), "attention_fwd")
    _lib.check(lib.mmt_attention_fwd(_lib.ptr(qkv2), _lib.ptr(mask2), Bt, Hh, S2, dh, scale, p2, seed, site,
                                     _lib.ptr(ctx_b), _lib.ptr(lse_b), _lib.ptr(Pf),
                                     _lib.ptr(Pdf) if p2 > 0 else None, Sp2, _lib.stream_ptr()), "attention_fwd")
    torch.cuda.synchronize()
    assert torch.isfinite(Pf[..., :S2]).all() and torch.isfinite(Pdf[..., :S2]).all()
    eP = float((Pf[..., :S2] - Pm[..., :S2]).abs().max())
    ePd = float((Pdf[..., :S2] - Pdm[..., :S2]).abs().max())
    same_mask = bool(((Pdf[..., :S2] == 0) == (Pdm[..., :S2] == 0)).all())
    print("saved probabilities S=%d p=%.1f: |dP| %.2e |dPd| %.2e ctx %.2e lse %.2e" %
          (S2, p2, eP, ePd, H.rel_err(ctx_b, ctx_a), float((lse_a - lse_b).abs().max())))
    assert eP < 2e-4 and ePd < 3e-4 and same_mask
    # ctx: the P V product sees normalised instead of un-normalised probabilities, i.e. different tf32
    # truncations of its A operand -- same accuracy class, not bit-identical
    assert H.rel_err(ctx_b, ctx_a) < 1e-3 and float((lse_a - lse_b).abs().max()) < 1e-4


def test_graphed_train_step_matches_eager_and_draws_fresh_dropout(dev):
  """The CUDA-graph replay of the whole train step gives the eager step's loss (p = 0) and, with
  dropout, a different mask on every replay (device-side step counter)."""
  from mmt_b200 import _lib
  from mmt_b200.graph import GraphedTrainStep
  from mmt_b200.model.loss import MaxMarginRankingLoss
  from mmt_b200.optim import FusedAdam
  for p_drop in (0.0, 0.1):
    ed, vb, P, batch, cfg = H.make_case(["s3d", "vggish"], 8, 14, layers=2, dropout=p_drop)
    crit = MaxMarginRankingLoss(0.05, True)
    nets, losses = [], []
    for graphed in (False, True):
      net = H.build_cuda_net(ed, vb, P, batch, dropout=p_drop, precision="tf32").train()
      net._step = 0
      opt = FusedAdam(net, lr=1e-4)
      kw = H.batch_kwargs(batch, "cuda")
      ls = []
      if graphed:
        g = GraphedTrainStep(net, crit, opt, kw, net.txt_bert.hidden[:, 0].clone(), lambda t: None, warmup=2)
        for _ in range(3):
          ls.append(float(g.replay()))
        g.close()
      else:
        for _ in range(5):
          opt.zero_grad()
          l = crit(net(**kw)["cross_view_conf_matrix"])
          l.backward()
          opt.step()
          ls.append(float(l))
      losses.append(ls)
    eager, graph = losses
    if p_drop == 0.0:
      # replays 1..3 == eager steps 3..5 (two eager warm-up steps preceded the capture)
      for a, b in zip(eager[2:], graph):
        assert abs(a - b) < 2e-5 * max(1.0, abs(a)), (eager, graph)
    else:
      assert len(set(round(x, 7) for x in graph)) == len(graph), graph      # fresh masks each replay
      assert all(math.isfinite(x) for x in graph)


@pytest.mark.parametrize("name,kw", [
    # BASELINE configs[2]: ActivityNet geometry, 7 experts, T=62 -> S=442 (two key blocks in the fused
    # attention kernel), max_position_embeddings 102, type_vocab 10
    ("C3-activitynet-S442", dict(modalities=["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"],
                                 B=3, T=62, layers=2, max_pos=102, type_vocab=10)),
    # BASELINE configs[4] geometry: LSMDC (face_dim 128, type_vocab 10)
    ("C5-lsmdc-face128", dict(modalities=["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"],
                              B=4, T=30, layers=2, max_pos=32, type_vocab=10, face_dim=128)),
])
def test_other_baseline_geometries_tf32(dev, name, kw):
  from mmt_b200.model.loss import MaxMarginRankingLoss
  ed, vb, P, batch, cfg = H.make_case(**kw)
  ref = O.cenet_forward(P, batch, cfg, training=True, out="conf", text_feat=batch["text_feat"])
  net = H.build_cuda_net(ed, vb, P, batch, precision="tf32").train()
  out = net(**H.batch_kwargs(batch, "cuda"))["cross_view_conf_matrix"]
  MaxMarginRankingLoss(0.05, True)(out).backward()
  torch.cuda.synchronize()
  e = H.rel_err(out, ref["cross_view_conf_matrix"])
  print("%s: conf max-rel %.2e" % (name, e))
  assert e < 2e-3          # tiny batches: the max-norm figure of a 3x3 / 4x4 matrix is noisy
  assert all(torch.isfinite(p.grad).all() for p in net._hot_params() if p.grad is not None)


def test_eval_scale_similarity_indep(dev):
  """Evaluation-scale call of sharded_cross_view_inner_product (trainer/trainer.py:396-403):
  300 videos x 20 captions, CPU tensors in, 'indep' merge; t2v ranks equal the oracle's."""
  from mmt_b200.model.model import sharded_cross_view_inner_product
  g = torch.Generator().manual_seed(8)
  mods = ["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"]
  nv, caps, d = 300, 20, 512
  vid = {m: torch.nn.functional.normalize(torch.randn(nv, d, generator=g), dim=-1) for m in mods}
  txt = {m: torch.nn.functional.normalize(torch.randn(nv, caps, d, generator=g), dim=-1) for m in mods}
  vw = torch.full((nv, len(mods)), 1.0 / len(mods))
  tw = torch.softmax(torch.randn(nv, caps, len(mods), generator=g), -1)
  ref = O.sharded_cross_view_inner_product(vid, txt, vw, tw, mods, "indep")
  got = sharded_cross_view_inner_product(vid, {m: v.clone() for m, v in txt.items()}, vw, tw, mods, "indep")
  assert got.device.type == "cpu" and tuple(got.shape) == (nv * caps, nv)
  np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=0, atol=2e-6)
  r1, r2 = O.retrieval_ranks(got.numpy()), O.retrieval_ranks(ref.numpy())
  assert float(np.mean(r1 == r2)) > 0.999           # identical up to exact fp32 ties


def test_single_sample_training_batch_raises_like_batchnorm(dev):
  ed, vb, P, batch, cfg = H.make_case(["s3d", "vggish"], 1, 6, layers=1)
  net = H.build_cuda_net(ed, vb, P, batch).train()
  with pytest.raises(ValueError, match="more than 1 value per channel"):
    net(**H.batch_kwargs(batch, "cuda"))
  net.eval()
  with torch.no_grad():
    out = net(**H.batch_kwargs(batch, "cuda"))["cross_view_conf_matrix"]     # eval: running stats
  assert tuple(out.shape) == (1, 1) and torch.isfinite(out).all()


def test_retrieval_metrics_on_gpu_match_reference_golden_and_oracle(dev, golden_dir):
  """Eval path: ranks counted on the device reproduce every scalar the reference's
  t2v_metrics / v2t_metrics produced for the fixture (ties, masked captions) and, at eval scale with
  heavy ties, the oracle's restatement rank for rank (integer / half-integer values: exact)."""
  import numpy as np
  from mmt_b200.model import metric
  fx = np.load(os.path.join(golden_dir, "metrics.npz"))
  sims, qm = fx["sims"], fx["query_masks"]
  t2v, v2t = metric.t2v_metrics(sims, qm), metric.v2t_metrics(torch.from_numpy(sims).to(dev), qm)
  for k in ("R1", "R5", "R10", "R50", "MedR", "MeanR", "geometric_mean_R1-R5-R10"):
    np.testing.assert_allclose(t2v[k], fx["t2v/" + k], rtol=1e-9, err_msg="t2v/" + k)
    np.testing.assert_allclose(v2t[k], fx["v2t/" + k], rtol=1e-9, err_msg="v2t/" + k)
  rng = np.random.RandomState(11)
  for nv, caps, decimals in ((1000, 20, 1), (333, 1, 2), (64, 3, 0)):
    big = np.round(rng.randn(nv * caps, nv), decimals).astype(np.float32)
    m = (rng.rand(nv, caps) > 0.1).astype(np.int32)
    m[:, 0] = 1
    if caps > 1:
      m[5, :] = 0                                     # a video without any caption: v2t rank = inf
    r_t2v = metric.retrieval_ranks(big, None, v2t=False)
    r_v2t = metric.retrieval_ranks(big, m, v2t=True)
    assert np.array_equal(r_t2v, O.retrieval_ranks(big))
    assert np.array_equal(r_v2t, O.retrieval_ranks_v2t(big, m))


@pytest.mark.parametrize("M,N,K", [(512, 384, 320), (1000, 520, 200), (2048, 3072, 512)])
def test_gemm_bf16_operand_mode_experimental(dev, M, N, K):
  """MMT_PREC_BF16 (experimental): bf16 K-major operands made by mmt_cast_bf16, kind::f16 MMAs on the
  CTA-pair kernel, fp32 accumulate and epilogue.  Products of bf16 values are exact in fp32, so against
  an fp64 product of the ROUNDED operands only the accumulation order differs."""
  from mmt_b200 import _lib
  lib = _lib.load()
  g = torch.Generator().manual_seed(K)
  A = torch.randn(M, K, generator=g).to(dev)
  W = (torch.randn(N, K, generator=g) * 0.05).to(dev)
  bias = torch.randn(N, generator=g).to(dev)
  Ab = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
  Wb = torch.empty(N, K, dtype=torch.bfloat16, device=dev)
  _lib.check(lib.mmt_cast_bf16(_lib.ptr(A), _lib.ptr(Ab), M * K, _lib.stream_ptr()), "cast")
  _lib.check(lib.mmt_cast_bf16(_lib.ptr(W), _lib.ptr(Wb), N * K, _lib.stream_ptr()), "cast")
  torch.cuda.synchronize()
  assert torch.equal(Ab, A.to(torch.bfloat16)) and torch.equal(Wb, W.to(torch.bfloat16))   # round to nearest even
  C = torch.full((M, N), float("nan"), device=dev)
  u = torch.empty(M, N, device=dev)
  _lib.gemm(M, N, K, Ab, K, 1, Wb, K, 1, C, N, bias=bias, precision=_lib.PREC_BF16)
  ref = Ab.double() @ Wb.double().t() + bias.double()
  assert torch.isfinite(C).all()
  assert H.rel_err(C, ref) < 2e-6
  _lib.gemm(M, N, K, Ab, K, 1, Wb, K, 1, C, N, bias=bias, epilogue=_lib.EPI_GELU, aux=u, precision=_lib.PREC_BF16)
  assert H.rel_err(u, ref) < 2e-6 and H.rel_err(C, O.gelu(ref)) < 2e-5
  # operands contiguous along m / n (what dgrad and wgrad products need): all four combinations
  if M % 8 == 0 and N % 8 == 0:
    At, Wt = Ab.t().contiguous(), Wb.t().contiguous()          # [K, M], [K, N]
    for a_t, w_t in ((False, True), (True, False), (True, True)):
      C.fill_(float("nan"))
      _lib.gemm(M, N, K, At if a_t else Ab, 1 if a_t else K, M if a_t else 1, Wt if w_t else Wb,
                1 if w_t else K, N if w_t else 1, C, N, bias=bias, precision=_lib.PREC_BF16)
      assert H.rel_err(C, ref) < 2e-6, (a_t, w_t)
