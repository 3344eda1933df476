#!/usr/bin/env python
"""bench.py -- video-caption pairs/sec of one MMT train step (hot path) on N B200s.

A "step" is one pass of the hot path over one synthetic batch, exactly the reference's timed
region (trainer/trainer.py:175-204): zero_grad -> CENet.forward(out='conf') ->
MaxMarginRankingLoss -> backward -> Adam.step, with the reference's dropout probabilities.
Scope "hot path only" (scope A): the third-party text encoder is replaced on
BOTH arms by fixed [B, 768] CLS features.

  python bench.py --gpus N --steps K --warmup W            (this repo's sm_100a path)
  python bench.py --impl reference ...                       (CPU port of the reference, oracle/)
Under torchrun every rank runs; rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

METRIC = "video-caption pairs/sec (train step)"
UNIT = "pairs/s"

WORKLOADS = {
    # BASELINE.json configs[1]: MSRVTT jsfusion, 7 experts, d=512, S=218 (pad 224), batch 64 / GPU
    "C2": dict(name="MSRVTT-jsfusion C2: 7 experts, T=30, S=218, d=512, L=4, H=4, ff=3072",
               modalities=["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"], B=64, T=30,
               max_pos=32, type_vocab=19, face_dim=512),
    # configs[0]: 2 experts, S=31 (pad 32), batch 8 -- the reference's own CPU-runnable case
    "C1": dict(name="MSRVTT-jsfusion C1: 2 experts, T=14, S=31, batch 8",
               modalities=["s3d", "vggish"], B=8, T=14, max_pos=32, type_vocab=19, face_dim=512),
    # configs[2]: ActivityNet geometry with 7 experts, S=442 (pad 448), batch 32
    "C3": dict(name="ActivityNet C3: 7 experts, T=62, S=442, batch 32",
               modalities=["face", "ocr", "rgb", "s3d", "scene", "speech", "vggish"], B=32, T=62,
               max_pos=102, type_vocab=10, face_dim=512),
}
DROPOUT = 0.1          # hidden / attention / moe-text dropout of every published config
LR, WD = 5e-5, 0.0     # configs_pub/eccv20/*.json optimizer


def hotpath_flops(w, B):
  """Algorithmic FLOPs of one train step (fwd + bwd = 3x fwd GEMM FLOPs), SURVEY.md §8(d)."""
  M, T, d, ff, L = len(w["modalities"]), w["T"], 512, 3072, 4
  S = 1 + M * (T + 1)
  import workloads as W
  ed = W.compute_dims(w["modalities"], w["face_dim"])
  sin = sum(v["dim"] for v in ed.values())
  lin = L * 2 * B * S * (4 * d * d + 2 * d * ff)
  attn = L * 4 * B * S * S * d
  k1 = 2 * B * (T + 1) * sin * d          # ReduceDim: forward + weight gradient only (its inputs carry no gradient)
  k11 = M * 2 * B * (768 * d + d * d)
  return 3.0 * (lin + attn + k11) + 2.0 * k1


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
       "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.idx, self.rows, self.proc = gpu_index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits", "-lms", "100"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      f = [x.strip() for x in r.split(",")]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2]))
      except ValueError:
        continue
      for n, v in zip(names, f[5:9]):
        if v.lower().startswith("active"):
          reasons.add(n)
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ workloads
def make_batches(w, B, n_batches, seed0, pin=False):
  import workloads as W
  ed = W.compute_dims(w["modalities"], w["face_dim"])
  out = []
  for i in range(n_batches):
    b = W.synth_batch(ed, B, w["T"], seed=seed0 + i, dense=False)
    if pin:
      for k in ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool"):
        for m in b[k]:
          b[k][m] = b[k][m].pin_memory()
      b["text_feat"] = b["text_feat"].pin_memory()
      b["token_ids"] = b["token_ids"].pin_memory()
    out.append(b)
  return ed, out


VB_FULL = {"vocab_size_or_config_json_file": 10, "hidden_size": 512, "num_hidden_layers": 4,
           "num_attention_heads": 4, "intermediate_size": 3072, "hidden_act": "gelu",
           "initializer_range": 0.02, "layer_norm_eps": 1e-12}     # configs_pub/eccv20/*.json vid_bert_params


def vb_params(w, dropout=DROPOUT):
  return dict(VB_FULL, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout,
              max_position_embeddings=w["max_pos"], type_vocab_size=w["type_vocab"])


def bench_config(w, world, precision):
  """The benchmark definition -- IDENTICAL for both arms (the driver compares the two lines' `config`)."""
  B = w["B"]
  return {"workload": w["name"], "batch_per_gpu": B, "global_batch": B * world,
          "scope": "hot path only: text encoder replaced by fixed [B,768] CLS features on both arms",
          "step": "zero_grad+forward+MaxMarginRankingLoss+backward+Adam, dropout 0.1",
          "parallelism": "dp%d: batch sharded over ranks, one all-gather of embeddings + gradient all-reduce" % world,
          "l2": "GPU arm: ring of 4 distinct input batches (53.7 MB each at C2) and a >1 GB activation working set per "
                "step, far above the 126 MB L2; no explicit flush",
          "precision": "GPU arm: %s GEMM / attention operands, fp32 accumulation, statistics, residuals, master "
                       "weights, gradients and loss; CPU arm: fp32" % precision}


def batch_bytes(b):
  n = 0
  for k in ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool"):
    n += sum(v.numel() * v.element_size() for v in b[k].values())
  n += b["text_feat"].numel() * 4 + b["token_ids"].numel() * 4 + b["query_masks"].numel() * 4
  return n


# ------------------------------------------------------------------------------------ our arm
class TextFeed(torch.nn.Module):
  """Stands for txt_bert on both arms (hot-path-only scope): returns the step's [R, W, 768]
  hidden states whose CLS row is the synthetic text feature."""

  def __init__(self):
    super().__init__()
    import types
    self.config = types.SimpleNamespace(hidden_size=768)
    self.cls = None

  def forward(self, input_ids, **kw):
    return (self.cls.unsqueeze(1),)


def run_b200(args):
  import torch.distributed as dist
  from mmt_b200 import _lib
  from mmt_b200.model.loss import MaxMarginRankingLoss
  from mmt_b200.model.model import CENet
  from mmt_b200.optim import FusedAdam
  import workloads as W

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")     # lets NCCL collectives be captured in a CUDA graph
    dist.init_process_group("nccl", device_id=dev)
  _lib.load()
  torch.manual_seed(0)

  w = WORKLOADS[args.workload]
  B = w["B"]
  vb = vb_params(w)
  NB = 4                                    # ring of distinct input batches: 4 x 53.5 MB > L2
  ed, batches = make_batches(w, B, NB, 1234 + 100 * rank, pin=True)
  P = W.init_params(ed, vb, seed=0)
  PREC = {"fp32": _lib.PREC_FP32, "tf32": _lib.PREC_TF32, "f16": _lib.PREC_F16, "bf16": _lib.PREC_BF16}[args.precision]

  def build_net(dropout, data_parallel):
    f = TextFeed()
    n = CENet(l2renorm=False, expert_dims=ed, tokenizer=None, keep_missing_modalities=True,
              test_caption_mode="indep", txt_inp="bertftn", txt_agg="bertftn", txt_wgh="emb",
              vid_wgh="none", vid_cont="bert", vid_inp="both", pos_enc="tint", out_tok="mxp",
              vid_bert_params=vb_params(w, dropout), txt_pro="gbn",
              txt_bert_params={"hidden_dropout_prob": dropout, "attention_probs_dropout_prob": dropout}, txt_bert=f)
    n.load_state_dict(P, strict=True)
    n.to(dev).train()
    n.cfg.precision = PREC
    if data_parallel:
      n.enable_data_parallel()
    return n, f

  net, feed = build_net(DROPOUT, world > 1)
  crit = MaxMarginRankingLoss(margin=0.05, fix_norm=True)
  opt = FusedAdam(net, lr=LR, weight_decay=WD)

  def to_dev(b):
    kw = {}
    for k in ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool"):
      kw[k] = {m: v.to(dev, non_blocking=True) for m, v in b[k].items()}
    kw["token_ids"] = b["token_ids"].to(dev, non_blocking=True)
    kw["query_masks"] = b["query_masks"]
    return kw, b["text_feat"].to(dev, non_blocking=True)

  resident = [to_dev(b) for b in batches]
  torch.cuda.synchronize()

  def step(kw, text):
    feed.cls = text
    opt.zero_grad()
    out = net(**kw, out="conf", device=dev)
    loss = crit(out["cross_view_conf_matrix"])
    loss.backward()
    opt.step()           # `opt` is looked up at call time (the stock-Adam e2e arm swaps it)
    return loss

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
      fn(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)

  # ---- device-resident throughput ("value") ----
  # With --graph the step is replayed as ONE CUDA graph (mmt_b200/graph.py): static input tensors
  # refreshed from the ring of resident batches with device-to-device copies inside the timed
  # region, dropout seeds / Adam step advancing through a device counter.
  for i in range(args.warmup):
    step(*resident[i % NB])
  n0 = _lib.launch_count()
  step(*resident[0])
  torch.cuda.synchronize()
  launches_per_step = _lib.launch_count() - n0
  graphed = None
  if args.graph:
    from mmt_b200.graph import GraphedTrainStep
    skw = {k: ({m: t.clone() for m, t in v.items()} if isinstance(v, dict) else v)
           for k, v in resident[0][0].items()}
    stext = resident[0][1].clone()
    graphed = GraphedTrainStep(net, crit, opt, skw, stext, lambda t: setattr(feed, "cls", t))

  def value_step(i):
    if graphed is None:
      return step(*resident[i % NB])
    graphed.load(*resident[i % NB])
    return graphed.replay()

  for i in range(3):
    value_step(i)
  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()
  ms = timed(value_step, args.steps)
  launches = launches_per_step * args.steps
  clk = clocks.stop() if rank == 0 else None
  value = B * world * args.steps / (ms / 1e3)
  if graphed is not None:
    graphed.close()

  if args.trace and rank != 0:
    for i in range(3):                      # every rank takes part in the traced steps' collectives
      value_step(i)
    torch.cuda.synchronize()
  if args.trace and rank == 0:
    # development aid: kernel timeline of 3 steps on rank 0 (torch.profiler / CUPTI): busy vs idle time of
    # the device and the share of NCCL kernels, printed to stderr (never part of a reported number)
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
      for i in range(3):
        value_step(i)
      torch.cuda.synchronize()
    ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type.name == "CUDA"]
    ev.sort()
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    busy, cur_s, cur_e = 0.0, ev[0][0], ev[0][1]
    for s_, e_, _ in ev[1:]:
      if s_ > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s_, e_
      else:
        cur_e = max(cur_e, e_)
    busy += cur_e - cur_s
    nccl = sum(e_ - s_ for s_, e_, n in ev if "nccl" in n.lower())
    import collections
    agg = collections.defaultdict(float)
    for s_, e_, n in ev:
      agg[n[:70]] += e_ - s_
    sys.stderr.write("[trace] 3 steps: span %.2f ms, device busy %.2f ms, idle %.2f ms, nccl kernels %.2f ms\n" %
                     ((t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, nccl / 1e3))
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
      sys.stderr.write("[trace]   %8.1f us  %s\n" % (v, n))
  if args.trace and world > 1:
    dist.barrier()

  # ---- end to end through the public API with HOST buffers ("e2e") ----
  # Every timed step performs one pinned-host -> device copy of a full input batch and one
  # device -> host read of the loss.  The copy of step i+1's batch runs on a side stream while
  # step i computes (two static device slots; the per-step loss.item() keeps the slots safe),
  # which is what a DataLoader with pin_memory + non_blocking copies gives the reference trainer.
  copy_stream = torch.cuda.Stream()
  slots = [to_dev(batches[0]), to_dev(batches[1])]
  torch.cuda.synchronize()
  ready = [torch.cuda.Event(), torch.cuda.Event()]

  def stage(slot, b):
    with torch.cuda.stream(copy_stream):
      kw, text = slots[slot]
      for k in ("features", "features_t", "features_ind", "features_avgpool", "features_maxpool"):
        for m in kw[k]:
          kw[k][m].copy_(b[k][m], non_blocking=True)
      kw["token_ids"].copy_(b["token_ids"], non_blocking=True)
      text.copy_(b["text_feat"], non_blocking=True)
      ready[slot].record(copy_stream)

  last = {}
  stage(0, batches[0])

  def e2e_step(i):
    slot = i & 1
    torch.cuda.current_stream().wait_event(ready[slot])
    stage(slot ^ 1, batches[(i + 1) % NB])             # H2D of the next batch overlaps this step
    kw, text = slots[slot]
    last["loss"] = step(kw, text).item()               # device -> host read of the result

  for i in range(2):
    e2e_step(i)
  ms_e2e = timed(lambda i: e2e_step(i + 2), args.steps)
  e2e_eager = B * world * args.steps / (ms_e2e / 1e3)
  e2e_value, e2e_api = e2e_eager, "CENet.forward + MaxMarginRankingLoss + backward + FusedAdam.step, eager launches"
  # the optimizer the UNCHANGED reference train.py:95-100 constructs: torch.optim.Adam over the module's parameters
  # (views of the flat buffer).  Same host-buffer protocol, a few steps.
  e2e_torch_adam = None
  if world == 1 and not args.no_torch_adam:
    opt_fused = opt
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=LR, weight_decay=WD)
    for i in range(2):
      e2e_step(i)
    n_ta = max(4, args.steps // 2)
    ms_ta = timed(lambda i: e2e_step(i + 2), n_ta)
    e2e_torch_adam = B * n_ta / (ms_ta / 1e3)
    opt = opt_fused
  if not args.no_graph_e2e and (world == 1 or args.graph_dp):
    # Same host-buffer protocol through mmt_b200.graph.GraphedTrainStep (the repo's train-step API:
    # the same module / loss / optimizer objects captured once, replayed as one CUDA graph).  With a
    # loss read-back every step the host cannot run ahead, so the ~150 eager launches of a step are
    # exposed; one graph launch is not.  The staged batch is copied device-to-device into the
    # graph's static inputs (inside the timed region).
    from mmt_b200.graph import GraphedTrainStep
    # TWO captures, one per staging slot: the slot's device tensors ARE the graph's static inputs, so the pinned-host
    # batch is copied H2D straight into the inputs of the graph that runs next -- no device-to-device hop.  The
    # two graphs share the net, the optimizer and one device-side step counter.
    def set_text(t):
      feed.cls = t
    g_a = GraphedTrainStep(net, crit, opt, slots[0][0], slots[0][1], set_text)
    g_b = GraphedTrainStep(net, crit, opt, slots[1][0], slots[1][1], set_text, share_with=g_a)
    graphs = [g_a, g_b]
    stage(0, batches[0])

    def e2e_graph_step(i):
      slot = i & 1
      torch.cuda.current_stream().wait_event(ready[slot])
      stage(slot ^ 1, batches[(i + 1) % NB])           # H2D of the next batch into the OTHER graph's inputs
      last["loss"] = graphs[slot].replay().item()      # device -> host read of the result

    for i in range(2):
      e2e_graph_step(i)
    ms_g = timed(lambda i: e2e_graph_step(i + 2), args.steps)
    g_b.close()
    g_a.close()
    if ms_g < ms_e2e:
      ms_e2e = ms_g
      e2e_value = B * world * args.steps / (ms_g / 1e3)
      e2e_api = ("mmt_b200.graph.GraphedTrainStep x2 (one capture per input slot, shared step counter): H2D into the "
                 "idle graph's static inputs + replay")

  res = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
      "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "f16": "f16", "bf16": "bf16"}[args.precision],
      "data": "synthetic",
      "config": bench_config(w, world, args.precision),
      "launch": ("value: whole step replayed as one CUDA graph (%d kernels per step); "
                 "e2e: see e2e.api" % launches_per_step) if graphed is not None else "value: eager launches",
      "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
              "h2d_bytes_per_step": batch_bytes(batches[0]) * world, "d2h_bytes_per_step": 4 * world,
              "api": e2e_api, "eager_value": e2e_eager,
              "eager_torch_adam_value": e2e_torch_adam,
              "note": "value = best public train-step API; eager_value = the CENet.forward path the unchanged reference "
                      "trainer runs with mmt_b200.optim.FusedAdam; eager_torch_adam_value = the same with the stock "
                      "torch.optim.Adam that train.py:95-100 builds"},
      "gpu_launches": launches, "clocks": clk,
      "algorithmic_tflops_per_step": hotpath_flops(w, B) / 1e12,
      "achieved_tflops": hotpath_flops(w, B) * world / (ms / args.steps / 1e3) / 1e12,
  }

  # ---- scope B: the FULL CENet train step, text encoder included (bert-base-cased geometry, random init: the
  # pretrained weights cannot be fetched here), on this repo's kernels (mmt_b200/model/txt_bert.py) ----
  if world == 1 and not args.no_scope_b and _lib.is16(PREC):
    res["scope_b"] = scope_b(args, w, B, dev, P, ed, resident, crit, PREC)

  # ---- parity of THIS configuration (dropout off, fresh copies of the initial weights), outside every timed region
  if not args.no_parity_check:
    res["parity_check"] = parity_check(args, w, B, world, rank, dev, build_net, batches, P, crit)
  prof = profile_facts()
  if prof:
    res["attn_tensor_pipe_pct"] = prof.get("attn_tensor_pipe_pct")
  if rank == 0:
    res["roofline"] = roofline_ffn(net, w, B, dev, args)
    if not args.no_hbm_probe:
      res["roofline_hbm_maxmargin"] = roofline_maxmargin(dev)
    if world == 1 and not args.no_cpu_baseline:
      res["cpu_baseline"] = cpu_baseline(w, steps=2, warmup=1, budget_s=40.0)
  barrier()
  if rank == 0:
    print(json.dumps(res))
  if world > 1:
    dist.destroy_process_group()


def scope_b(args, w, B, dev, P, ed, resident, crit, PREC):
  """Full `CENet` step (SURVEY.md §8(d) scope B): token ids -> text encoder (12 x BERT-base layers, W = 30) -> CLS ->
  text head, plus the whole video side, forward + backward + fused Adam over both flat parameter buffers."""
  from mmt_b200.model.model import CENet
  from mmt_b200.model.txt_bert import TxtBert
  from mmt_b200.optim import FusedAdam
  tb = TxtBert(hidden_dropout_prob=DROPOUT, attention_probs_dropout_prob=DROPOUT, precision=PREC)
  net = CENet(l2renorm=False, expert_dims=ed, tokenizer=None, keep_missing_modalities=True,
              test_caption_mode="indep", txt_inp="bertftn", txt_agg="bertftn", txt_wgh="emb",
              vid_wgh="none", vid_cont="bert", vid_inp="both", pos_enc="tint", out_tok="mxp",
              vid_bert_params=vb_params(w), txt_pro="gbn",
              txt_bert_params={"hidden_dropout_prob": DROPOUT, "attention_probs_dropout_prob": DROPOUT}, txt_bert=tb)
  net.load_state_dict(P, strict=False)
  net.to(dev).train()
  net.cfg.precision = PREC
  opt = FusedAdam(net, lr=LR, weight_decay=WD)

  def step(kw):
    opt.zero_grad()
    loss = crit(net(**kw, out="conf", device=dev)["cross_view_conf_matrix"])
    loss.backward()
    opt.step()
    return loss

  NB = len(resident)
  for i in range(4):
    step(resident[i % NB][0])
  torch.cuda.synchronize()
  n = max(4, args.steps // 2)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(n):
    step(resident[i % NB][0])
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / n
  Wt = resident[0][0]["token_ids"].shape[2]
  txt_flops = 3.0 * 12 * 2 * B * Wt * (4 * 768 * 768 + 2 * 768 * 3072)
  out = {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": n,
         "text_encoder": "mmt_b200.model.txt_bert.TxtBert, BERT-base geometry (12L/768/12H, vocab 28996), W=%d, random init" % Wt,
         "trainable_parameters": int(sum(p.numel() for p in net.parameters() if p.requires_grad)),
         "algorithmic_tflops_per_step": (hotpath_flops(w, B) + txt_flops) / 1e12}
  out["achieved_tflops"] = out["algorithmic_tflops_per_step"] / (ms / 1e3)
  del net, opt, tb
  torch.cuda.empty_cache()
  return out


def profile_facts():
  """Figures that only a profiler can give (tensor-pipe %, DRAM bytes), read from the committed ncu extract of THIS
  round's build -- never measured inside bench.py (a number taken under a profiler is not a bench value)."""
  p = os.path.join(ROOT, "profiles", "r02_kernels.json")
  if os.path.isfile(p):
    try:
      return json.load(open(p))
    except Exception:
      return None
  return None


def parity_check(args, w, B, world, rank, dev, build_net, batches, P, crit):
  """N = 1: one dropout-free train step of the benchmark configuration against the oracle port (conf, loss, every
  gradient).  N > 1: the data-parallel step (rank-local batches) against a single-device step of the GLOBAL batch on
  rank 0 -- so the scaling runs carry the data-parallel path's correctness."""
  import torch.distributed as dist
  import mmt_test_helpers as H
  import workloads as W

  def to_kw(b):
    kw = H.batch_kwargs(b, dev)
    return kw

  out = {"dropout": 0.0, "batch_per_gpu": B}
  net0, feed0 = build_net(0.0, world > 1)
  feed0.cls = batches[0]["text_feat"].to(dev)
  conf = net0(**to_kw(batches[0]), out="conf", device=dev)["cross_view_conf_matrix"]
  loss = crit(conf)
  loss.backward()
  torch.cuda.synchronize()
  if world == 1:
    from oracle import mmt_oracle as O         # the checker (CPU), not the thing measured
    ed = W.compute_dims(w["modalities"], w["face_dim"])
    cfg = {"expert_dims": ed, "vid_bert_params": vb_params(w, 0.0), "txt_dropout": 0.0, "test_caption_mode": "indep"}
    torch.set_num_threads(min(16, usable_cpus()))
    conf_ref, loss_ref, grads = H.oracle_step(P, batches[0], cfg)
    g_max, g_l2, worst, worst_l2 = H.grad_errors(net0, grads)
    out.update({"against": "oracle port (fp32, CPU) on the same batch and weights",
                "conf_max_rel": H.rel_err(conf, conf_ref), "conf_rel_l2": H.rel_l2(conf, conf_ref),
                "loss_rel": abs(float(loss) - loss_ref) / abs(loss_ref),
                "grad_max_rel": g_max, "grad_rel_l2": g_l2, "worst_tensor": [worst[0], worst[1]]})
    out["pass"] = bool(out["conf_max_rel"] < 1e-3 and out["conf_rel_l2"] < 1e-3 and out["loss_rel"] < 1e-3 and
                       g_max < 1e-3 and g_l2 < 1e-3) if args.precision in ("f16", "tf32", "fp32") else None
  else:
    # every rank regenerates all ranks' first batches (seeds are a function of the rank) -> the global batch
    from mmt_b200.parallel import head_segments
    gb = {}
    parts = [make_batches(w, B, 1, 1234 + 100 * r)[1][0] for r in range(world)]
    for k in parts[0]:
      if isinstance(parts[0][k], dict):
        gb[k] = {m: torch.cat([p_[k][m] for p_ in parts], 0) for m in parts[0][k]}
      else:
        gb[k] = torch.cat([p_[k] for p_ in parts], 0)
    if rank == 0:
      net1, feed1 = build_net(0.0, False)
      feed1.cls = gb["text_feat"].to(dev)
      conf1 = net1(**to_kw(gb), out="conf", device=dev)["cross_view_conf_matrix"]
      loss1 = crit(conf1)
      loss1.backward()
      torch.cuda.synchronize()
      g0, g1 = net0._grad_flat(), net1._grad_flat()
      out.update({"against": "single-device step of the global batch (%d) on rank 0" % (B * world),
                  "conf_max_rel": H.rel_err(conf, conf1), "loss_rel": abs(float(loss) - float(loss1)) / abs(float(loss1)),
                  "grad_max_rel": float((g0 - g1).abs().max() / g1.abs().max()),
                  "grad_rel_l2": float((g0 - g1).norm() / g1.norm())})
      out["pass"] = bool(out["conf_max_rel"] < 1e-4 and out["loss_rel"] < 1e-5 and out["grad_max_rel"] < 1e-3)
      del net1
    dist.barrier()
  del net0
  torch.cuda.empty_cache()
  return out


def measured_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.isfile(p):
    d = json.load(open(p))
    return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
  return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def roofline_ffn(net, w, B, dev, args):
  """Dominant kernel = the FFN-up GEMM+bias+erf-GELU ([B*S,512] x [3072,512]^T), timed alone with
  CUDA events on the launching stream over inputs re-used from L2-cold buffers."""
  from mmt_b200 import _lib
  M_ = len(w["modalities"])
  S = 1 + M_ * (w["T"] + 1)
  BS, d, ff = B * S, 512, 3072
  L = net.layout
  reps = 6
  p = "vid_bert.encoder.layer.0."
  is16 = _lib.is16(net.cfg.precision)
  if is16:
    dt = _lib.dt_of(net.cfg.precision)
    tdt = _lib.torch_dtype(dt)
    net._prepare16()
    a = [torch.randn(BS, d, device=dev).to(tdt) for _ in range(reps)]
    f = [torch.empty(BS, ff, device=dev, dtype=tdt) for _ in range(reps)]
    u = [torch.empty(BS, ff, device=dev, dtype=tdt) for _ in range(reps)]

    def launch(i):
      _lib.gemm16(dt, BS, ff, d, a[i], d, 0, net.cfg.w16.flat16, d, 0, b_off=L.off(p + "intermediate.dense.weight"),
                  bias=net.flat, bias_off=L.off(p + "intermediate.dense.bias"), epilogue=_lib.EPI_GELU, aux16=u[i],
                  aux_ld=ff, C16=f[i], c16_ld=ff)
  else:
    a = [torch.randn(BS, d, device=dev) for _ in range(reps)]
    f = [torch.empty(BS, ff, device=dev) for _ in range(reps)]
    u = [torch.empty(BS, ff, device=dev) for _ in range(reps)]

    def launch(i):
      _lib.gemm(BS, ff, d, a[i], d, 1, net.flat, d, 1, f[i], ff, b_off=L.off(p + "intermediate.dense.weight"),
                bias=net.flat, bias_off=L.off(p + "intermediate.dense.bias"), epilogue=_lib.EPI_GELU,
                aux=u[i], precision=net.cfg.precision)

  for i in range(3):
    launch(i % reps)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n = 12
  e0.record()
  for i in range(n):
    launch(i % reps)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / n
  flops = 2.0 * BS * d * ff
  hbm, bf16, src = measured_peaks()
  peak = bf16 / 2.0 if args.precision == "tf32" else bf16
  ach = flops / (ms / 1e3) / 1e12
  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu capture of this round's build
  prof = profile_facts() or {}
  traffic = (prof.get("ffn_up_gemm") or {}).get("dram_bytes") if (BS == 13952 and args.precision == "f16") else None
  alg = BS * d * 2 + ff * d * 2 + (2 * BS * ff * 2 if is16 else 2 * BS * ff * 4)
  return {"kernel": "FFN-up GEMM+bias+erf-GELU(+GELU') %dx%dx%d (%s operands)" % (BS, ff, d, args.precision),
          "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
          "traffic": traffic, "algorithmic_bytes": alg, "ms_per_launch": ms,
          "peak_note": "%s dense bf16 burst rate (fp16 runs at the same rate; tf32 = half)" % src,
          "traffic_source": "profiles/r02_kernels.json" if traffic else None}


def roofline_maxmargin(dev):
  """north_star's HBM-bound kernel: MaxMarginRankingLoss forward over an N x N matrix >> L2
  (N=16384: 1.07 GB; algorithmic bytes = 4*N^2 read)."""
  from mmt_b200 import engine
  n = 16384
  x = torch.rand(n, n, device=dev) * 2 - 1
  for _ in range(3):
    engine.max_margin(x, 0.05, True, want_grad=False)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 5
  e0.record()
  for _ in range(reps):
    engine.max_margin(x, 0.05, True, want_grad=False)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / reps
  hbm, _, src = measured_peaks()
  ach = 4.0 * n * n / (ms / 1e3) / 1e9
  prof = profile_facts() or {}
  traffic = (prof.get("max_margin_fwd") or {}).get("dram_bytes")
  return {"kernel": "max_margin forward N=16384", "bound": "hbm", "achieved": ach, "peak": hbm,
          "unit": "GB/s", "frac": ach / hbm, "traffic": traffic, "algorithmic_bytes": 4.0 * n * n,
          "ms_per_launch": ms, "peak_note": src, "traffic_source": "profiles/r02_kernels.json" if traffic else None}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_step_fn(w, B, seed=1234):
  """The reference's train step restated on CPU (oracle port): same op sequence as the reference
  modules (unfused GELU, materialised attention, Python token assembly is vectorised)."""
  from oracle import mmt_oracle as O
  import workloads as W
  ed = W.compute_dims(w["modalities"], w["face_dim"])
  vb = vb_params(w)
  P = W.init_params(ed, vb, seed=0)
  params = [v.requires_grad_(True) for k, v in P.items()
            if v.is_floating_point() and "running" not in k and "pooler" not in k]
  opt = torch.optim.Adam(params, lr=LR, weight_decay=WD)
  cfg = {"expert_dims": ed, "vid_bert_params": vb, "txt_dropout": DROPOUT,
         "test_caption_mode": "indep"}
  batch = W.synth_batch(ed, B, w["T"], seed=seed)

  def step():
    opt.zero_grad()
    out = O.cenet_forward(P, batch, cfg, training=True, out="conf", text_feat=batch["text_feat"])
    loss = O.max_margin_ranking_loss(out["cross_view_conf_matrix"], 0.05, True)
    loss.backward()
    opt.step()
    return float(loss.detach())

  return step


def usable_cpus():
  """CPUs this process may actually use: affinity mask capped by the cgroup CPU quota."""
  try:
    n = len(os.sched_getaffinity(0))
  except AttributeError:
    n = os.cpu_count() or 1
  try:
    q, per = open("/sys/fs/cgroup/cpu.max").read().split()
    if q != "max":
      n = max(1, min(n, int(float(q) / float(per) + 0.5)))
  except Exception:
    pass
  return n


def pick_threads(w):
  """The reference does not set a thread count (torch default = all CPUs it sees).  To give the
  CPU arm its best case, probe a few intra-op thread counts on a tiny step of the same model and
  keep the fastest."""
  n = usable_cpus()
  cands = sorted(set([c for c in (4, 8, 16, 32, 64) if c <= n] + [n]))
  small = dict(w, B=4)
  best, best_t = cands[0], None
  for c in cands:
    torch.set_num_threads(c)
    fn = cpu_step_fn(small, 4)
    fn()
    t0 = time.time()
    fn()
    dt = time.time() - t0
    if best_t is None or dt < best_t:
      best, best_t = c, dt
  torch.set_num_threads(best)
  return best


def cpu_baseline(w, steps, warmup, budget_s):
  cores = pick_threads(w)
  B = w["B"]
  fn = cpu_step_fn(w, B)
  t0 = time.time()
  fn()
  first = time.time() - t0
  if first * (steps + warmup) > budget_s and B > 8:      # bound the sample
    B = max(8, int(B * budget_s / (first * (steps + warmup))) // 8 * 8)
    fn = cpu_step_fn(w, B)
    fn()
  for _ in range(max(0, warmup - 1)):
    fn()
  t0 = time.time()
  for _ in range(steps):
    fn()
  dt = (time.time() - t0) / steps
  return {"value": B / dt, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": "%d timed train steps (after %d warm-up) of the same workload at batch %d, oracle port (torch CPU, %d threads)" % (steps, warmup, B, cores),
          "ms_per_step": dt * 1e3}


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  w = WORKLOADS[args.workload]
  cores = pick_threads(w)
  B = w["B"]
  total = args.steps + args.warmup
  fn = cpu_step_fn(w, B)
  t0 = time.time()
  fn()
  first = time.time() - t0
  if first * total > 150.0 and B > 8:
    B = max(8, int(B * 150.0 / (first * total)) // 8 * 8)
    fn = cpu_step_fn(w, B)
  for _ in range(args.warmup):
    fn()
  t0 = time.time()
  for _ in range(args.steps):
    fn()
  dt = (time.time() - t0) / args.steps
  val = B / dt
  sample = "each step = one train step of the workload at batch %d (bounded sample of batch %d), oracle port of the reference on %d host threads" % (B, w["B"], cores)
  print(json.dumps({
      "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": bench_config(w, args.gpus, args.precision),
      "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
  ap.add_argument("--precision", default=os.environ.get("MMT_PRECISION", "f16"),
                  choices=["fp32", "tf32", "f16", "bf16"],
                  help="operand precision of the GEMMs / attention (accumulation, statistics, residuals, gradients "
                       "are fp32 in every mode): f16 = fp16 operands (default; tf32's mantissa at half the bytes), "
                       "bf16 (BASELINE config 5), tf32, fp32 = CUDA-core exact mode")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--graph", action="store_true",
                  help="replay the step as one CUDA graph for `value` (mmt_b200/graph.py); measured gain on "
                       "B200 is < 1 % because the step is GPU-bound, so eager launches are the default")
  ap.add_argument("--no-hbm-probe", action="store_true")
  ap.add_argument("--no-scope-b", action="store_true", help="skip the full-CENet (text encoder included) measurement")
  ap.add_argument("--no-parity-check", action="store_true", help="skip the dropout-free parity / data-parallel check")
  ap.add_argument("--no-torch-adam", action="store_true", help="skip the stock torch.optim.Adam e2e arm")
  ap.add_argument("--trace", action="store_true", help="print a kernel-timeline summary of 3 steps (rank 0)")
  ap.add_argument("--graph-dp", action="store_true",
                  help="also capture the data-parallel step (NCCL inside the CUDA graph) for the e2e arm at N > 1")
  ap.add_argument("--no-graph-e2e", action="store_true",
                  help="e2e through eager CENet.forward only (skip the GraphedTrainStep arm)")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    if args.warmup < 3:
      args.warmup = 3
    run_b200(args)


if __name__ == "__main__":
  main()
