run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) bench.py --gpus 8 --steps 15 --warmup 4 --no-parity-check --no-hbm-probe > gpurun_out/bench_8gpu_$tag.log 2> gpurun_out/bench_8gpu_$tag.err; python - <<PY
import json
try:
  l=[x for x in open("gpurun_out/bench_8gpu_$tag.log") if x.startswith("{")]
  d=json.loads(l[-1]); print("$tag", "value %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
  print("$tag FAILED", e)
PY
}
run default A=1
run simple NCCL_PROTO=Simple
run nvls NCCL_ALGO=NVLS,Ring
