"""Turns the ncu reports of a round (gpurun_out/*.ncu-rep) into the committed evidence under profiles/:
  profiles/<round>_<name>_ncu_raw.csv   the raw-page rows of the kernels of interest (selected metrics)
  profiles/<round>_kernels.json         per kernel: duration, DRAM bytes, tensor-pipe %, ... (read by bench.py)
usage: python tools/ncu_extract.py r02 gpurun_out/prof_a.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
           "sm__inst_executed_pipe_xu.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def rows_of(rep):
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  r = list(csv.reader(io.StringIO(out)))
  return r[0], r[1], r[2:]


def main():
  tag, reps = sys.argv[1], sys.argv[2:]
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  facts = {}
  for rep in reps:
    H, U, rows = rows_of(rep)
    idx = {h: i for i, h in enumerate(H)}
    keep = ["ID", "Kernel Name"] + [m for m in METRICS if m in idx]
    name = os.path.basename(rep).replace(".ncu-rep", "").replace("prof_", "")
    with open(os.path.join(root, "profiles", "%s_%s_ncu_raw.csv" % (tag, name)), "w", newline="") as f:
      wr = csv.writer(f)
      wr.writerow(keep)
      wr.writerow([U[idx[k]] for k in keep])
      for r in rows:
        wr.writerow([r[idx[k]] for k in keep])
    for n, r in enumerate(rows):
      kn = r[idx["Kernel Name"]]
      short = kn.split("(")[0].split("::")[-1].replace("<unnamed>", "")
      def val(m):
        if m not in idx or r[idx[m]] == "":
          return None
        v = float(r[idx[m]].replace(",", ""))
        return v * UNIT.get(U[idx[m]], 1.0)
      facts.setdefault(name, []).append({
          "kernel": short, "grid": r[idx["launch__grid_size"]] if "launch__grid_size" in idx else None,
          "duration_us": val("gpu__time_duration.sum"),
          "dram_bytes": (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0),
          "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
          "tensor_pipe_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
          "sm_throughput_pct": val("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
          "dram_throughput_pct": val("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
          "registers": val("launch__registers_per_thread")})
  print(json.dumps(facts, indent=1))


if __name__ == "__main__":
  main()
