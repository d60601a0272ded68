#!/usr/bin/env python
"""Condense an ncu report of the stage kernels into profiles/: key metrics per kernel
(profiles/<name>.json) and DRAM bytes per launch (profiles/traffic.json, read by bench.py).
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/r2_ncu_stage_kernels_X.json [exact|fast] [workload name of the bench line, default rmat_scale20_ef16] [bench arguments, for the record]"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_active.min",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        # pipe shares (the tensor pipe is what the fast mode's mma.sync layers run on)
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
rep, out = sys.argv[1], sys.argv[2]
mode = sys.argv[3] if len(sys.argv) > 3 else "exact"
workload = sys.argv[4] if len(sys.argv) > 4 else "rmat_scale20_ef16"
bench_args = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
name_i = hdr.index("Kernel Name")
kernels = []
for r in data:
    k = {"kernel": r[name_i]}
    for m in KEEP:
        if m in hdr:
            i = hdr.index(m)
            k[m] = {"value": r[i], "unit": units[i]}
    kernels.append(k)
Path(out).write_text(json.dumps({
    "source": "ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 9 -c 3 python bench.py "
              "--steps 2 --warmup 3 --no-cpu-baseline " + bench_args + " (the .ncu-rep stays in gpurun_out/)",
    "workload": workload,
    "note": "one kernel at a time, cold L2 after the bench's 256 MiB flush; durations are not bench values",
    "kernels": kernels}, indent=1))


def to_bytes(cell):
    v, u = float(cell["value"]), cell["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


tp = ROOT / "profiles" / "traffic.json"
traffic = json.loads(tp.read_text()) if tp.exists() else {}
traffic = {k: v for k, v in traffic.items() if isinstance(v, dict)}          # keyed by workload since round 2
for st, k in enumerate(kernels[:3]):
    traffic.setdefault(workload, {})[f"stage{st}_{mode}"] = int(to_bytes(k["dram__bytes_read.sum"]) + to_bytes(k["dram__bytes_write.sum"]))
tp.write_text(json.dumps(traffic, indent=1))
for k in kernels:
    print(k["kernel"][:40], k["gpu__time_duration.sum"]["value"], "us  dram",
          k["dram__bytes_read.sum"]["value"], k["dram__bytes_read.sum"]["unit"], "+", k["dram__bytes_write.sum"]["value"],
          "issue%", k["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"],
          "cycles min/avg/max", k["sm__cycles_active.min"]["value"], k["sm__cycles_active.avg"]["value"], k["sm__cycles_active.max"]["value"])
