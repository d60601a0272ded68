#!/bin/bash
# usage: bash tools/gpu_variant.sh <tag> "<nvcc extra flags>"   -- rebuild libgvc with extra flags on the GPU box and bench it
tag=$1; shift
GVC_NVCC_EXTRA="$*" python -c "
import sys; sys.path.insert(0,'.')
import gnn_mwvc_b200
from gnn_mwvc_b200 import build
build.build_libgvc(force=True)" || { echo BUILD FAILED; exit 1; }
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || tail -5 gpurun_out/bench_${tag}.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mode fast > gpurun_out/bench_${tag}_fast.json 2>> gpurun_out/bench_${tag}.err
python - <<PY
import json
for f in ("gpurun_out/bench_${tag}.json","gpurun_out/bench_${tag}_fast.json"):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "stage_ms", [round(x,3) for x in r["stage_ms"]], "fwd_frac %.3f"%r["forward_frac"], "e2e ms %.3f"%d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "FAILED", e)
PY
