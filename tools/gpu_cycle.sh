#!/bin/bash
# One GPU iteration: parity tests, bench (exact + fast), then an ncu capture of the stage kernels.
# usage (under gpurun): bash tools/gpu_cycle.sh <tag> [ncu]
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || tail -5 gpurun_out/bench_${tag}.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mode fast > gpurun_out/bench_${tag}_fast.json 2>> gpurun_out/bench_${tag}.err
python - <<PY
import json
for f in ("gpurun_out/bench_${tag}.json","gpurun_out/bench_${tag}_fast.json"):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "stage_ms", [round(x,3) for x in r["stage_ms"]], "fwd_frac %.3f"%r["forward_frac"], "e2e ms %.3f (upload %.3f, %.1f GB/s; resident %.3f)"%(d["e2e"]["ms_per_step"], d["e2e"]["graph_upload_ms"], d["e2e"]["graph_upload_GBps"], d["e2e"]["csr_resident"]["ms_per_step"]), d["clocks"])
    except Exception as e: print(f, "FAILED", e)
PY
if [ "$2" = "ncu" ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${tag}.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 9 -c 3 -o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${tag}.log 2>&1
  tail -2 gpurun_out/ncu_${tag}.log
fi
if [ "$2" = "ncu" ]; then
  # launch list of the same command (per-launch device time; cold-cache and serialised: compare shares)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${tag}.log 2>&1
  tail -1 gpurun_out/ncu_launches_${tag}.log
fi
