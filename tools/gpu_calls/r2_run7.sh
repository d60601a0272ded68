#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_run7.json 2> gpurun_out/r2_bench_run7.err; echo "bench rc=$?"
timeout 900 python bench.py --scale 23 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale23_1gpu_exact_run7.json 2> gpurun_out/r2_scale23_run7.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_run7.json", "gpurun_out/r2_scale23_1gpu_exact_run7.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
        print(f, "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "stage_ms", [round(x,3) for x in r["stage_ms"]], "fwd_frac %.3f"%r["forward_frac"], "e2e ms %.3f first %.1f"%(d["e2e"]["ms_per_step"], d["e2e"].get("first_call_ms", 0)), d["other_mode"]["ms_per_step"])
    except Exception as e: print(f, "FAILED", e)
PY
tail -3 gpurun_out/r2_bench_run7.err
