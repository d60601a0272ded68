#!/bin/bash
# Round 2 re-baseline after the container was re-created: GPU tests, default bench (as the driver runs it),
# reference arm, the 100 M-edge graph, the grid and dense-only workloads in both modes.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc=$?"
for m in exact fast; do
timeout 900 python bench.py --scale 23 --mode $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_scale23_1gpu_$m.json 2> gpurun_out/r2_scale23_$m.err
done
for w in grid isolated; do for m in exact fast; do
timeout 600 python bench.py --workload $w --mode $m --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_wl_${w}_${m}.err > gpurun_out/r2_bench_${w}_${m}.json
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d.get("roofline") or {}
        print(f, d["config"].get("workload"), d["config"].get("mode"), "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.3f"%(d["value"]/1e9), "stage_ms", [round(x,3) for x in r.get("stage_ms",[])], "frac %.3f fwd_frac %.3f"%(r.get("frac",0), r.get("forward_frac",0)), "e2e", json.dumps(d.get("e2e"))[:300])
    except Exception as e: print(f, "FAILED", e)
PY
tail -3 gpurun_out/r2_bench_default.err
