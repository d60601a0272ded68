#!/bin/bash
# persisting L2 window on the hot rows (48 MB default): tests, bench lines of the large workloads and the default one
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in exact fast; do
timeout 900 python bench.py --scale 23 --mode $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final5_bench_scale23_1gpu_$m.json 2>/dev/null
timeout 900 python bench.py --workload grid --mode $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final5_bench_grid_$m.json 2>/dev/null
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final5_bench_default_quick.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_final5_bench_*.json")):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
    print(f, "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.3f"%(d["value"]/1e9), [round(x,3) for x in r["stage_ms"]], "frac %.3f fwd %.3f"%(r["frac"], r["forward_frac"]))
PY
