#!/bin/bash
# bench lines of record for the large workloads after the rows went into schedule order, + ncu of scale 23
mkdir -p gpurun_out
for m in exact fast; do
timeout 900 python bench.py --scale 23 --mode $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final3_bench_scale23_1gpu_$m.json 2>/dev/null
timeout 900 python bench.py --workload grid --mode $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final3_bench_grid_$m.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_final3_bench_*.json")):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
    print(f, "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.3f"%(d["value"]/1e9), [round(x,3) for x in r["stage_ms"]], "frac %.3f fwd %.3f"%(r["frac"], r["forward_frac"]), "launches", d["gpu_launches"])
PY
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 9 -c 3 -o /tmp/scale23 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scale 23 > gpurun_out/r2_ncu_scale23b.log 2>&1
python tools/ncu_summary.py /tmp/scale23.ncu-rep gpurun_out/r2_ncu_stage_kernels_scale23.json exact rmat_scale23_ef16 "--scale 23"
cp profiles/traffic.json gpurun_out/r2_traffic.json
