#!/bin/bash
# Final single-GPU measurements of round 2: tests, bench lines, ncu launch list + full captures (default workload,
# the 100 M-edge graph, the grid).  A number printed under ncu is never a bench value.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2_final_bench_default.json 2> gpurun_out/r2_final_bench_default.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference_arm.json 2> gpurun_out/r2_final_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --mode fast --no-cpu-baseline > gpurun_out/r2_final_bench_default_fast.json 2>/dev/null
for m in exact fast; do
timeout 900 python bench.py --scale 23 --mode $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_bench_scale23_1gpu_$m.json 2> gpurun_out/r2_final_scale23_$m.err
done
for w in grid isolated; do for m in exact fast; do
timeout 600 python bench.py --workload $w --mode $m --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r2_final_bench_${w}_${m}.json
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_final_bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d.get("roofline") or {}
        print(f, d["config"].get("workload"), d["config"].get("mode"), "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.3f"%(d["value"]/1e9), "stage_ms", [round(x,3) for x in r.get("stage_ms",[])], "frac %.3f fwd_frac %.3f"%(r.get("frac",0), r.get("forward_frac",0)), "e2e ms", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e: print(f, "FAILED", e)
PY
# ncu: launch list of the default command, then full captures, condensed on the box (the reports are 30 MB each
# and gpurun_out/ brings back 64 MiB at most: only the default workload's report travels)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; tail -1 gpurun_out/r2_ncu_launches.log | cut -c1-200
cap() {  # name, workload key, bench args
  ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 9 -c 3 -o /tmp/$1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline $3 > gpurun_out/r2_ncu_$1.log 2>&1
  tail -1 gpurun_out/r2_ncu_$1.log | cut -c1-200
  python tools/ncu_summary.py /tmp/$1.ncu-rep gpurun_out/r2_ncu_stage_kernels_$1.json exact $2 "$3"
}
cap default rmat_scale20_ef16 ""
cap scale23 rmat_scale23_ef16 "--scale 23"
cap grid grid_4472x4472 "--workload grid"
cp profiles/traffic.json gpurun_out/r2_traffic.json
cp /tmp/default.ncu-rep gpurun_out/r2_final_prof_default.ncu-rep
ls -la gpurun_out/
