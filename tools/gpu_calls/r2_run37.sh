#!/bin/bash
# threshold of the row reordering lowered to 500 000 vertices: tests, default bench (as the driver runs it), fast mode
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_final4_bench_default.json 2> gpurun_out/r2_final4.err; echo "bench rc=$?"
timeout 900 python bench.py --mode fast --no-cpu-baseline > gpurun_out/r2_final4_bench_default_fast.json 2>/dev/null
python - <<'PY'
import json
for f in ("gpurun_out/r2_final4_bench_default.json", "gpurun_out/r2_final4_bench_default_fast.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
    print(f, "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.3f"%(d["value"]/1e9), [round(x,3) for x in r["stage_ms"]], "frac %.3f fwd %.3f"%(r["frac"], r["forward_frac"]), "e2e %.3f c_abi %.3f resident %.3f"%(d["e2e"]["ms_per_step"], d["e2e"]["c_abi"]["ms_per_step"], d["e2e"]["csr_resident"]["ms_per_step"]), "launches", d["gpu_launches"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 9 -c 3 -o /tmp/default4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_default4.log 2>&1
python tools/ncu_summary.py /tmp/default4.ncu-rep gpurun_out/r2_ncu_stage_kernels.json exact rmat_scale20_ef16 ""
cp profiles/traffic.json gpurun_out/r2_traffic.json
