#!/bin/bash
# final check of the tree as committed: tests, smoke, default bench + reference arm; fast-mode ncu captures (tensor pipe share)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_final2_bench_default.json 2> gpurun_out/r2_final2_bench_default.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final2_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_final2_bench_default.json") if l.startswith("{")][-1]); r=d["roofline"]
print("ms/step %.3f"%d["ms_per_step"], "stage_ms", [round(x,3) for x in r["stage_ms"]], "frac %.3f fwd %.3f"%(r["frac"], r["forward_frac"]), "traffic", r["traffic"], "e2e ms %.3f first %.1f"%(d["e2e"]["ms_per_step"], d["e2e"].get("first_call_ms", 0)), "c_abi %.3f"%d["e2e"]["c_abi"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"])
PY
cap() {  # name, workload key, bench args
  ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 9 -c 3 -o /tmp/$1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline $3 > gpurun_out/r2_ncu_$1.log 2>&1
  tail -1 gpurun_out/r2_ncu_$1.log | cut -c1-120
  python tools/ncu_summary.py /tmp/$1.ncu-rep gpurun_out/r2_ncu_stage_kernels_$1.json fast $2 "$3"
}
cap default_fast rmat_scale20_ef16 "--mode fast"
cap grid_fast grid_4472x4472 "--workload grid --mode fast"
cp profiles/traffic.json gpurun_out/r2_traffic.json
