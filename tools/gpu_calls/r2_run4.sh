#!/bin/bash
# Round 2, GPU call 4: parallel exact hub sums -- tests, chain probe, bench at scale 20 and 23
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/chain_probe.py 262144 > gpurun_out/r2_chain_probe_262144_px.json 2> gpurun_out/r2_chain_probe4.err; cat gpurun_out/r2_chain_probe_262144_px.json
timeout 300 python tools/chain_probe.py 65536 200000 > gpurun_out/r2_chain_probe_65536_busy_px.json 2>> gpurun_out/r2_chain_probe4.err; cat gpurun_out/r2_chain_probe_65536_busy_px.json
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_run4.json 2> gpurun_out/r2_bench_run4.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_run4.json",):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
        print(f, "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "stage_ms", [round(x,3) for x in r["stage_ms"]], "fwd_frac %.3f"%r["forward_frac"], "e2e ms %.3f first %.1f"%(d["e2e"]["ms_per_step"], d["e2e"].get("first_call_ms", 0)), "c_abi %.3f"%d["e2e"]["c_abi"]["ms_per_step"], d["other_mode"])
    except Exception as e: print(f, "FAILED", e)
PY
tail -3 gpurun_out/r2_bench_run4.err
timeout 900 python bench.py --scale 23 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale23_1gpu_exact_px.json 2> gpurun_out/r2_scale23_px.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_scale23_1gpu_exact_px.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("scale23", "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "stage_ms", [round(x,3) for x in r["stage_ms"]], "fwd_frac %.3f"%r["forward_frac"], d["other_mode"])
except Exception as e: print("scale23 FAILED", e)
PY
for t in 4 8 16; do GVC_UPLOAD_THREADS=$t GVC_TRACE=1 timeout 300 python tools/predict_probe.py rmat 20 2> gpurun_out/r2_probe4_trace_t$t.txt | tail -1; grep "workers" gpurun_out/r2_probe4_trace_t$t.txt | tail -1; done
