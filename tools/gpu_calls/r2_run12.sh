#!/bin/bash
mkdir -p gpurun_out
python tools/microbench/dump_rmat_col.py /tmp/col.bin
./tools/microbench/bulk_gather /tmp/col.bin 2>&1 | tee gpurun_out/r2_microbench_bulk_gather_rmat_ids.txt
./tools/microbench/gather_bw /tmp/col.bin 2>&1 | tee gpurun_out/r2_microbench_gather_bw_rmat_ids.txt
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r2_bench_default_v2.json 2> gpurun_out/r2_bench_default_v2.err; echo "bench rc=$?"
timeout 900 python tools/replay_config5.py 1000000 > gpurun_out/r2_config5_er1m.json 2> gpurun_out/r2_config5_er1m.err; tail -c 1500 gpurun_out/r2_config5_er1m.json
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_default_v2.json") if l.startswith("{")][-1]); r=d["roofline"]
print("ms/step %.3f"%d["ms_per_step"], "stage_ms", [round(x,3) for x in r["stage_ms"]], "e2e ms %.3f first %.1f"%(d["e2e"]["ms_per_step"], d["e2e"].get("first_call_ms", 0)), "c_abi %.3f"%d["e2e"]["c_abi"]["ms_per_step"])
PY
