#!/bin/bash
# final verification of the round's last code: GPU suite, smoke, the driver's bench command and reference arm; ncu capture of the
# 100 M-edge graph with the L2 window (DRAM bytes against the capture without it)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final6_bench_default.json 2> gpurun_out/r2_final6_bench_default.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_final6_bench_reference_arm.json 2> gpurun_out/r2_final6_ref.err
python - <<'PY'
import json
for f in ["gpurun_out/r2_final6_bench_default.json","gpurun_out/r2_final6_bench_reference_arm.json"]:
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.4f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e", d["e2e"].get("ms_per_step"), d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("forward_frac"), d.get("clocks"), d.get("cpu_baseline",{}).get("value"), "launches", d.get("gpu_launches"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scale 23 > gpurun_out/r2_final6_plain23.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 9 -c 3 -o gpurun_out/r2_final6_prof_scale23 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scale 23 > gpurun_out/r2_final6_ncu23.log 2>&1
tail -2 gpurun_out/r2_final6_ncu23.log
ls -la gpurun_out/r2_final6_prof_scale23.ncu-rep
