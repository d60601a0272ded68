#!/bin/bash
# Round 2, GPU call 2: all GPU tests; default bench + reference arm; predict() end to end (probe, config 5 replay)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_run2.json 2> gpurun_out/r2_bench_run2.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2_bench_run2.json; tail -5 gpurun_out/r2_bench_run2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_ref_run2.json 2> gpurun_out/r2_bench_ref_run2.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_ref_run2.json
for a in "er 1000000" "er 1000000 packed" "rmat 20" "rmat 20 packed"; do
  timeout 600 python tools/predict_probe.py $a 2>> gpurun_out/r2_probe.err | tail -1
done
GVC_TRACE=1 timeout 600 python tools/predict_probe.py rmat 20 2> gpurun_out/r2_probe_trace_rmat20.txt | tail -1
GVC_TRACE=1 timeout 600 python tools/predict_probe.py er 1000000 2> gpurun_out/r2_probe_trace_er1m.txt | tail -1
timeout 1200 python tools/replay_config5.py 1000000 b200_exact,b200_exact_packed,reference_cpu 2>&1 | grep -v "^gvc profile: predict" | tail -8
