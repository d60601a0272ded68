#!/bin/bash
# Round 2, GPU call 3: all GPU tests after the upload rework; predict() probes; config 5 replay
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for a in "er 1000000" "rmat 20"; do
  timeout 600 python tools/predict_probe.py $a 2>> gpurun_out/r2_probe3.err | tail -1
done
GVC_TRACE=1 timeout 600 python tools/predict_probe.py rmat 20 2> gpurun_out/r2_probe3_trace_rmat20.txt | tail -1
GVC_TRACE=1 timeout 600 python tools/predict_probe.py er 1000000 2> gpurun_out/r2_probe3_trace_er1m.txt | tail -1
timeout 1200 python tools/replay_config5.py 1000000 b200_exact 2>&1 | tail -16
