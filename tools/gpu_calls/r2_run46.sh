#!/bin/bash
# small predicts with the single-thread streamed upload (GVC_UPLOAD_SINGLE_KB: 0 = helper threads as before, default 2048), then the GPU suite
mkdir -p gpurun_out
for n in 2000 20000 50000 100000; do
  for kb in 0 2048 8192; do
    echo -n "single_kb=$kb "; GVC_UPLOAD_SINGLE_KB=$kb timeout 200 python tools/predict_probe.py er $n 2>/dev/null | tail -1
  done
done > gpurun_out/r2_small_predict_single.txt 2>&1
cat gpurun_out/r2_small_predict_single.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
