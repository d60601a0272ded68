#!/bin/bash
# where does a small predict go? (ER 2 000 / 20 000 / 200 000 vertices, stream and packed upload, traced and untraced)
mkdir -p gpurun_out
for n in 2000 20000 200000; do
  for up in "" packed; do
    timeout 200 python tools/predict_probe.py er $n $up 2>/dev/null | tail -1
  done
  GVC_TRACE=1 timeout 200 python tools/predict_probe.py er $n 2>&1 | grep -A12 "predict call 7" | head -14
done > gpurun_out/r2_small_predict.txt 2>&1
cat gpurun_out/r2_small_predict.txt
