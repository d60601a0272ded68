#!/bin/bash
mkdir -p gpurun_out
./tools/microbench/fp32_rate 2000 | tee gpurun_out/r2_microbench_fp32_rate.txt
GVC_TRACE=1 timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20_traced_v2.json 2> gpurun_out/r2_predict_probe_rmat20_traced_v2.err
timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20_v2.json 2> /dev/null
GVC_SLOT_KB=2048 timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20_v2_2mb.json 2>/dev/null
GVC_SLOT_KB=512 timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20_v2_512k.json 2>/dev/null
GVC_UPLOAD_THREADS=12 timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20_v2_t12.json 2>/dev/null
cat gpurun_out/r2_predict_probe_rmat20*v2*.json
grep -A12 "predict call 5" gpurun_out/r2_predict_probe_rmat20_traced_v2.err | head -30
timeout 1800 python -m pytest tests -m gpu -x -q -k "stream or dropin or reference_interface" 2>&1 | tail -8
