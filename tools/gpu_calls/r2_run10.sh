#!/bin/bash
# Where does predict() go at the bench size?  (trace = synchronised steps; plain = wall time per call)
mkdir -p gpurun_out
GVC_TRACE=1 timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20_traced.json 2> gpurun_out/r2_predict_probe_rmat20_traced.err
timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20.json 2> gpurun_out/r2_predict_probe_rmat20.err
GVC_UPLOAD_THREADS=16 timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20_t16.json 2>/dev/null
GVC_UPLOAD_THREADS=4 timeout 600 python tools/predict_probe.py rmat 20 > gpurun_out/r2_predict_probe_rmat20_t4.json 2>/dev/null
cat gpurun_out/r2_predict_probe_rmat20*.json
grep -A40 "predict call 5" gpurun_out/r2_predict_probe_rmat20_traced.err | head -60
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA"
