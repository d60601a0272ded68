#!/bin/bash
# upload item size decoupled from the ring's slot size, faster gathered copy: tests, predict probe, config 5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "stream or dropin or reference_interface or train" 2>&1 | tail -3
timeout 600 python tools/predict_probe.py rmat 20 2>/dev/null | tail -1
timeout 600 python tools/predict_probe.py er 1000000 2>/dev/null | tail -1
timeout 900 python tools/replay_config5.py 1000000 b200_exact 2>&1 | grep "gvc profile" | head -14
python -c "
import json; d=json.load(open('gpurun_out/config5_n1000000_b200_exact.json'))['b200_exact']; print(d['md5'], d['predict_total_s'], d['profile'])"
