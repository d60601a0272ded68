#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parallel_exact or bench_size or degree_ladder or forward_vs_oracle" 2>&1 | tail -5
timeout 300 python tools/chain_probe.py 262144 > gpurun_out/r2_chain_probe_262144_v2.json 2> gpurun_out/r2_chain_probe.err; cat gpurun_out/r2_chain_probe_262144_v2.json
timeout 300 python tools/chain_probe.py 65536 200000 > gpurun_out/r2_chain_probe_65536_busy_v2.json 2>> gpurun_out/r2_chain_probe.err; cat gpurun_out/r2_chain_probe_65536_busy_v2.json
timeout 600 python tools/shard_probe.py 23 8 > gpurun_out/r2_shard_probe_23_8_v2.json 2> gpurun_out/r2_shard_probe.err; cat gpurun_out/r2_shard_probe_23_8_v2.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; print('scale20 ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']])"
