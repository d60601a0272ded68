#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/chain_probe.py 262144 > gpurun_out/r2_chain_probe_262144_v4.json 2> gpurun_out/r2_chain_probe.err; cat gpurun_out/r2_chain_probe_262144_v4.json
timeout 600 python tools/shard_probe.py 23 8 > gpurun_out/r2_shard_probe_23_8_v4.json 2> gpurun_out/r2_shard_probe.err; cat gpurun_out/r2_shard_probe_23_8_v4.json
for w in rmat grid isolated; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_wl.err > gpurun_out/r2_bench_v3_${w}_exact.json; python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/r2_bench_v3_${w}_exact.json') if l.startswith('{')][-1]); r=d['roofline']; print(d['config']['workload'], d['config']['mode'], 'ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']], 'Gedges/s %.2f'%(d['value']/1e9), 'fwd_frac %.3f'%r['forward_frac'], 'other', d['other_mode']['ms_per_step'])" || tail -3 gpurun_out/r2_wl.err
done
timeout 900 python bench.py --scale 23 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r2_bench_v3_scale23_exact.json; python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/r2_bench_v3_scale23_exact.json') if l.startswith('{')][-1]); r=d['roofline']; print(d['config']['workload'], d['config']['mode'], 'ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']], 'Gedges/s %.2f'%(d['value']/1e9), 'fwd_frac %.3f'%r['forward_frac'], 'frac %.3f'%r['frac'])"
