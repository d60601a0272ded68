#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for w in isolated grid rmat; do for m in fast exact; do
timeout 600 python bench.py --workload $w --mode $m --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_wl8.err > gpurun_out/r2_bench8_${w}_${m}.json; python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/r2_bench8_${w}_${m}.json') if l.startswith('{')][-1]); r=d['roofline']; print(d['config']['workload'], d['config']['mode'], 'ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']], 'Gedges/s %.2f'%(d['value']/1e9), 'fwd_frac %.3f'%r['forward_frac'])" || tail -3 gpurun_out/r2_wl8.err
done; done
