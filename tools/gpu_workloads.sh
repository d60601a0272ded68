#!/bin/bash
# bench the diagnostic workloads (dense-only, grid) in both modes
for w in isolated grid; do for m in exact fast; do
python bench.py --workload $w --mode $m --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/wl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['config']['workload'], d['config']['mode'], 'ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']], 'Gedges/s %.2f'%(d['value']/1e9), 'fwd_frac %.3f'%r['forward_frac'])"
done; done
