#!/bin/bash
# rows in schedule order: forced on small graphs (tests), then the whole suite with the default threshold and with
# the renumbering forced for every graph; scale-23 and default bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_row_order.py -x -q 2>&1 | tail -12
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
GVC_ROW_ORDER_MIN_VERTICES=0 GVC_ROW_ORDER_AFTER=0 timeout 2400 python -m pytest tests -m gpu -q -k "not default_threshold and not forwarded_again" 2>&1 | tail -6
for a in "--scale 23" "--scale 23 --mode fast" "" "--workload grid"; do
timeout 900 python bench.py $a --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_wl.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; print(d['config']['workload'], d['config']['mode'], 'ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']], 'Gedges/s %.2f'%(d['value']/1e9), 'frac %.3f fwd_frac %.3f'%(r['frac'], r['forward_frac']), 'e2e', round(d['e2e']['ms_per_step'],2))" || tail -3 gpurun_out/r2_wl.err
done
