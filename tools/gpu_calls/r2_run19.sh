#!/bin/bash
# tuning variants (built in the container into gnn-mwvc_b200/_variants, selected with GVC_LIB) at both sizes, exact mode
mkdir -p gpurun_out
for tag in base mid3 mid4 heavy3 heavy5 heavy6; do
  if [ $tag = base ]; then unset GVC_LIB; else export GVC_LIB=$PWD/gnn-mwvc_b200/_variants/libgvc_$tag.so; fi
  for sc in 0 23; do
    timeout 600 python bench.py --scale $sc --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r2_variant_${tag}_$sc.json
    python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/r2_variant_${tag}_$sc.json') if l.startswith('{')][-1]); r=d['roofline']; print('$tag', d['config']['workload'], 'ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']], 'fwd_frac %.3f'%r['forward_frac'], 'fast', round(d['other_mode']['ms_per_step'],3))" || echo "$tag $sc FAILED"
  done
done
