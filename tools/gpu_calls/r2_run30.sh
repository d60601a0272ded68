#!/bin/bash
# low-degree tile gather variants (built into gnn-mwvc_b200/_variants): grid and R-MAT, both modes; parity of the variants
mkdir -p gpurun_out
for tag in base tile3 tile2; do
  if [ $tag = base ]; then unset GVC_LIB; else export GVC_LIB=$PWD/gnn-mwvc_b200/_variants/libgvc_$tag.so; fi
  echo "== $tag"
  [ $tag != base ] && timeout 900 python -m pytest tests -m gpu -x -q -k "golden_vectors or bench_size or degree_ladder or forward_vs_oracle" 2>&1 | tail -1
  for w in grid rmat; do
    timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_wl.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; print(d['config']['workload'], 'exact ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']], 'fast ms %.3f'%d['other_mode']['ms_per_step'])"
  done
done
