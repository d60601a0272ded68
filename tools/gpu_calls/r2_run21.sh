#!/bin/bash
mkdir -p gpurun_out
for tag in base walk9; do
  if [ $tag = base ]; then unset GVC_LIB; else export GVC_LIB=$PWD/gnn-mwvc_b200/_variants/libgvc_$tag.so; fi
  echo "== $tag"
  timeout 900 python -m pytest tests -m gpu -x -q -k "parallel_exact or bench_size or degree_ladder" 2>&1 | tail -2
  timeout 300 python tools/chain_probe.py 262144 2> gpurun_out/r2_chain_probe.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('star', d['exact_ms'], [p['walk_kcycles'] for p in d['px']])"
  timeout 600 python tools/shard_probe.py 23 8 2> gpurun_out/r2_shard_probe.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('shard', d['exact'], d['fast'])"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_wl.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; print(d['config']['workload'], 'ms %.3f'%d['ms_per_step'], [round(x,3) for x in r['stage_ms']])"
done
