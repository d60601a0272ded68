#!/bin/bash
# One point of the weak-scaling curve, exact and fast mode (usage under `gpurun --gpus N`: bash tools/scale_one.sh N)
n=${1:-2}
for m in exact fast; do
  out=gpurun_out/scale_${m}_n$n.json
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --mode $m > $out 2> gpurun_out/scale_${m}_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 --mode $m > $out 2> gpurun_out/scale_${m}_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$out") if l.startswith("{")][-1])
    print("n=$n $m", d["config"]["workload"], "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "e2e ms %.3f"%d["e2e"]["ms_per_step"], d.get("phase_ms"))
except Exception as e:
    print("n=$n $m FAILED", e); print(open("gpurun_out/scale_${m}_n$n.err").read()[-1500:])
PY
done
