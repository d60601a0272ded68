#!/bin/bash
# 1 -> N GPU weak-scaling run, as the driver does it (usage under `gpurun --gpus 8`: bash tools/scale_run.sh 8)
max=${1:-8}
for n in 1 2 4 8; do
  [ $n -gt $max ] && break
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --mode ${MODE:-exact} > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 --mode ${MODE:-exact} > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print("n=$n", d["config"]["workload"], "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "e2e ms %.3f"%d["e2e"]["ms_per_step"], d.get("phase_ms"))
except Exception as e:
    print("n=$n FAILED", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
