#!/bin/bash
# Round 2 multi-GPU measurements (usage under `gpurun --gpus N`: bash tools/r2_multi.sh N [tests])
n=${1:-2}
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu or group_of_devices or sharded_over" 2>&1 | tail -5
fi
run() {  # name, extra bench args
  out=gpurun_out/r2_multi_$1_n$n.json
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 ${@:2} > $out 2> gpurun_out/r2_multi_$1_n$n.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$out") if l.startswith("{")][-1])
    print("n=$n $1", d["config"]["workload"], "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "e2e ms %.3f"%d["e2e"]["ms_per_step"], d.get("phase_ms"), "parity", d.get("parity_vs_1gpu"))
except Exception as e:
    print("n=$n $1 FAILED", e); print(open("gpurun_out/r2_multi_$1_n$n.err").read()[-1500:])
PY
}
run weak_exact --mode exact
run weak_fast --mode fast
if [ $n -lt 8 ]; then run strong23_exact --mode exact --scale 23; run strong23_fast --mode fast --scale 23; fi
