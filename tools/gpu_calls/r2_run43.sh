#!/bin/bash
# two-GPU sanity of the driver's scaling command with the final code (rows in schedule order + L2 window came after the last multi-GPU run)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu or group_of_devices or sharded_over" 2>&1 | tail -3
for m in exact fast; do
out=gpurun_out/r2_final6_multi_weak_${m}_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 --mode $m > $out 2> gpurun_out/r2_final6_multi_${m}_n2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$out") if l.startswith("{")][-1])
    print("n=2 $m", d["config"]["workload"], "ms/step %.3f"%d["ms_per_step"], "Gedges/s %.2f"%(d["value"]/1e9), "e2e ms %.3f"%d["e2e"]["ms_per_step"], d.get("phase_ms"), "parity", d.get("parity_vs_1gpu"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2_final6_multi_${m}_n2.err").read()[-1500:])
PY
done
