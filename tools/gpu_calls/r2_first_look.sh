#!/bin/bash
# Round 2, first GPU call: measurements the plan depends on (nothing here is a bench value of record).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_first_smi.txt
timeout 120 tools/microbench/bulk_gather > gpurun_out/r2_bulk_gather.txt 2>&1; echo "bulk_gather rc=$?"
timeout 120 tools/microbench/gather_bw > gpurun_out/r2_gather_bw.txt 2>&1; echo "gather_bw rc=$?"
timeout 600 python -m pytest tests/test_gpu_zz_reference_interface.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/chain_probe.py 262144 > gpurun_out/r2_chain_probe_262144.json 2> gpurun_out/r2_chain_probe.err; cat gpurun_out/r2_chain_probe_262144.json
timeout 300 python tools/chain_probe.py 65536 200000 > gpurun_out/r2_chain_probe_65536_busy.json 2>> gpurun_out/r2_chain_probe.err; cat gpurun_out/r2_chain_probe_65536_busy.json
timeout 600 python bench.py --scale 23 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale23_1gpu_exact_before.json 2> gpurun_out/r2_scale23.err; tail -c 1500 gpurun_out/r2_scale23_1gpu_exact_before.json
for m in exact fast; do
timeout 600 python bench.py --workload grid --mode $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_grid_${m}_before.json 2> gpurun_out/r2_grid_${m}.err; tail -c 900 gpurun_out/r2_grid_${m}_before.json
done
