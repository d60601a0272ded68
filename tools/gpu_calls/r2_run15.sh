#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -x -q -k "parallel_exact or bench_size or degree_ladder or forward_vs_oracle" 2>&1 | tail -3
timeout 300 python tools/chain_probe.py 262144 > gpurun_out/r2_chain_probe_262144_v3.json 2> gpurun_out/r2_chain_probe.err; cat gpurun_out/r2_chain_probe_262144_v3.json
timeout 600 python tools/shard_probe.py 23 8 > gpurun_out/r2_shard_probe_23_8_v3.json 2> gpurun_out/r2_shard_probe.err; cat gpurun_out/r2_shard_probe_23_8_v3.json
