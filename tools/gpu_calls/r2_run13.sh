#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/chain_probe.py 262144 > gpurun_out/r2_chain_probe_262144.json 2> gpurun_out/r2_chain_probe.err; cat gpurun_out/r2_chain_probe_262144.json
timeout 600 python tools/shard_probe.py 23 8 > gpurun_out/r2_shard_probe_23_8.json 2> gpurun_out/r2_shard_probe.err; cat gpurun_out/r2_shard_probe_23_8.json
timeout 600 python tools/shard_probe.py 23 4 > gpurun_out/r2_shard_probe_23_4.json 2>> gpurun_out/r2_shard_probe.err; cat gpurun_out/r2_shard_probe_23_4.json
tail -3 gpurun_out/r2_shard_probe.err
