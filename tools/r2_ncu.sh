#!/bin/bash
# ncu evidence for profiles/: launch list + full capture of the three exact stage kernels, default bench command
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 9 -c 3 -o gpurun_out/r2_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
ls -la gpurun_out/r2_prof.ncu-rep
