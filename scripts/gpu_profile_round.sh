#!/bin/bash
# Runs on the GPU box (through gpurun): launch lists + one full ncu capture of each dominant
# kernel, for the bench.py command lines.  Outputs under gpurun_out/; summarise with
# scripts/summarise_profiles.py in the build container and commit the result under profiles/.
set -u
R=${1:-r01}
mkdir -p gpurun_out
# 1. every launch with its device time, same command as the bench (short)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/launches_${R}_cfg1.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${R}_cfg1.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_${R}_cfg2.csv python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_${R}_cfg2.log 2>&1
# 2. the dominant kernels, full sections, at the bench batch sizes
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dot_fast -s 2 -c 1 \
    -o gpurun_out/prof_${R}_dot python scripts/run_once.py cfg1 4 3 > gpurun_out/prof_${R}_dot.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 1 -c 1 \
    -o gpurun_out/prof_${R}_hero python scripts/run_once.py cfg2 8 2 > gpurun_out/prof_${R}_hero.log 2>&1
ls -la gpurun_out | tail -12
