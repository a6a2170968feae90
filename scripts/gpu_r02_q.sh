#!/bin/bash
# Round-2 GPU call Q: one full ncu capture (source-level stall samples) of the current hero kernel.
set -u
O=gpurun_out
mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 1 -c 1 \
    -o $O/prof_r02q_hero python scripts/run_once.py cfg2 8 2 > $O/prof_r02q_hero.log 2>&1; echo "ncu rc=$?"
ls -la $O | tail -5
