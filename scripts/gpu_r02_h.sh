#!/bin/bash
# Round-2 GPU call H: hero v6 (layer 1 in two N = 64 halves pipelined with the epilogue; pose bias staged in shared memory).
set -u
O=gpurun_out
mkdir -p $O
LIBDIR=$PWD/simplerecon_b200/lib
echo "== 0. smoke"
timeout 120 python __graft_entry__.py smoke > $O/r02h_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02h_smoke.log
[ $rc -ne 0 ] && exit 1
echo "== 1. hero A/B (cfg2, B=8, 30 steps)"
for lib in default noreg; do
  if [ $lib = default ]; then L=$LIBDIR/libsrcv_b200.so; else L=$LIBDIR/libsrcv_b200_$lib.so; fi
  SRCV_B200_LIB=$L timeout 120 python bench.py --workload cfg2 --steps 30 --warmup 3 --no-cpu-baseline --no-also 2>$O/r02h_hero_$lib.err | tail -n 1 > $O/r02h_hero_$lib.json
  python -c "import json; d=json.load(open('$O/r02h_hero_$lib.json')); print('$lib', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['sweep_us_per_launch'],1), round(d['e2e']['value'],1), d['clocks'])" || tail -n 2 $O/r02h_hero_$lib.err
done
echo "== 2. ncu: default, full sections + source (B=4)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 1 -c 1 \
    -o $O/prof_r02h_hero python scripts/run_once.py cfg2 4 2 > $O/r02h_ncu_hero.log 2>&1; echo "ncu rc=$?"
ls -la $O | tail -n 4
