#!/bin/bash
# Round-2 GPU call F: hero kernel v5 (layer-2 epilogue on the MMA warpgroup).
set -u
O=gpurun_out
mkdir -p $O
LIBDIR=$PWD/simplerecon_b200/lib
echo "== 0. smoke (default build = v5, setmaxnreg 112/120/24)"
timeout 120 python __graft_entry__.py smoke > $O/r02f_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02f_smoke.log
DEF=$LIBDIR/libsrcv_b200.so
if [ $rc -ne 0 ]; then
  echo "default build failed the smoke: falling back to the no-setmaxnreg build"
  DEF=$LIBDIR/libsrcv_b200_noreg.so
  SRCV_B200_LIB=$DEF timeout 120 python __graft_entry__.py smoke > $O/r02f_smoke_noreg.log 2>&1; rc2=$?; echo "noreg smoke rc=$rc2"; tail -n 1 $O/r02f_smoke_noreg.log
  if [ $rc2 -ne 0 ]; then echo "both v3 builds fail: stopping"; exit 1; fi
fi
echo "== 1. hero A/B (cfg2, B=8, 30 steps): v5 default / v5 noreg / v5 contig"
for lib in default noreg contig; do
  if [ $lib = default ]; then L=$DEF; else L=$LIBDIR/libsrcv_b200_$lib.so; fi
  SRCV_B200_LIB=$L timeout 120 python bench.py --workload cfg2 --steps 30 --warmup 3 --no-cpu-baseline --no-also 2>$O/r02f_hero_$lib.err | tail -n 1 > $O/r02f_hero_$lib.json
  python -c "import json; d=json.load(open('$O/r02f_hero_$lib.json')); print('$lib', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['sweep_us_per_launch'],1), round(d['e2e']['value'],1), d['clocks'])" || tail -n 2 $O/r02f_hero_$lib.err
done
export SRCV_B200_LIB=$DEF
echo "== 2. parity on v3: tcgen05 tests + hero / golden / training-contract parity tests"
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -m gpu -x --timeout 200 -k "tc or golden or hero or mlp or autocast or strided or per_frame or shard" > $O/r02f_parity.log 2>&1; echo "rc=$?"; tail -n 4 $O/r02f_parity.log
timeout 300 python -m pytest tests/test_gpu_producer.py tests/test_gpu_tsdf.py -q -m gpu -x --timeout 200 > $O/r02f_producer_tsdf.log 2>&1; echo "producer+tsdf rc=$?"; tail -n 3 $O/r02f_producer_tsdf.log
echo "== 3. ncu: hero kernel v3, full sections + source (B=4)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 1 -c 1 \
    -o $O/prof_r02f_hero python scripts/run_once.py cfg2 4 2 > $O/r02f_ncu_hero.log 2>&1; echo "ncu rc=$?"
ls -la $O | tail -n 6
echo "== 4. TSDF bench after the 32-bit index fix"
for cfg in "0.01 1" "0.01 4"; do
  set -- $cfg
  timeout 150 python scripts/bench_tsdf.py --voxel $1 --frames $2 --steps 20 2>$O/r02f_tsdf.err | tail -n 1 | tee -a $O/r02f_tsdf.jsonl | head -c 420; echo
done
