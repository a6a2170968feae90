#!/bin/bash
# Round-2 GPU call E: hero kernel v4 (double-buffered A1, K1 = 192) vs v3.
set -u
O=gpurun_out
mkdir -p $O
LIBDIR=$PWD/simplerecon_b200/lib
echo "== 0. smoke (default build = v4, setmaxnreg 112/120/24)"
timeout 120 python __graft_entry__.py smoke > $O/r02e_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02e_smoke.log
DEF=$LIBDIR/libsrcv_b200.so
if [ $rc -ne 0 ]; then
  echo "default build failed the smoke: falling back to the no-setmaxnreg build"
  DEF=$LIBDIR/libsrcv_b200_noreg.so
  SRCV_B200_LIB=$DEF timeout 120 python __graft_entry__.py smoke > $O/r02e_smoke_noreg.log 2>&1; rc2=$?; echo "noreg smoke rc=$rc2"; tail -n 1 $O/r02e_smoke_noreg.log
  if [ $rc2 -ne 0 ]; then echo "both v3 builds fail: stopping"; exit 1; fi
fi
echo "== 1. hero A/B (cfg2, B=8, 30 steps): v4 default / v4 noreg / v4 contig / v3"
for lib in default noreg contig; do
  if [ $lib = default ]; then L=$DEF; else L=$LIBDIR/libsrcv_b200_$lib.so; fi
  SRCV_B200_LIB=$L timeout 120 python bench.py --workload cfg2 --steps 30 --warmup 3 --no-cpu-baseline --no-also 2>$O/r02e_hero_$lib.err | tail -n 1 > $O/r02e_hero_$lib.json
  python -c "import json; d=json.load(open('$O/r02e_hero_$lib.json')); print('$lib', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['sweep_us_per_launch'],1), round(d['e2e']['value'],1), d['clocks'])" || tail -n 2 $O/r02e_hero_$lib.err
done
export SRCV_B200_LIB=$DEF
echo "== 2. parity on v3: tcgen05 tests + hero / golden / training-contract parity tests"
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -m gpu -x --timeout 200 -k "tc or golden or hero or mlp or autocast or strided or per_frame or shard" > $O/r02e_parity.log 2>&1; echo "rc=$?"; tail -n 4 $O/r02e_parity.log
timeout 300 python -m pytest tests/test_gpu_mvs.py tests/test_gpu_tsdf.py -q -m gpu -x --timeout 200 > $O/r02e_mvs_tsdf.log 2>&1; echo "mvs+tsdf rc=$?"; tail -n 3 $O/r02e_mvs_tsdf.log
echo "== 3. ncu: hero kernel v3, full sections + source (B=4)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 1 -c 1 \
    -o $O/prof_r02e_hero python scripts/run_once.py cfg2 4 2 > $O/r02e_ncu_hero.log 2>&1; echo "ncu rc=$?"
ls -la $O | tail -n 6
