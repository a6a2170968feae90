#!/bin/bash
# Round-2 GPU call C (after the setmaxnreg budget fix): smoke first — everything else only if it passes.
set -u
O=gpurun_out
mkdir -p $O
echo "== 0. smoke (default build, setmaxnreg 112/24)"
timeout 120 python __graft_entry__.py smoke > $O/r02c_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02c_smoke.log
if [ $rc -ne 0 ]; then
  echo "default build failed the smoke: falling back to the no-setmaxnreg build for the rest of this call"
  export SRCV_B200_LIB=$PWD/simplerecon_b200/lib/libsrcv_b200_noreg.so
  timeout 120 python __graft_entry__.py smoke > $O/r02c_smoke_noreg.log 2>&1; rc2=$?; echo "noreg smoke rc=$rc2"; tail -n 1 $O/r02c_smoke_noreg.log
  if [ $rc2 -ne 0 ]; then echo "both builds fail: stopping"; exit 1; fi
fi
echo "== 1. hero A/B (cfg2, B=8, 30 steps): default / noreg / contig"
for lib in default noreg contig; do
  if [ $lib != default ]; then L=$PWD/simplerecon_b200/lib/libsrcv_b200_$lib.so; else L=${SRCV_B200_LIB:-}; fi
  SRCV_B200_LIB=$L timeout 120 python bench.py --workload cfg2 --steps 30 --warmup 3 --no-cpu-baseline --no-also 2>$O/r02c_hero_$lib.err | tail -n 1 > $O/r02c_hero_$lib.json
  python -c "import json; d=json.load(open('$O/r02c_hero_$lib.json')); print('$lib', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['sweep_us_per_launch'],1), round(d['e2e']['value'],1), d['clocks'])" || tail -n 3 $O/r02c_hero_$lib.err
done
echo "== 2. ncu: hero kernel v2, full sections + source (B=4)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 1 -c 1 \
    -o $O/prof_r02c_hero python scripts/run_once.py cfg2 4 2 > $O/r02c_ncu_hero.log 2>&1; echo "ncu rc=$?"
echo "== 3. hero + TSDF + producer + depth-parity dump tests"
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tsdf.py tests/test_gpu_depth_parity.py -q -m gpu -x --timeout 200 > $O/r02c_newtests.log 2>&1; echo "rc=$?"; tail -n 4 $O/r02c_newtests.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 200 -k "golden or hero or autocast or strided or per_frame" > $O/r02c_parity.log 2>&1; echo "rc=$?"; tail -n 4 $O/r02c_parity.log
echo "== 4. TSDF bench: 1 cm room volume (438 MB), 1 and 4 frames per call; 4 cm, 8 frames"
for cfg in "0.01 1" "0.01 4" "0.04 8"; do
  set -- $cfg
  timeout 150 python scripts/bench_tsdf.py --voxel $1 --frames $2 --steps 20 2>$O/r02c_tsdf.err | tail -n 1 | tee -a $O/r02c_tsdf.jsonl | head -c 600; echo
done
ls -la $O | tail -n 8
