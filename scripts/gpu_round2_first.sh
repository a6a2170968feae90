#!/bin/bash
# First GPU call of the next round (run through gpurun, one GPU): everything that was written
# after round 1's GPU budget was spent, in one box acquisition.  Every step has its own timeout;
# logs land in gpurun_out/r02_first_*.  Build the experiment library BEFORE calling gpurun:
#   python -m simplerecon_b200.build --out simplerecon_b200/lib/libsrcv_b200_uw.so --extra=-DSRCV_TC_UNIFORM_WARP --extra=-DSRCV_TC_EARLY_FLAGS --extra=-DSRCV_TC_TILE32 --extra=-DSRCV_TC_CLAMPED_TAPS
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
set -u
O=gpurun_out
mkdir -p $O
UW=simplerecon_b200/lib/libsrcv_b200_uw.so
echo "== 1. new GPU tests (torch operators, metadata-MLP backward)"
timeout 240 python -m pytest tests/test_zz_gpu_torch_ops.py tests/test_zzz_gpu_mlp_backward.py -q -m gpu \
    > $O/r02_first_newtests.log 2>&1; echo "rc=$?"; tail -n 4 $O/r02_first_newtests.log
echo "== 2. full GPU suite"
timeout 400 python -m pytest tests -q -m gpu -x > $O/r02_first_gpu_suite.log 2>&1; echo "rc=$?"; tail -n 3 $O/r02_first_gpu_suite.log
echo "== 3. backward kernels: time per training step"
for w in cfg1 cfg2; do
  timeout 240 python scripts/bench_backward.py --workload $w --steps 5 --warmup 2 2>/dev/null | tail -n 1 | tee $O/r02_first_backward_$w.json
done
echo "== 4. hero kernel: default vs -DSRCV_TC_UNIFORM_WARP (parity first, then the bench)"
if [ -f $UW ]; then
  SRCV_B200_LIB=$UW timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -m gpu -k "tc or hero or mlp or golden" \
      > $O/r02_first_uw_parity.log 2>&1; echo "uw parity rc=$?"; tail -n 2 $O/r02_first_uw_parity.log
  for lib in default uw; do
    if [ $lib = uw ]; then export SRCV_B200_LIB=$UW; else unset SRCV_B200_LIB; fi
    timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 > $O/r02_first_hero_$lib.json
    python -c "import json,sys; d=json.load(open('$O/r02_first_hero_$lib.json')); print('$lib', d['value'], d['ms_per_step'], d['clocks'])"
  done
  unset SRCV_B200_LIB
else
  echo "no $UW — build it first (see the header of this script)"
fi
echo "== 5. ncu: the MLP backward kernel (one launch, full sections)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_backward_kernel -c 1 \
    -o $O/prof_r02_mlp_bwd python scripts/bench_backward.py --workload cfg2 --batch 1 --steps 1 --warmup 0 \
    > $O/r02_first_ncu_bwd.log 2>&1; echo "ncu rc=$?"
echo "== 6. measured L1 gather ceiling for the dot sweep's load shape (scripts/probes)"
if [ -x scripts/probes/l1_gather_probe ]; then
  timeout 60 ./scripts/probes/l1_gather_probe | tee $O/r02_first_l1_probe.jsonl
else
  echo "build it first: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/l1_gather_probe scripts/probes/l1_gather_probe.cu"
fi
ls -la $O | tail -n 12
echo "== 7. ncu: hero kernel, variant build (all four switches), full sections + source"
if [ -f $UW ]; then
  SRCV_B200_LIB=$UW timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 1 -c 1 \
      -o $O/prof_r02_hero_uw python scripts/run_once.py cfg2 4 2 > $O/r02_first_ncu_hero_uw.log 2>&1; echo "ncu rc=$?"
fi
echo "== 8. dot bench (cfg1) + stress cfgS on the default build"
timeout 200 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline 2>$O/r02_first_cfg1.err | tail -n 1 > $O/r02_first_cfg1.json
timeout 300 python bench.py --workload stress_dot --steps 5 --warmup 3 --no-cpu-baseline 2>$O/r02_first_stress_dot.err | tail -n 1 > $O/r02_first_stress_dot.json; cat $O/r02_first_stress_dot.json | head -c 600
ls -la $O | tail -n 14
