#!/bin/bash
# Round-2 GPU call B: hero kernel v2 — smoke, full GPU suite, A/B of the experiment builds, ncu, bench line.
#   python -m simplerecon_b200.build --out simplerecon_b200/lib/libsrcv_b200_noreg.so --extra=-DSRCV_TC_NO_SETMAXNREG
#   python -m simplerecon_b200.build --out simplerecon_b200/lib/libsrcv_b200_contig.so --extra=-DSRCV_TC_CONTIG
#   gpurun --timeout 1800 -- 'bash scripts/gpu_r02_b.sh'
set -u
O=gpurun_out
mkdir -p $O
echo "== 0. smoke (default build)"
timeout 180 python __graft_entry__.py smoke > $O/r02b_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 2 $O/r02b_smoke.log
if [ $rc -ne 0 ]; then
  echo "default build failed the smoke: trying the no-setmaxnreg build"
  SRCV_B200_LIB=simplerecon_b200/lib/libsrcv_b200_noreg.so timeout 180 python __graft_entry__.py smoke > $O/r02b_smoke_noreg.log 2>&1; echo "noreg smoke rc=$?"; tail -n 2 $O/r02b_smoke_noreg.log
fi
echo "== 1. full GPU suite"
timeout 900 python -m pytest tests -q -m gpu -x --timeout 400 > $O/r02b_gpu_suite.log 2>&1; echo "rc=$?"; tail -n 5 $O/r02b_gpu_suite.log
echo "== 2. hero A/B: default / noreg / contig  (cfg2, B=8, 30 steps)"
for lib in default noreg contig; do
  if [ $lib = default ]; then unset SRCV_B200_LIB; else export SRCV_B200_LIB=simplerecon_b200/lib/libsrcv_b200_$lib.so; fi
  timeout 300 python bench.py --workload cfg2 --steps 30 --warmup 3 --no-cpu-baseline --no-also 2>$O/r02b_hero_$lib.err | tail -n 1 > $O/r02b_hero_$lib.json
  python -c "import json; d=json.load(open('$O/r02b_hero_$lib.json')); print('$lib', round(d['value'],1), round(d['ms_per_step'],4), d['roofline']['sweep_us_per_launch'], d['e2e']['value'], d['clocks'])" || tail -n 3 $O/r02b_hero_$lib.err
done
unset SRCV_B200_LIB
echo "== 3. ncu: hero kernel v2, full sections + source (B=4)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 1 -c 1 \
    -o $O/prof_r02b_hero python scripts/run_once.py cfg2 4 2 > $O/r02b_ncu_hero.log 2>&1; echo "ncu rc=$?"
echo "== 4. the default bench line (hero + dot under 'also' + cpu baseline)"
timeout 900 python bench.py --steps 50 --warmup 5 2>$O/r02b_bench_default.err | tail -n 1 > $O/r02b_bench_default.json
python -c "
import json; d=json.load(open('$O/r02b_bench_default.json'))
print('hero', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['windows_ms_per_step'], d['roofline']['frac'], d['clocks'])
a=list(d['also'].values())[0]; print('dot', a['value'], a['ms_per_step'], 'e2e', a['e2e']['value'], a['roofline']['frac'], a['roofline']['binding'])
print('cpu', d['cpu_baseline'])" || tail -n 5 $O/r02b_bench_default.err
echo "== 5. reference arm, same config"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>$O/r02b_ref.err | tail -n 1 > $O/r02b_bench_reference.json; head -c 900 $O/r02b_bench_reference.json; echo
ls -la $O | tail -n 12
