#!/bin/bash
# Round-2 GPU call T: validation of the shipped build — full GPU suite, default bench line (hero +
# dot under "also" + cpu baseline), reference arm, stress shape, ncu launch list, dot kernel capture,
# MVDepthLoss timing, sanitizers.
set -u
O=gpurun_out
mkdir -p $O
echo "== 0. smoke"
timeout 120 python __graft_entry__.py smoke > $O/r02t_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02t_smoke.log
[ $rc -ne 0 ] && exit 1
echo "== 1. full GPU suite"
timeout 900 python -m pytest tests -q -m gpu --timeout 400 > $O/r02t_gpu_suite.log 2>&1; echo "rc=$?"; tail -n 6 $O/r02t_gpu_suite.log
echo "== 2. default bench line, then the reference arm"
timeout 600 python bench.py --steps 50 --warmup 5 2>$O/r02t_bench_default.err | tail -n 1 > $O/r02t_bench_default.json
python -c "
import json; d=json.load(open('$O/r02t_bench_default.json'))
print('hero', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), [round(x,3) for x in d['e2e']['windows_ms_per_step']], 'frac', d['roofline']['frac'], d['roofline']['issued_mma'], d['clocks'], d['host_binding'])
a=list(d['also'].values())[0]; print('dot', round(a['value'],1), a['ms_per_step'], 'e2e', round(a['e2e']['value'],1), [round(x,3) for x in a['e2e']['windows_ms_per_step']], a['roofline']['frac'], a['roofline']['binding'])
print('cpu', d['cpu_baseline'])" || tail -n 5 $O/r02t_bench_default.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 2>$O/r02t_ref.err | tail -n 1 > $O/r02t_bench_reference.json; head -c 900 $O/r02t_bench_reference.json; echo
echo "== 2b. stress shapes (feature map 480 x 640)"
timeout 300 python bench.py --workload stress_dot --steps 10 --warmup 3 --no-cpu-baseline --no-also 2>$O/r02t_stress_dot.err | tail -n 1 > $O/r02t_bench_stress_dot.json; head -c 300 $O/r02t_bench_stress_dot.json; echo
timeout 300 python bench.py --workload stress_hero --steps 5 --warmup 3 --no-cpu-baseline --no-also 2>$O/r02t_stress_hero.err | tail -n 1 > $O/r02t_bench_stress_hero.json; head -c 300 $O/r02t_bench_stress_hero.json; echo
echo "== 2c. MVDepthLoss"
timeout 200 python scripts/bench_mvloss.py > $O/r02t_mvloss.json 2>$O/r02t_mvloss.err; tail -n 1 $O/r02t_mvloss.json
echo "== 3. ncu launch list of the bench command (every launch with its device time)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02t_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/r02t_launches.log 2>&1; echo "rc=$?"
echo "== 4. ncu --set full: dot sweep (B=4)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dot_fast -s 2 -c 1 \
    -o $O/prof_r02t_dot python scripts/run_once.py cfg1 4 3 > $O/r02t_ncu_dot.log 2>&1; echo "rc=$?"
echo "== 5. compute-sanitizer (small shapes): memcheck on both sweeps and the loss"
SRCV_SMALL=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/run_once.py cfg2 1 1 > $O/r02t_memcheck_hero.log 2>&1; echo "memcheck hero rc=$?"; tail -n 2 $O/r02t_memcheck_hero.log
SRCV_SMALL=1 timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/run_once.py cfg1 1 1 > $O/r02t_memcheck_dot.log 2>&1; echo "memcheck dot rc=$?"; tail -n 2 $O/r02t_memcheck_dot.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/bench_mvloss.py --batch 1 --views 3 --height 24 --width 32 --steps 2 > $O/r02t_memcheck_mvloss.log 2>&1; echo "memcheck mvloss rc=$?"; tail -n 2 $O/r02t_memcheck_mvloss.log
ls -la $O | tail -n 8
