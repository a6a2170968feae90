#!/bin/bash
# Round-2 GPU call V: last validation of the final commit — smoke, full GPU suite, the default bench
# line three times (run-to-run spread on one box), BASELINE configs[3] (D = 96, 4 frames per GPU),
# launch list of the dot workload.
set -u
O=gpurun_out
mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/r02v_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02v_smoke.log
[ $rc -ne 0 ] && exit 1
timeout 900 python -m pytest tests -q -m gpu --timeout 400 > $O/r02v_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -n 3 $O/r02v_gpu_suite.log
for i in 1 2 3; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>$O/r02v_bench$i.err | tail -n 1 > $O/r02v_bench$i.json
  python -c "
import json; d=json.load(open('$O/r02v_bench$i.json')); a=list(d['also'].values())[0]
print('run $i hero', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), '| dot', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), d['clocks'])"
done
timeout 300 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02v_cfg3.err | tail -n 1 > $O/r02v_bench_cfg3.json
python -c "
import json; d=json.load(open('$O/r02v_bench_cfg3.json')); print('cfg3', d['config']['workload'], round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), d['roofline']['frac'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_r02_cfg1.csv \
    python bench.py --workload cfg1 --steps 4 --warmup 3 --no-cpu-baseline --no-also > $O/r02v_launches_cfg1.log 2>&1; echo "launch list rc=$?"
