#!/bin/bash
# Round-2 GPU call X: the final commit once more — smoke, full GPU suite, default bench line with the
# CPU baseline, reference arm (the two commands the driver runs at round end).
set -u
O=gpurun_out
mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/r02x_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02x_smoke.log
[ $rc -ne 0 ] && exit 1
timeout 900 python -m pytest tests -q -m gpu --timeout 400 > $O/r02x_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -n 3 $O/r02x_gpu_suite.log
timeout 600 python bench.py --steps 50 --warmup 5 2>$O/r02x_bench_default.err | tail -n 1 > $O/r02x_bench_default.json
python -c "
import json; d=json.load(open('$O/r02x_bench_default.json')); a=list(d['also'].values())[0]
print('hero', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], d['roofline']['issued_mma']['frac_of_sustained'], '| dot', round(a['value'],1), 'e2e', round(a['e2e']['value'],1), d['clocks'], 'launches', d['gpu_launches'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>$O/r02x_ref.err | tail -n 1 > $O/r02x_bench_reference.json; head -c 300 $O/r02x_bench_reference.json; echo
