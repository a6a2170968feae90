#!/bin/bash
# Round-2 GPU call N: hero kernel with half-view builder units (software-pipelined gathers,
# 5 / 5 / 4+tail unit split) against the shipped whole-view builders; gather ablation.
#   python -m simplerecon_b200.build --out simplerecon_b200/lib/libsrcv_b200_hu.so --extra=-DSRCV_TC_HALF_UNITS  (+ _tl, _abl)
set -u
O=gpurun_out
L=$PWD/simplerecon_b200/lib
mkdir -p $O
timeout 180 python __graft_entry__.py smoke > $O/r02n_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02n_smoke.log
[ $rc -ne 0 ] && exit 1
SRCV_B200_LIB=$L/libsrcv_b200_hu.so timeout 180 python __graft_entry__.py smoke > $O/r02n_smoke_hu.log 2>&1; rc=$?; echo "smoke(hu) rc=$rc"; tail -n 1 $O/r02n_smoke_hu.log
for v in "" _hu _hu_abl; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02n_bench$v.err | tail -n 1 > $O/r02n_bench$v.json
  python -c "
import json; d=json.load(open('$O/r02n_bench$v.json'))
print('hero$v', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], d['clocks'])" || tail -n 5 $O/r02n_bench$v.err
done
for v in "" _hu; do
  SRCV_B200_LIB=$L/libsrcv_b200${v}_tl.so timeout 120 python scripts/hero_timeline.py $O/r02n_hero_timeline$v.json > /dev/null 2>$O/r02n_tl$v.err; echo "timeline rc=$?"
  python -c "import json; d=json.load(open('$O/r02n_hero_timeline$v.json')); print(json.dumps(d['summary_clk'])); print(d['tiles_8_to_39'][0])" || tail -n 5 $O/r02n_tl$v.err
done
timeout 300 python -m pytest tests/test_gpu_mvloss.py -q -m gpu -x --timeout 200 > $O/r02n_mvloss.log 2>&1; echo "mvloss rc=$?"; tail -n 3 $O/r02n_mvloss.log
SRCV_B200_LIB=$L/libsrcv_b200_hu.so timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q -m gpu -x --timeout 300 -k "mlp or hero or golden or tc" > $O/r02n_parity_hu.log 2>&1; echo "parity(hu) rc=$?"; tail -n 3 $O/r02n_parity_hu.log
