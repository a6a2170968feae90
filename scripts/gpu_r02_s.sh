#!/bin/bash
# Round-2 GPU call S: hero kernel with SIXTEEN builder warps (24 warps per CTA, two K blocks per builder
# thread; register splits A = 80 / 120 / 24 and B = 88 / 104 / 24) against the shipped 12-builder kernel.
# Every run of a new setmaxnreg budget sits under a short timeout (a budget error spins forever).
set -u
O=gpurun_out
L=$PWD/simplerecon_b200/lib
mkdir -p $O
timeout 180 python __graft_entry__.py smoke > $O/r02s_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02s_smoke.log
[ $rc -ne 0 ] && exit 1
for v in _w24b _w24a; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 100 python __graft_entry__.py smoke > $O/r02s_smoke$v.log 2>&1; rc=$?; echo "smoke($v) rc=$rc"; tail -n 1 $O/r02s_smoke$v.log
  [ $rc -ne 0 ] && { echo "variant $v failed its smoke: skipping the rest"; exit 1; }
done
for v in "" _w24b _w24a "" _w24b _w24a; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02s_bench$v.err | tail -n 1 > $O/r02s_bench$v.json
  python -c "
import json; d=json.load(open('$O/r02s_bench$v.json'))
print('hero$v', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], d['clocks'])" || tail -n 5 $O/r02s_bench$v.err
done
SRCV_B200_LIB=$L/libsrcv_b200_w24b_tl.so timeout 120 python scripts/hero_timeline.py $O/r02s_hero_timeline_w24b.json > /dev/null 2>$O/r02s_tl.err; echo "timeline rc=$?"
python -c "import json; d=json.load(open('$O/r02s_hero_timeline_w24b.json')); print(json.dumps(d['summary_clk'])); print(d['tiles_8_to_39'][0])" || tail -n 5 $O/r02s_tl.err
SRCV_B200_LIB=$L/libsrcv_b200_w24b.so timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_producer.py -q -m gpu -x --timeout 300 -k "mlp or hero or golden or tc or producer" > $O/r02s_parity.log 2>&1; echo "parity(w24b) rc=$?"; tail -n 3 $O/r02s_parity.log
