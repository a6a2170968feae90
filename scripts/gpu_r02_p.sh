#!/bin/bash
# Round-2 GPU call P: hero builders re-balanced 2.5 / 2.5 / 2 + tail (view 2 built as two half-view units
# by slots 0 and 1) against the previous kernel (lib/libsrcv_b200_prev.so) and with contiguous tile
# ranges per CTA (-DSRCV_TC_CONTIG); parity of the new default.
set -u
O=gpurun_out
L=$PWD/simplerecon_b200/lib
mkdir -p $O
timeout 180 python __graft_entry__.py smoke > $O/r02p_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02p_smoke.log
[ $rc -ne 0 ] && exit 1
for v in _prev "" _contig _prev "" _contig; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02p_bench$v.err | tail -n 1 > $O/r02p_bench$v.json
  python -c "
import json; d=json.load(open('$O/r02p_bench$v.json'))
print('hero$v', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], d['clocks'])" || tail -n 5 $O/r02p_bench$v.err
done
SRCV_B200_LIB=$L/libsrcv_b200_tl.so timeout 120 python scripts/hero_timeline.py $O/r02p_hero_timeline.json > /dev/null 2>$O/r02p_tl.err; echo "timeline rc=$?"
python -c "import json; d=json.load(open('$O/r02p_hero_timeline.json')); print(json.dumps(d['summary_clk'])); print(d['tiles_8_to_39'][0])" || tail -n 5 $O/r02p_tl.err
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_producer.py -q -m gpu -x --timeout 300 -k "mlp or hero or golden or tc or producer" > $O/r02p_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 $O/r02p_parity.log
