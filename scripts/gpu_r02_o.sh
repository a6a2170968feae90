#!/bin/bash
# Round-2 GPU call O: hero kernel with packed fp32 issue (FFMA2 / FMUL2 in the builders' blend and both
# epilogues) and the whole layer-1 bias (b1 + the frame's pose measures) folded into the MMA, against the
# previous kernel (lib/libsrcv_b200_base.so = HEAD's sources) on the same box; then the full GPU suite.
set -u
O=gpurun_out
L=$PWD/simplerecon_b200/lib
mkdir -p $O
timeout 180 python __graft_entry__.py smoke > $O/r02o_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02o_smoke.log
[ $rc -ne 0 ] && exit 1
for v in _base "" _base ""; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02o_bench$v.err | tail -n 1 > $O/r02o_bench$v.json
  python -c "
import json; d=json.load(open('$O/r02o_bench$v.json'))
print('hero$v', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], d['clocks'])" || tail -n 5 $O/r02o_bench$v.err
done
SRCV_B200_LIB=$L/libsrcv_b200_tl.so timeout 120 python scripts/hero_timeline.py $O/r02o_hero_timeline.json > /dev/null 2>$O/r02o_tl.err; echo "timeline rc=$?"
python -c "import json; d=json.load(open('$O/r02o_hero_timeline.json')); print(json.dumps(d['summary_clk'])); print(d['tiles_8_to_39'][0])" || tail -n 5 $O/r02o_tl.err
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > $O/r02o_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -n 4 $O/r02o_gpu_suite.log
