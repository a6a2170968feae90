#!/bin/bash
# Round-2 GPU call K: hero kernel with layer 1 in two column halves / layer 2 in two K halves
# (epilogue overlaps the tensor pipe): smoke, parity, bench line, timeline.
set -u
O=gpurun_out
mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/r02k_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02k_smoke.log
[ $rc -ne 0 ] && exit 1
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 300 -k "mlp or hero or golden" > $O/r02k_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 $O/r02k_parity.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>$O/r02k_bench.err | tail -n 1 > $O/r02k_bench.json
python -c "
import json; d=json.load(open('$O/r02k_bench.json'))
print('hero', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], d['clocks'])" || tail -n 5 $O/r02k_bench.err
SRCV_B200_LIB=$PWD/simplerecon_b200/lib/libsrcv_b200_tl.so timeout 120 python scripts/hero_timeline.py $O/r02k_hero_timeline.json > /dev/null 2>$O/r02k_tl.err; echo "timeline rc=$?"
python -c "import json; d=json.load(open('$O/r02k_hero_timeline.json')); print(json.dumps(d['summary_clk'], indent=1)); print(d['tiles_8_to_39'][0])" || tail -n 5 $O/r02k_tl.err
