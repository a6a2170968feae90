#!/bin/bash
# Round-2 GPU call M: hero kernel with elect.sync-guarded MMA issue (bare UTCHMMA, no per-MMA
# ELECT/branch wrapper), without and with the layer-1 column-half / layer-2 K-half split.
set -u
O=gpurun_out
L=$PWD/simplerecon_b200/lib
mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/r02m_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02m_smoke.log
[ $rc -ne 0 ] && exit 1
for v in "" _split; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02m_bench$v.err | tail -n 1 > $O/r02m_bench$v.json
  python -c "
import json; d=json.load(open('$O/r02m_bench$v.json'))
print('hero$v', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], d['clocks'])" || tail -n 5 $O/r02m_bench$v.err
  SRCV_B200_LIB=$L/libsrcv_b200_tl$v.so timeout 120 python scripts/hero_timeline.py $O/r02m_hero_timeline$v.json > /dev/null 2>$O/r02m_tl$v.err; echo "timeline rc=$?"
  python -c "import json; d=json.load(open('$O/r02m_hero_timeline$v.json')); print(json.dumps(d['summary_clk'])); print(d['tiles_8_to_39'][0])" || tail -n 5 $O/r02m_tl$v.err
done
SRCV_B200_LIB=$L/libsrcv_b200_split.so timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 300 -k "mlp or hero or golden" > $O/r02m_parity_split.log 2>&1; echo "parity(split) rc=$?"; tail -n 3 $O/r02m_parity_split.log
timeout 60 scripts/_bin/mma_probe > $O/r02m_mma_probe.jsonl 2>&1; echo "probe rc=$?"
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 300 -k "mlp or hero or golden" > $O/r02m_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 $O/r02m_parity.log
