#!/bin/bash
# Round-2 GPU call W: hero kernel "lean" (tile decode through a multiply-high division, the redundant
# second normalisation of the cosine similarity dropped) against the previous kernel
# (lib/libsrcv_b200_prev.so = HEAD's sources); parity of the new default.
set -u
O=gpurun_out
L=$PWD/simplerecon_b200/lib
mkdir -p $O
timeout 180 python __graft_entry__.py smoke > $O/r02w_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02w_smoke.log
[ $rc -ne 0 ] && exit 1
for v in _prev "" _prev "" _prev ""; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02w_bench$v.err | tail -n 1 > $O/r02w_bench$v.json
  python -c "
import json; d=json.load(open('$O/r02w_bench$v.json'))
print('hero$v', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1))" || tail -n 5 $O/r02w_bench$v.err
done
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_producer.py tests/test_gpu_depth_parity.py -q -m gpu -x --timeout 300 -k "mlp or hero or golden or tc or producer or depth" > $O/r02w_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 $O/r02w_parity.log
