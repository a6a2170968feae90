#!/bin/bash
# Round-2 GPU call R: packed fp32 FMAs (FFMA2) in the dot sweep's tap reduction (default: capped at 64
# registers; _nomin: 68 registers / 28 warps) and in the MLP backward's GEMM loops, against the previous
# library (lib/libsrcv_b200_prev.so); parity of the dot and backward paths.
set -u
O=gpurun_out
L=$PWD/simplerecon_b200/lib
mkdir -p $O
timeout 180 python __graft_entry__.py smoke > $O/r02r_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -n 1 $O/r02r_smoke.log
[ $rc -ne 0 ] && exit 1
for v in _prev "" _nomin _prev "" _nomin; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 300 python bench.py --workload cfg1 --steps 60 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02r_bench_dot$v.err | tail -n 1 > $O/r02r_bench_dot$v.json
  python -c "
import json; d=json.load(open('$O/r02r_bench_dot$v.json'))
print('dot$v', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'sweep_us', d['roofline'].get('kernel_us'), 'frac', d['roofline']['frac'])" || tail -n 5 $O/r02r_bench_dot$v.err
done
for v in _prev ""; do
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 300 python scripts/bench_backward.py --workload cfg2 --steps 5 --warmup 2 > $O/r02r_bwd_mlp$v.json 2>$O/r02r_bwd$v.err; echo "bwd mlp$v rc=$?"; tail -n 1 $O/r02r_bwd_mlp$v.json
  SRCV_B200_LIB=$L/libsrcv_b200$v.so timeout 300 python scripts/bench_backward.py --workload cfg1 --steps 10 --warmup 3 > $O/r02r_bwd_dot$v.json 2>>$O/r02r_bwd$v.err; echo "bwd dot$v rc=$?"; tail -n 1 $O/r02r_bwd_dot$v.json
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zzz_gpu_mlp_backward.py tests/test_zz_gpu_torch_ops.py -q -m gpu -x --timeout 600 > $O/r02r_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 $O/r02r_parity.log
