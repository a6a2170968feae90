#!/bin/bash
# Round-2 GPU call U (gpurun --gpus 2): the driver's multi-GPU launch of both bench arms on the hero
# default — one rank per GPU over NCCL, frames sharded, no data-path collective.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 5 2>$O/r02u_bench_n2.err | tail -n 1 > $O/r02u_bench_n2.json; echo "rc=$?"
python -c "
import json; d=json.load(open('$O/r02u_bench_n2.json'))
print('N=2 hero', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), d['config']['global_batch'], d['config']['parallelism'], d['clocks'])
a=list(d.get('also',{}).values()); print('dot', round(a[0]['value'],1) if a else None)" || tail -n 8 $O/r02u_bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>$O/r02u_ref_n2.err | tail -n 1 > $O/r02u_bench_reference_n2.json; echo "rc=$?"; head -c 600 $O/r02u_bench_reference_n2.json; echo
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-also 2>$O/r02u_bench_n1.err | tail -n 1 > $O/r02u_bench_n1.json
python -c "
import json; d=json.load(open('$O/r02u_bench_n1.json')); print('N=1 hero', round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1))"
