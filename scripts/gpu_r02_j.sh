#!/bin/bash
# Round-2 GPU call J: hero timeline (clock64 stamps of CTA 0), TSDF bench after the cull rewrite.
set -u
O=gpurun_out
mkdir -p $O
SRCV_B200_LIB=$PWD/simplerecon_b200/lib/libsrcv_b200_tl.so timeout 120 python scripts/hero_timeline.py > $O/r02j_hero_timeline.json 2>$O/r02j_tl.err; echo "timeline rc=$?"
python -c "import json; d=json.load(open('$O/r02j_hero_timeline.json')); print(json.dumps(d['summary_clk'], indent=1))" || tail -n 5 $O/r02j_tl.err
timeout 200 python -m pytest tests/test_gpu_tsdf.py tests/test_gpu_parity.py -q -m gpu -x --timeout 200 -k "tsdf or errors_on_device or integrate" > $O/r02j_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 $O/r02j_tests.log
for cfg in "0.01 1" "0.01 4" "0.04 8"; do
  set -- $cfg
  timeout 150 python scripts/bench_tsdf.py --voxel $1 --frames $2 --steps 20 2>$O/r02j_tsdf.err | tail -n 1 | tee -a $O/r02j_tsdf.jsonl | head -c 420; echo
done
