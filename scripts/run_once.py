"""Runs a few forward passes of one workload (for ncu / compute-sanitizer).

    python scripts/run_once.py cfg1 [batch] [iters] [variant]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import simplerecon_b200 as S  # noqa: E402
from simplerecon_b200 import _native  # noqa: E402
from simplerecon_b200.synthetic import CONFIGS, make_workload_tuple, mlp_state, to_device  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
import dataclasses  # noqa: E402
import os  # noqa: E402

w = next(c for c in CONFIGS if c.name.startswith(name))
if os.environ.get("SRCV_SMALL"):          # sanitizer runs: same kernels, 37x53 map, 5 planes
    w = dataclasses.replace(w, height=37, width=53, planes=5)
batch = int(sys.argv[2]) if len(sys.argv) > 2 else w.batch
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
variant = sys.argv[4] if len(sys.argv) > 4 else "auto"
_native.set_variant({"auto": 0, "generic": 1, "fast": 2}[variant])
t = to_device(make_workload_tuple(w, batch=batch), "cuda")
if w.kind == "mlp":
    m = S.FeatureVolumeManager(w.height, w.width, w.planes, [0, 128, 128, 1], w.channels, w.views)
    m.load_state_dict({**m.state_dict(), **mlp_state(w.views, w.channels)})
    kw = dict(return_mask=True)
else:
    m = S.CostVolumeManager(w.height, w.width, w.planes)
    kw = {}
m = m.cuda().eval()
with torch.inference_mode():
    for _ in range(iters):
        out = m(**t, **kw)
torch.cuda.synchronize()
print(name, batch, _native.last_variant(), float(out[0].abs().max()))
