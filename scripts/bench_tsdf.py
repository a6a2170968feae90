"""Timing + roofline of the TSDF integration kernel, one JSON line.

    python scripts/bench_tsdf.py [--voxel 0.01] [--frames 1] [--steps 20]

Workload: a 6 x 5 x 3 m room; at --voxel 0.01 the volume is 632 x 528 x 328 voxels = 438 MB of fp16
values + weights (> the 126 MB L2), `frames` depth maps of 240 x 320 per call.
Algorithmic bytes per launch: every voxel a frame of the batch updates is read and
written once (2 x 4 bytes) + the depth maps; voxels outside every frustum cost nothing.
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import simplerecon_b200 as S  # noqa: E402
from simplerecon_b200.synthetic import make_tsdf_case  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voxel", type=float, default=0.01)
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--hw", type=int, nargs=2, default=[240, 320])
a = ap.parse_args()
cases = [make_tsdf_case(seed=100 + i, frames=a.frames, voxel_size=a.voxel, height=a.hw[0], width=a.hw[1],
                        room=(6.0, 5.0, 3.0)) for i in range(4)]
vol = S.TSDF.from_bounds(cases[0]["bounds"], a.voxel)
fuser = S.TSDFFuser(vol, max_depth=3.0)
dev = [{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in c.items()} for c in cases]
dev = [{**d, "depth": d["depth"].half(), "cam_T_world": d["cam_T_world"].half(), "K": d["K"].half()} for d in dev]
# bytes: count the voxels each call touches on a fresh volume (a first pass, untimed)
touched = []
for d in dev:
    v0 = S.TSDF.from_bounds(cases[0]["bounds"], a.voxel)
    S.TSDFFuser(v0, max_depth=3.0).integrate_depth(d["depth"], d["cam_T_world"], d["K"])
    touched.append(int((v0.tsdf_weights > 0).sum()))
    del v0
for i in range(3):
    fuser.integrate_depth(dev[i % 4]["depth"], dev[i % 4]["cam_T_world"], dev[i % 4]["K"])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    d = dev[i % 4]
    fuser.integrate_depth(d["depth"], d["cam_T_world"], d["K"])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
nvox = vol.tsdf_values.numel()
alg = 8 * sum(touched) / len(touched) + a.frames * a.hw[0] * a.hw[1] * 2
dense = 8 * nvox
peaks = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text()) \
    if (Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").is_file() else {"hbm_gbs": 6650.0}
print(json.dumps({
    "kernel": "tsdf_integrate_f16", "volume": list(vol.tsdf_values.shape), "voxels": nvox,
    "volume_MB": 4 * nvox / 1e6, "frames_per_call": a.frames, "depth_hw": a.hw, "ms_per_call": ms,
    "frames_per_s": a.frames / (ms * 1e-3), "touched_voxels_per_call": sum(touched) / len(touched),
    "touched_frac": sum(touched) / len(touched) / nvox,
    "roofline": {"bound": "hbm", "algorithmic_bytes_per_launch": alg, "achieved": alg / (ms * 1e-3) / 1e9,
                 "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                 "dense_sweep_equiv_GBps": dense / (ms * 1e-3) / 1e9,
                 "note": "algorithmic = 8 B per voxel some frame updates + depth maps; dense_sweep_equiv = what a "
                         "read-modify-write of the WHOLE volume in the same time would move (the reference touches "
                         "every voxel ~40 times per frame)"},
}))
