"""GPU: the TSDF integration kernel through the C ABI (via the TSDF / TSDFFuser mirror of the
reference classes) against the CPU oracle — bit for bit: the arithmetic is fp16 with one rounding
per operation, restated from reference tools/tsdf.py:204-320 and pinned to the imported reference
in tests/test_tsdf_oracle_vs_reference.py."""
import pytest
import torch

import simplerecon_b200 as S
from oracle import tsdf_oracle as T
from simplerecon_b200 import _native
from simplerecon_b200.synthetic import make_tsdf_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,frames,voxel,hw,masked", [
    (11, 1, 0.06, (48, 64), False),
    (12, 4, 0.05, (96, 128), False),
    (13, 3, 0.07, (60, 80), True),
    (14, 20, 0.08, (48, 64), False),        # > 16 frames: two launches
])
def test_integrate_matches_oracle_bitwise(cuda_device, seed, frames, voxel, hw, masked):
    c = make_tsdf_case(seed=seed, frames=frames, voxel_size=voxel, height=hw[0], width=hw[1], masked=masked)
    vol = S.TSDF.from_bounds(c["bounds"], voxel)
    fuser = S.TSDFFuser(vol, max_depth=c["max_depth"])
    tv, tw, origin = T.new_volume(c["bounds"], voxel)
    n0 = _native.launch_count()
    for _ in range(2):
        fuser.integrate_depth(c["depth"].cuda(), c["cam_T_world"].cuda(), c["K"].cuda(),
                              None if c["mask"] is None else c["mask"].cuda())
        T.integrate(tv, tw, origin, voxel, c["depth"], c["cam_T_world"], c["K"], c["mask"],
                    min_depth=fuser.min_depth, max_depth=c["max_depth"])
        torch.cuda.synchronize()
        assert int((tw > 0).sum()) > 1000
        ow, ot = vol.tsdf_weights.cpu(), vol.tsdf_values.cpu()
        assert torch.equal(ow, tw), f"weights differ at {(ow != tw).sum().item()} voxels"
        assert torch.equal(ot, tv), f"tsdf differs at {(ot != tv).sum().item()} voxels"
    assert _native.launch_count() - n0 == 2 * 2 * (1 + (frames - 1) // 16)
    assert _native.last_variant() == "tsdf_integrate_f16"


def test_full_size_volume_properties(cuda_device):
    """A room at the reference's default 4 cm resolution, 8 frames of 240x320 predicted depth (the
    depth output of a 640x480 frame): (1) integrating the batch in one call equals integrating the
    frames one by one (the kernel applies frames in order); (2) voxels no frustum reaches keep the
    initial state; (3) weights stay in [0, 1], values in [-1, 1]; (4) deterministic."""
    c = make_tsdf_case(seed=21, frames=8, voxel_size=0.04, height=240, width=320, room=(6.0, 5.0, 3.0))
    d, E, K = c["depth"].cuda(), c["cam_T_world"].cuda(), c["K"].cuda()
    a = S.TSDF.from_bounds(c["bounds"], 0.04)
    b = S.TSDF.from_bounds(c["bounds"], 0.04)
    S.TSDFFuser(a, max_depth=3.0).integrate_depth(d, E, K)
    fb = S.TSDFFuser(b, max_depth=3.0)
    for i in range(8):
        fb.integrate_depth(d[i:i + 1], E[i:i + 1], K[i:i + 1])
    assert torch.equal(a.tsdf_values, b.tsdf_values) and torch.equal(a.tsdf_weights, b.tsdf_weights)
    touched = a.tsdf_weights > 0
    assert 0.02 < touched.float().mean().item() < 0.9
    assert torch.all(a.tsdf_values[~touched] == -1)
    assert a.tsdf_weights.max().item() <= 1.0 and a.tsdf_values.abs().max().item() <= 1.0
    a2 = S.TSDF.from_bounds(c["bounds"], 0.04)
    S.TSDFFuser(a2, max_depth=3.0).integrate_depth(d, E, K)
    assert torch.equal(a.tsdf_values, a2.tsdf_values) and torch.equal(a.tsdf_weights, a2.tsdf_weights)
