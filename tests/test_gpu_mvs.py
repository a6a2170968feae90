"""GPU: the multi-view consistency kernel through the C ABI against the CPU oracle (fp32, the
reference's operation order; discontinuous decisions compared by mismatch count)."""
import pytest
import torch

from simplerecon_b200 import _native, point_cloud_fusion as pcf
from simplerecon_b200.synthetic import make_mvs_scene
from tests.test_emu_mvs import check_against_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n,hw", [(11, 8, (60, 80)), (12, 40, (48, 64))])
def test_consistency_matches_oracle(cuda_device, seed, n, hw):
    sc = make_mvs_scene(seed=seed, frames=n, height=hw[0], width=hw[1])
    scan = pcf._Scan(sc["depths"], sc["cam_T_world"], sc["K"], torch.device("cuda"))
    for ref_idx in (0, n // 3, n - 1):
        check_against_oracle(scan, sc, ref_idx)
    assert _native.last_variant() == "mvs_consistency_f32"


def test_process_scene_full_size(cuda_device):
    """24 frames at the fuser's 480x640 (pc_fusion.py:121-125): per-frame results of process_scene equal
    process_depth on the same frame; validity is deterministic."""
    sc = make_mvs_scene(seed=21, frames=24, height=480, width=640)
    d, im, P, K = (sc[k].cuda() for k in ("depths", "images", "cam_T_world", "K"))
    fp, fr, av = pcf.process_scene(d, im, P, K, 0.1, 3)
    assert av.shape == (24, 480, 640) and fp.shape == (int(av.sum()), 3) and fr.shape == fp.shape
    assert 0.2 < av.mean() < 0.99
    src = torch.arange(24) != 5
    pts5, rgb5, valid5 = pcf.process_depth(d[5], im[5], d[src], im[src], P[5], P[src], K[5], K[src], 0.1, 3)
    assert (valid5 == av[5]).all()
    lo = int(av[:5].sum())
    assert abs(fp[lo:lo + pts5.shape[0]] - pts5).max() < 1e-6
