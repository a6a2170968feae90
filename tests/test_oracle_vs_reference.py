"""CPU, build container only: the oracle against the LIVE reference classes on
randomised shapes (skipped where /root/reference is absent, e.g. the GPU box)."""
import pytest
import torch

from oracle import costvolume_oracle as O
from oracle.ref_import import load_reference, reference_available
from simplerecon_b200.synthetic import make_tuple, mlp_state
from tests.parity import assert_cost_close

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


@pytest.mark.parametrize("B,K,H,W,D,seed", [(2, 3, 20, 28, 6, 1), (1, 7, 15, 21, 5, 2), (3, 1, 8, 8, 3, 3)])
def test_dot_matches_live_reference(B, K, H, W, D, seed):
    R = load_reference()
    t = make_tuple(B, K, H, W, seed=seed)
    ref = R.CostVolumeManager(H, W, num_depth_bins=D)
    with torch.no_grad():
        rc, rl, rp, rm = ref(**t)
    oc, ol, op, om = O.forward_dot(**t, num_depth_bins=D, sampler="aten")
    assert torch.equal(op.contiguous(), rp.contiguous())
    assert (oc - rc).abs().max().item() <= 2e-6 * float(rc.abs().max())
    assert torch.equal(ol, rl) and rm is None and om is None
    oc2, *_ = O.forward_dot(**t, num_depth_bins=D, sampler="explicit")
    assert_cost_close("dot", oc2, rc)


@pytest.mark.parametrize("B,K,H,W,D,seed", [(2, 3, 12, 16, 4, 4), (1, 7, 10, 14, 3, 5)])
def test_mlp_matches_live_reference_slow_and_fast(B, K, H, W, D, seed):
    R = load_reference()
    t = make_tuple(B, K, H, W, seed=seed)
    sd = mlp_state(K, 16, seed=seed)
    ref = R.FeatureVolumeManager(H, W, num_depth_bins=D, mlp_channels=[0, 128, 128, 1],
                                 matching_dim_size=16, num_source_views=K)
    ref.load_state_dict({**ref.state_dict(), **sd})
    with torch.no_grad():
        rc, rl, rp, rm = ref(**t, return_mask=True)
        fc, fl, fp, fm = ref.to_fast()(**t, return_mask=True)
    w = O.mlp_weights_from_state_dict(sd)
    oc, ol, op, om = O.forward_mlp(**t, weights=w, num_depth_bins=D, return_mask=True, sampler="aten")
    for c in (rc, fc):
        assert (oc - c).abs().max().item() <= 5e-6 * float(c.abs().max()) + 1e-7
    assert torch.equal(om, rm) and torch.equal(om, fm)
    assert torch.equal(ol, rl)


def test_pose_distance_matches_live_reference():
    R = load_reference()
    t = make_tuple(3, 5, 4, 4, seed=9)
    comb, r, tm = O.pose_distance(t["src_poses"])
    rc, rr, rt = R.pose_distance(t["src_poses"].reshape(-1, 4, 4))
    assert torch.allclose(comb.reshape(-1), rc) and torch.allclose(r.reshape(-1), rr)
    assert torch.allclose(tm.reshape(-1), rt)


def test_dot_gradients_match_live_reference_autograd():
    """The oracle's autograd (which the GPU backward kernel is tested against) equals autograd
    through the reference's own CostVolumeManager for the two feature inputs."""
    R = load_reference()
    B, K, H, W, D = 2, 3, 10, 14, 4
    t = make_tuple(B, K, H, W, seed=17)
    g = torch.randn(B, D, H, W, generator=torch.Generator().manual_seed(1))
    grads = []
    for impl in ("ref", "oracle"):
        tt = dict(t)
        tt["cur_feats"] = t["cur_feats"].clone().requires_grad_(True)
        tt["src_feats"] = t["src_feats"].clone().requires_grad_(True)
        if impl == "ref":
            cost, *_ = R.CostVolumeManager(H, W, num_depth_bins=D)(**tt)
        else:
            cost, *_ = O.forward_dot(**tt, num_depth_bins=D, sampler="explicit")
        (cost * g).sum().backward()
        grads.append((tt["cur_feats"].grad, tt["src_feats"].grad))
    for a, b in zip(*grads):
        assert (a - b).abs().max().item() <= 2e-5 * float(a.abs().max()) + 1e-6


def test_per_frame_depth_range_matches_live_reference():
    """(B,1,1,1) min/max depth: the reference's generate_depth_planes broadcasts one range per
    frame (modules/cost_volume.py:124-127); the oracle restates that."""
    R = load_reference()
    B, K, H, W, D = 3, 2, 10, 12, 5
    t = make_tuple(B, K, H, W, seed=31)
    t["min_depth"] = torch.tensor([0.25, 0.5, 0.3]).view(B, 1, 1, 1)
    t["max_depth"] = torch.tensor([5.0, 8.0, 2.0]).view(B, 1, 1, 1)
    ref = R.CostVolumeManager(H, W, num_depth_bins=D)
    with torch.no_grad():
        rc, rl, rp, _ = ref(**t)
    oc, ol, op, _ = O.forward_dot(**t, num_depth_bins=D, sampler="aten")
    assert torch.equal(op.contiguous(), rp.contiguous())
    assert (oc - rc).abs().max().item() <= 2e-6 * float(rc.abs().max()) and torch.equal(ol, rl)
