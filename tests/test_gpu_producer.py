"""GPU: producer-side fusion through the C ABI — InstanceNorm2d fused with the
chunk-planar layout, the sweeps on caller-provided chunk-planar features, raw poses in the prep."""
import pytest
import torch

import simplerecon_b200 as S
from oracle import costvolume_oracle as O
from simplerecon_b200 import _native
from simplerecon_b200.synthetic import make_tuple, mlp_state, to_device
from tests.parity import assert_cost_close, assert_mask_close

pytestmark = pytest.mark.gpu


def _back(c4):
    return c4.movedim(-1, -3).flatten(-4, -3)        # (..., C/4, H, W, 4) -> (..., C, H, W)


@pytest.mark.parametrize("kind", ["dot", "mlp"])
def test_chunk_planar_inputs_and_raw_poses(cuda_device, kind):
    B, K, C, H, W, D = 2, 7, 16, 30, 40, 8
    t = make_tuple(B, K, H, W, channels=C, seed=51)
    g = torch.Generator().manual_seed(52)
    x = (2.5 * torch.randn(B, 1 + K, C, H, W, generator=g) - 0.3).cuda()
    cur_c4, src_c4 = S.instance_norm_to_chunk_planar(x)
    assert _native.last_variant() == "instnorm_chunk_planar"
    ref = torch.nn.InstanceNorm2d(C)(x.reshape(B * (1 + K), C, H, W)).reshape(B, 1 + K, C, H, W)   # networks.py:201
    assert (_back(cur_c4) - ref[:, 0]).abs().max().item() < 3e-6
    assert (_back(src_c4) - ref[:, 1:]).abs().max().item() < 3e-6
    td = to_device(t, "cuda")
    nchw = dict(td, cur_feats=_back(cur_c4).contiguous(), src_feats=_back(src_c4).contiguous())
    c4 = dict(td, cur_feats=cur_c4, src_feats=src_c4)
    if kind == "dot":
        m = S.CostVolumeManager(H, W, num_depth_bins=D).cuda().eval()
    else:
        m = S.FeatureVolumeManager(H, W, num_depth_bins=D, mlp_channels=[0, 128, 128, 1], matching_dim_size=C,
                                   num_source_views=K)
        m.load_state_dict({**m.state_dict(), **mlp_state(K, C, seed=2)})
        m = m.cuda().eval()
    world_T_cur = torch.linalg.inv(make_tuple(B, 1, H, W, seed=53)["src_extrinsics"][:, 0].double())
    cur_T_world = torch.linalg.inv(world_T_cur)
    raw = {k: v.float().cuda() for k, v in dict(
        src_cam_T_world=t["src_extrinsics"].double() @ cur_T_world[:, None],
        src_world_T_cam=world_T_cur[:, None] @ t["src_poses"].double(),
        cur_cam_T_world=cur_T_world, cur_world_T_cam=world_T_cur).items()}
    with torch.inference_mode():
        n0 = _native.launch_count()
        a = m(**nchw, return_mask=True)
        n_nchw = _native.launch_count() - n0
        b = m(**c4, return_mask=True)
        r = m(**{**nchw, "src_extrinsics": None, "src_poses": None}, return_mask=True, raw_poses=raw)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])          # same values, no prep copy
    if a[3] is not None:
        assert torch.equal(a[3], b[3])
    assert n_nchw >= 2
    assert (r[0] - a[0]).abs().max().item() <= 2e-4 * float(a[0].abs().max())
    # and the chunk-planar path against the oracle on the normalised features
    tc = dict(t, cur_feats=nchw["cur_feats"].cpu(), src_feats=nchw["src_feats"].cpu())
    if kind == "dot":
        oc, *_ = O.forward_dot(**tc, num_depth_bins=D)
        assert_cost_close("dot", b[0], oc, what="chunk-planar dot")
    else:
        oc, _, _, om = O.forward_mlp(**tc, weights=O.mlp_weights_from_state_dict(mlp_state(K, C, seed=2)),
                                     num_depth_bins=D, return_mask=True)
        assert_cost_close("mlp", b[0], oc, what="chunk-planar hero")
        assert_mask_close(b[3], om)
