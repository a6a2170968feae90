"""GPU: the ``b200cv::*`` torch operators run the same C-ABI sweeps as the manager classes
(which tests/test_gpu_parity.py pins to the oracle / golden vectors), so their outputs must
agree with the managers' to rounding noise, under inference_mode and through autograd.
(File name sorts last on purpose: operator-surface checks run after the parity suite.)"""
import pytest
import torch

import simplerecon_b200 as S
import simplerecon_b200.torch_ops  # noqa: F401
from simplerecon_b200.synthetic import make_tuple, mlp_state, to_device

pytestmark = pytest.mark.gpu


def _case(kind, B=2, K=7, C=16, H=48, W=64, D=16, seed=5):
    inputs = to_device(make_tuple(batch=B, views=K, height=H, width=W, channels=C, seed=seed), "cuda")
    if kind == "dot":
        mgr = S.CostVolumeManager(H, W, num_depth_bins=D)
    else:
        mgr = S.FeatureVolumeManager(H, W, num_depth_bins=D, mlp_channels=[0, 128, 128, 1],
                                     matching_dim_size=C, num_source_views=K)
        mgr.load_state_dict({**mgr.state_dict(), **mlp_state(views=K, channels=C, seed=0)})
    return inputs, mgr.cuda().eval()


def _close(a, b, rel=1e-6):
    assert a.shape == b.shape
    tol = rel * float(b.abs().max()) + 1e-7
    err = float((a - b).abs().max())
    assert err <= tol, f"max-abs {err:.3e} > {tol:.3e}"


def test_dot_forward_matches_manager(cuda_device):
    inputs, mgr = _case("dot")
    with torch.inference_mode():
        cost_m, low_m, planes_m, _ = mgr(**inputs)
        planes_bd = planes_m[:, :, 0, 0].contiguous()
        cost, low = torch.ops.b200cv.dot_forward(inputs["cur_feats"], inputs["src_feats"],
                                                 inputs["src_extrinsics"], inputs["src_Ks"],
                                                 inputs["cur_invK"], planes_bd)
        # per-pixel form of the same planes
        cost_pp, low_pp = torch.ops.b200cv.dot_forward(inputs["cur_feats"], inputs["src_feats"],
                                                       inputs["src_extrinsics"], inputs["src_Ks"],
                                                       inputs["cur_invK"], planes_m.contiguous())
    torch.cuda.synchronize()
    _close(cost, cost_m)
    _close(low, low_m)
    _close(cost_pp, cost_m, rel=2e-6)
    _close(low_pp, low_m)


def test_mlp_forward_matches_manager(cuda_device):
    inputs, mgr = _case("mlp")
    lin = [l for l in mgr.mlp.net if isinstance(l, torch.nn.Linear)]
    with torch.inference_mode():
        cost_m, low_m, planes_m, mask_m = mgr(**inputs, return_mask=True)
        cost, low, mask = torch.ops.b200cv.mlp_forward(
            inputs["cur_feats"], inputs["src_feats"], inputs["src_extrinsics"], inputs["src_poses"],
            inputs["src_Ks"], inputs["cur_invK"], planes_m[:, :, 0, 0].contiguous(),
            lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias)
    torch.cuda.synchronize()
    _close(cost, cost_m, rel=1e-5)
    _close(low, low_m)
    assert mask.dtype == torch.bool and torch.equal(mask, mask_m)


def test_dot_autograd_matches_manager(cuda_device):
    inputs, mgr = _case("dot", H=24, W=32, D=8)
    g = torch.randn(2, 8, 24, 32, device="cuda", generator=torch.Generator("cuda").manual_seed(1))

    def grads(fn):
        cur = inputs["cur_feats"].clone().requires_grad_(True)
        src = inputs["src_feats"].clone().requires_grad_(True)
        cost = fn(cur, src)
        (cost * g).sum().backward()
        return cur.grad, src.grad

    with torch.no_grad():                       # the kernel's own plane depths (FROM_RANGE mode)
        planes_bd = mgr(**inputs)[2][:, :, 0, 0].contiguous()
    ref = grads(lambda cur, src: mgr(**{**inputs, "cur_feats": cur, "src_feats": src})[0])
    ours = grads(lambda cur, src: torch.ops.b200cv.dot_forward(
        cur, src, inputs["src_extrinsics"], inputs["src_Ks"], inputs["cur_invK"], planes_bd)[0])
    torch.cuda.synchronize()
    _close(ours[0], ref[0], rel=1e-5)          # atomics: summation order differs run to run
    _close(ours[1], ref[1], rel=1e-5)


def test_opcheck_schema_and_fake(cuda_device):
    inputs, _ = _case("dot", H=24, W=32, D=8)
    planes = torch.linspace(0.5, 4.0, 8, device="cuda").repeat(2, 1)
    args = (inputs["cur_feats"], inputs["src_feats"], inputs["src_extrinsics"], inputs["src_Ks"],
            inputs["cur_invK"], planes)
    torch.library.opcheck(torch.ops.b200cv.dot_forward.default, args,
                          test_utils=("test_schema", "test_faketensor"))
