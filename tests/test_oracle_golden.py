"""CPU: the oracle restatement against the committed golden vectors (made from the
unmodified reference by tests/golden/make_golden.py)."""
import pytest
import torch

from oracle import costvolume_oracle as O
from tests.parity import (assert_cost_close, assert_lowest_close, assert_mask_close, golden_fullsize_names,
                          golden_names, load_golden, load_golden_fullsize)


def run_oracle(g, inputs, sd, sampler, dtype=torch.float32):
    kw = {k: (v.to(dtype) if torch.is_tensor(v) and v.dtype.is_floating_point else v)
          for k, v in inputs.items()}
    if g["kind"] == "dot":
        return O.forward_dot(**kw, num_depth_bins=g["D"], sampler=sampler)
    w = tuple(x.to(dtype) for x in O.mlp_weights_from_state_dict(sd))
    return O.forward_mlp(**kw, weights=w, num_depth_bins=g["D"], return_mask=True, sampler=sampler)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("sampler", ["aten", "explicit"])
def test_oracle_matches_golden(name, sampler):
    g, inputs, sd = load_golden(name)
    cost, lowest, planes, mask = run_oracle(g, inputs, sd, sampler)
    ref = g["ref_cost"]
    if sampler == "aten":
        # same ATen ops in the same order as the reference: equal to the last bit or two
        assert (cost - ref).abs().max().item() <= 2e-6 * float(ref.abs().max()) + 1e-7
    assert_cost_close(g["kind"], cost, ref, g["ref_cost64"], what=f"{name}/{sampler}")
    assert_lowest_close(g["kind"], lowest, planes, ref, what=name)
    assert torch.equal(lowest, g["ref_lowest"]) or sampler == "explicit"
    if "ref_mask" in g:
        assert_mask_close(mask, g["ref_mask"], what=name)
    else:
        assert mask is None


@pytest.mark.parametrize("name", golden_fullsize_names())
def test_oracle_matches_golden_at_the_bench_map_size(name):
    """120 x 160 x 7 views (tests/golden/make_golden_fullsize.py): regenerated inputs hash to the stored
    value, and the explicit-sampler oracle (the arithmetic the kernels implement) meets the reference."""
    g, inputs, sd = load_golden_fullsize(name)
    cost, lowest, planes, mask = run_oracle(g, inputs, sd, "explicit")
    assert_cost_close(g["kind"], cost, g["ref_cost"], g["ref_cost64"], what=name, e_ref=g["err32v64"])
    assert_lowest_close(g["kind"], lowest, planes, g["ref_cost"], what=name)
    if g["kind"] == "mlp":
        assert_mask_close(mask, g["ref_mask"], what=name)


@pytest.mark.parametrize("name", ["hero_mini_24x32_D8_K7", "cfg0_dot_48x64_D16_K2"])
def test_oracle_fp64_matches_reference_fp64(name):
    g, inputs, sd = load_golden(name)
    cost, *_ = run_oracle(g, inputs, sd, "explicit", torch.float64)
    ref = g["ref_cost64"]
    assert (cost - ref).abs().max().item() <= 1e-9 * float(ref.abs().max()) + 1e-12


def test_depth_planes_known_answer():
    # D=5, 0.25 -> 5.0
    d = O.depth_planes(0.25, 5.0, 5)
    exp = torch.tensor([0.25, 0.5287, 1.1180, 2.3644, 5.0])
    assert torch.allclose(d, exp, atol=2e-4)


def test_pose_distance_identity_and_translation():
    P = torch.eye(4).repeat(1, 2, 1, 1)
    P[0, 1, :3, 3] = torch.tensor([0.3, 0.0, 0.4])
    comb, r, t = O.pose_distance(P)
    assert torch.allclose(r, torch.zeros(1, 2)) and torch.allclose(t, torch.tensor([[0.0, 0.5]]))
    assert torch.allclose(comb, torch.tensor([[0.0, 0.5]]))


def test_bilinear_zeros_padding_and_identity():
    # identity warp samples the texel centres exactly; outside the map gives zeros
    src = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(1, 1, 2, 3, 4)
    v, u = torch.meshgrid(torch.arange(3), torch.arange(4), indexing="ij")
    px = (u.reshape(1, 1, -1) + 0.5).float()
    py = (v.reshape(1, 1, -1) + 0.5).float()
    for s in ("explicit", "aten"):
        out = O.sample_bilinear_zeros(src, px, py, s)
        assert torch.allclose(out.reshape(2, 3, 4), src[0, 0], atol=1e-5)
        far = O.sample_bilinear_zeros(src, px + 100.0, py, s)
        assert float(far.abs().max()) == 0.0
    # half a pixel past the left border: half weight on column 0, half on padding
    edge = O.sample_bilinear_zeros(src, torch.zeros(1, 1, 1), torch.full((1, 1, 1), 0.5), "explicit")
    assert torch.allclose(edge.reshape(-1), 0.5 * src[0, 0, :, 0, 0], atol=1e-5)
