"""CPU: the plain-C restatement (oracle/cv_oracle.c) against the golden vectors of the
unmodified reference and against the torch oracle — two independent restatements agreeing
with each other and with the reference."""
import pytest
import torch

from oracle import c_oracle as CO
from oracle import costvolume_oracle as O
from simplerecon_b200.synthetic import make_tuple, mlp_state
from tests.parity import (assert_cost_close, assert_lowest_close, assert_mask_close, golden_names,
                          load_golden)


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_matches_golden(name):
    g, inputs, sd = load_golden(name)
    if g["kind"] == "dot":
        cost, lowest, planes, mask = CO.forward_dot(**inputs, num_depth_bins=g["D"])
    else:
        cost, lowest, planes, mask = CO.forward_mlp(**inputs, weights=O.mlp_weights_from_state_dict(sd),
                                                    num_depth_bins=g["D"], return_mask=True)
    assert_cost_close(g["kind"], cost, g["ref_cost"], g["ref_cost64"], what=name)
    assert_lowest_close(g["kind"], lowest, planes, g["ref_cost"], what=name)
    if g["kind"] == "mlp":
        assert_mask_close(mask, g["ref_mask"], what=name)


def test_c_oracle_fp64_matches_reference_fp64():
    g, inputs, sd = load_golden("hero_mini_24x32_D8_K7")
    w = O.mlp_weights_from_state_dict(sd)
    cost, *_ = CO.forward_mlp(**inputs, weights=w, num_depth_bins=g["D"], double=True)
    assert (cost - g["ref_cost64"]).abs().max().item() <= 1e-9 * float(g["ref_cost64"].abs().max()) + 1e-12
    g, inputs, sd = load_golden("cfg0_dot_48x64_D16_K2")
    cost, *_ = CO.forward_dot(**inputs, num_depth_bins=g["D"], double=True)
    assert (cost - g["ref_cost64"]).abs().max().item() <= 1e-9 * float(g["ref_cost64"].abs().max()) + 1e-12


@pytest.mark.parametrize("kind", ["dot", "mlp"])
def test_c_oracle_agrees_with_torch_oracle(kind):
    t = make_tuple(2, 3, 17, 23, seed=123)
    if kind == "dot":
        a, al, ap, _ = CO.forward_dot(**t, num_depth_bins=6)
        b, bl, bp, _ = O.forward_dot(**t, num_depth_bins=6)
    else:
        w = O.mlp_weights_from_state_dict(mlp_state(3, 16, seed=9))
        a, al, ap, am = CO.forward_mlp(**t, weights=w, num_depth_bins=6, return_mask=True)
        b, bl, bp, bm = O.forward_mlp(**t, weights=w, num_depth_bins=6, return_mask=True)
        assert_mask_close(am, bm)
    assert_cost_close(kind, a, b)
    assert torch.allclose(ap.expand_as(bp), bp, rtol=1e-6)
