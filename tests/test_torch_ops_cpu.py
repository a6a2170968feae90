"""``b200cv::*`` operator registration (simplerecon_b200/torch_ops.py) — what can be checked
without a GPU: schemas, the fake (meta) kernels' shapes / dtypes, the autograd wiring on meta
tensors, argument validation, and that there is NO CPU kernel to fall back to."""
import pytest
import torch

import simplerecon_b200.torch_ops  # noqa: F401  (registers the operators)

B, K, C, H, W, D = 2, 7, 16, 12, 16, 8


def m(*s, dtype=torch.float32, device="meta"):
    return torch.empty(*s, dtype=dtype, device=device)


def dot_args(device="meta", planes=None):
    return (m(B, C, H, W, device=device), m(B, K, C, H, W, device=device), m(B, K, 4, 4, device=device),
            m(B, K, 4, 4, device=device), m(B, 4, 4, device=device),
            planes if planes is not None else m(B, D, device=device))


def mlp_args(device="meta", F=C * (K + 1) + 10 * K + 4):
    cur, src, E, Ks, invK, planes = dot_args(device)
    return (cur, src, E, m(B, K, 4, 4, device=device), Ks, invK, planes, m(128, F, device=device),
            m(128, device=device), m(128, 128, device=device), m(128, device=device),
            m(1, 128, device=device), m(1, device=device))


def test_schemas():
    s = str(torch.ops.b200cv.dot_forward.default._schema)
    assert s.startswith("b200cv::dot_forward(Tensor cur_feats, Tensor src_feats, Tensor src_extrinsics, "
                        "Tensor src_Ks, Tensor cur_invK, Tensor planes)") and s.endswith("-> (Tensor, Tensor)")
    s = str(torch.ops.b200cv.mlp_forward.default._schema)
    assert "Tensor src_poses" in s and "Tensor w3, Tensor b3" in s and s.endswith("-> (Tensor, Tensor, Tensor)")
    s = str(torch.ops.b200cv.dot_backward.default._schema)
    assert s.startswith("b200cv::dot_backward(Tensor grad_cost") and s.endswith("-> (Tensor, Tensor)")
    s = str(torch.ops.b200cv.mlp_backward.default._schema)
    assert s.startswith("b200cv::mlp_backward(Tensor grad_cost") and s.count("Tensor") == 14 + 8


@pytest.mark.parametrize("per_pixel", [False, True])
def test_fake_shapes(per_pixel):
    planes = m(B, D, H, W) if per_pixel else m(B, D)
    cost, lowest = torch.ops.b200cv.dot_forward(*dot_args(planes=planes))
    assert cost.shape == (B, D, H, W) and lowest.shape == (B, H, W) and cost.dtype == torch.float32
    cost, lowest, mask = torch.ops.b200cv.mlp_forward(*mlp_args())
    assert cost.shape == (B, D, H, W) and lowest.shape == (B, H, W)
    assert mask.shape == (B, H, W) and mask.dtype == torch.bool
    gcur, gsrc = torch.ops.b200cv.dot_backward(m(B, D, H, W), *dot_args())
    assert gcur.shape == (B, C, H, W) and gsrc.shape == (B, K, C, H, W)


def test_fake_tensor_mode_cuda_device():
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        cost, lowest = torch.ops.b200cv.dot_forward(*dot_args(device="cuda"))
        assert cost.device.type == "cuda" and cost.shape == (B, D, H, W)


def test_autograd_wiring_on_meta():
    cur, src, E, Ks, invK, planes = dot_args()
    cur.requires_grad_(True)
    src.requires_grad_(True)
    cost, lowest = torch.ops.b200cv.dot_forward(cur, src, E, Ks, invK, planes)
    assert cost.requires_grad and not lowest.requires_grad      # lowest comes from an argmax
    cost.sum().backward()
    assert cur.grad.shape == cur.shape and src.grad.shape == src.shape


def test_mlp_autograd_wiring_on_meta():
    args = list(mlp_args())
    for i in (0, 1, 7, 8, 9, 10, 11, 12):            # features and the six MLP parameters
        args[i].requires_grad_(True)
    cost, lowest, mask = torch.ops.b200cv.mlp_forward(*args)
    assert cost.requires_grad and not lowest.requires_grad and not mask.requires_grad
    cost.sum().backward()
    for i in (0, 1, 7, 8, 9, 10, 11, 12):
        assert args[i].grad is not None and args[i].grad.shape == args[i].shape
    for i in (2, 3, 4, 5, 6):                        # cameras and planes: no gradient
        assert args[i].grad is None
    outs = torch.ops.b200cv.mlp_backward(m(B, D, H, W), *mlp_args())
    assert [tuple(o.shape) for o in outs] == [tuple(a.shape) for j, a in enumerate(mlp_args()) if j in (0, 1, 7, 8, 9, 10, 11, 12)]


def test_validation():
    cur, src, E, Ks, invK, planes = dot_args()
    with pytest.raises(ValueError, match="cur_feats shape"):
        torch.ops.b200cv.dot_forward(m(B, C, H, W + 1), src, E, Ks, invK, planes)
    with pytest.raises(ValueError, match="planes must be"):
        torch.ops.b200cv.dot_forward(cur, src, E, Ks, invK, m(B + 1, D))
    with pytest.raises(ValueError, match="float32"):
        torch.ops.b200cv.dot_forward(cur, src.half(), E, Ks, invK, planes)
    with pytest.raises(ValueError, match="w1 must be"):
        torch.ops.b200cv.mlp_forward(*mlp_args(F=201))


def test_no_cpu_kernel():
    """CPU tensors reach no kernel: the dispatcher raises, nothing silently falls back."""
    z = lambda *s: torch.zeros(*s)
    with pytest.raises(NotImplementedError):
        torch.ops.b200cv.dot_forward(z(B, C, H, W), z(B, K, C, H, W), z(B, K, 4, 4), z(B, K, 4, 4), z(B, 4, 4),
                                     torch.ones(B, D))
