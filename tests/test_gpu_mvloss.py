"""GPU: the multi-view depth-loss kernels through the C ABI / the MVDepthLoss mirror against the CPU
oracle (value, masks, gradient vs fp64 autograd), at test sizes and at the training resolution."""
import pytest
import torch

from oracle import mvdepth_oracle as M
from simplerecon_b200 import _native, losses as L
from simplerecon_b200.synthetic import make_mvloss_batch
from tests.test_emu_mvloss import check_loss_against_oracle, check_masks_against_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,B,K,hw", [(0, 2, 3, (24, 32)), (1, 1, 7, (30, 40)), (2, 3, 2, (17, 23)), (3, 2, 16, (48, 64))])
def test_loss_and_gradient(cuda_device, seed, B, K, hw):
    check_loss_against_oracle(make_mvloss_batch(seed, B, K, *hw), device="cuda")
    assert _native.last_variant() == "mvloss_backward_f32"


def test_valid_masks(cuda_device):
    check_masks_against_oracle(make_mvloss_batch(4, 2, 3, 40, 56), device="cuda")


def test_training_resolution_is_deterministic(cuda_device):
    """depth_pred_s0 of the 512x384 training config is 192x256 (options.py:70-71, depth_model.py:477): batch 4,
    7 source views; two runs give the same bits (fixed-order reduction), value equals the oracle."""
    t = make_mvloss_batch(11, 4, 7, 192, 256)
    td = {k: v.cuda() for k, v in t.items()}
    loss_fn = L.MVDepthLoss(192, 256)
    outs = []
    for _ in range(2):
        p = td["depth_pred_b1hw"].clone().requires_grad_(True)
        l = loss_fn(**{**td, "depth_pred_b1hw": p})
        l.backward()
        outs.append((l.item(), p.grad.clone()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])
    o = M.mv_depth_loss(**t)
    assert abs(outs[0][0] - o.item()) <= 2e-5 * o.item()


def test_nan_terms_and_bf16(cuda_device):
    t = make_mvloss_batch(7, 1, 2, 16, 20)
    vm, _ = M.valid_mask(t["cur_depth_b1hw"], t["src_depth_bk1hw"][:, 0], t["cur_invK_b44"], t["src_K_bk44"][:, 0],
                         t["cur_world_T_cam_b44"], t["src_cam_T_world_bk44"][:, 0])
    idx = vm.flatten().nonzero()[0].item()
    t["depth_pred_b1hw"].view(-1)[idx] = -50.0
    td = {k: v.cuda() for k, v in t.items()}
    loss_fn = L.MVDepthLoss(16, 20)
    pred = td["depth_pred_b1hw"].clone().requires_grad_(True)
    loss = loss_fn(**{**td, "depth_pred_b1hw": pred})
    loss.backward()
    o = M.mv_depth_loss(**t)
    assert torch.isfinite(loss) and abs(loss.item() - o.item()) <= 2e-5 * o.item()
    assert pred.grad.view(-1)[idx].item() == 0.0 and torch.isfinite(pred.grad).all()
    ph = td["depth_pred_b1hw"].abs().bfloat16().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lh = loss_fn(**{**td, "depth_pred_b1hw": ph})
    lh.backward()
    assert ph.grad.dtype == torch.bfloat16 and torch.isfinite(lh)


def test_cpu_tensors_are_refused():
    t = make_mvloss_batch(9, 1, 2, 8, 12)
    with pytest.raises(RuntimeError):
        L.MVDepthLoss(8, 12)(**t)
