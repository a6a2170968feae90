"""CPU: the multi-view depth-loss kernels (csrc/srcv_mvloss.cu) compiled for the host (tests/emu) and
the Python mirror of the reference's MVDepthLoss, against the oracle (value, masks, gradient)."""
import contextlib
import types

import pytest
import torch

from oracle import mvdepth_oracle as M
from simplerecon_b200 import _native, losses as L
from simplerecon_b200.synthetic import make_mvloss_batch
from tests import emu


@pytest.fixture()
def emulated(monkeypatch):
    lib = emu.load_or_skip()
    monkeypatch.setattr(_native, "_lib", lib)
    monkeypatch.setattr(L, "_require_cuda", lambda t: None)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: types.SimpleNamespace(cuda_stream=0))
    real_empty = torch.empty

    def aligned_empty(*size, **kw):
        if kw.get("dtype") is torch.uint8 and len(size) == 1 and isinstance(size[0], int):
            buf = real_empty(size[0] + 256, **kw)
            off = (-buf.data_ptr()) % 256
            return buf[off:off + size[0]]
        return real_empty(*size, **kw)

    monkeypatch.setattr(torch, "empty", aligned_empty)
    return lib


def check_loss_against_oracle(t, device="cpu"):
    """value within 2e-5 relative; gradient against the fp64 oracle gradient, 1e-4 of its maximum except at the
    few pixels whose validity / nearest sample flips between the fp32 kernel and the fp64 evaluation."""
    td = {k: v.to(device) for k, v in t.items()}
    B, K = t["src_depth_bk1hw"].shape[:2]
    H, W = t["cur_depth_b1hw"].shape[-2:]
    loss_fn = L.MVDepthLoss(H, W)
    pred = td["depth_pred_b1hw"].clone().requires_grad_(True)
    loss = loss_fn(**{**td, "depth_pred_b1hw": pred})
    loss.backward()
    o32 = M.mv_depth_loss(**t)
    assert abs(loss.item() - o32.item()) <= 2e-5 * abs(o32.item()) + 1e-7, (loss.item(), o32.item())
    t64 = {k: v.double() for k, v in t.items()}
    p64 = t64["depth_pred_b1hw"].clone().requires_grad_(True)
    o64 = M.mv_depth_loss(**{**t64, "depth_pred_b1hw": p64})
    o64.backward()
    g, g64 = pred.grad.cpu().double(), p64.grad
    assert g.shape == g64.shape and torch.isfinite(g).all()
    bad = (g - g64).abs() > 1e-4 * g64.abs().max()
    assert bad.float().mean().item() < 2e-3, f"{bad.sum().item()} of {bad.numel()} gradient entries differ"
    assert g64.abs().max() > 0
    return loss_fn, td


@pytest.mark.parametrize("seed,B,K,hw", [(0, 2, 3, (24, 32)), (1, 1, 7, (30, 40)), (2, 3, 2, (17, 23)), (3, 1, 16, (12, 20))])
def test_loss_and_gradient(emulated, seed, B, K, hw):
    check_loss_against_oracle(make_mvloss_batch(seed, B, K, *hw))


def check_masks_against_oracle(t, device="cpu"):
    td = {k: v.to(device) for k, v in t.items()}
    H, W = t["cur_depth_b1hw"].shape[-2:]
    loss_fn = L.MVDepthLoss(H, W)
    for k in range(t["src_depth_bk1hw"].shape[1]):
        a = ("cur_depth_b1hw", ("src_depth_bk1hw", k), "cur_invK_b44", ("src_K_bk44", k), "cur_world_T_cam_b44",
             ("src_cam_T_world_bk44", k))
        pick = lambda d: [d[n] if isinstance(n, str) else d[n[0]][:, n[1]] for n in a]
        vm, s = loss_fn.get_valid_mask(*pick(td))
        ovm, os_ = M.valid_mask(*pick(t))
        assert vm.dtype == torch.bool and vm.shape == ovm.shape
        assert (vm.cpu() != ovm).float().mean().item() < 1e-3
        same = s.cpu() == os_
        assert same.float().mean().item() > 0.999          # nearest sample: an index flip changes the value
        assert 0.05 < ovm.float().mean().item() < 0.99


def test_valid_mask_and_pair_error(emulated):
    t = make_mvloss_batch(4, 2, 2, 20, 28)
    check_masks_against_oracle(t)
    loss_fn = L.MVDepthLoss(20, 28)
    e = loss_fn.get_error_for_pair(t["depth_pred_b1hw"], t["cur_depth_b1hw"], t["src_depth_bk1hw"][:, 1], t["cur_invK_b44"],
                                   t["src_K_bk44"][:, 1], t["cur_world_T_cam_b44"], t["src_cam_T_world_bk44"][:, 1])
    o = M.error_for_pair(t["depth_pred_b1hw"], t["cur_depth_b1hw"], t["src_depth_bk1hw"][:, 1], t["cur_invK_b44"],
                         t["src_K_bk44"][:, 1], t["cur_world_T_cam_b44"], t["src_cam_T_world_bk44"][:, 1])
    assert abs(e.item() - o.item()) <= 2e-5 * o.item()


def test_nan_terms_are_dropped_and_half_inputs_are_upcast(emulated):
    t = make_mvloss_batch(7, 1, 2, 16, 20)
    vm, _ = M.valid_mask(t["cur_depth_b1hw"], t["src_depth_bk1hw"][:, 0], t["cur_invK_b44"], t["src_K_bk44"][:, 0],
                         t["cur_world_T_cam_b44"], t["src_cam_T_world_bk44"][:, 0])
    idx = vm.flatten().nonzero()[0].item()
    t["depth_pred_b1hw"].view(-1)[idx] = -50.0       # behind the source cameras: log() is NaN, nanmean drops it
    loss_fn = L.MVDepthLoss(16, 20)
    pred = t["depth_pred_b1hw"].clone().requires_grad_(True)
    loss = loss_fn(**{**t, "depth_pred_b1hw": pred})
    loss.backward()
    o = M.mv_depth_loss(**t)
    assert torch.isfinite(loss) and abs(loss.item() - o.item()) <= 2e-5 * o.item()
    assert pred.grad.view(-1)[idx].item() == 0.0 and torch.isfinite(pred.grad).all()   # as the reference's autograd
    # bf16 prediction (autocast): computed in fp32 from the rounded values, gradient comes back as bf16
    t2 = make_mvloss_batch(8, 1, 2, 16, 20)
    ph = t2["depth_pred_b1hw"].bfloat16().requires_grad_(True)
    lh = loss_fn(**{**t2, "depth_pred_b1hw": ph})
    lh.backward()
    oh = M.mv_depth_loss(**{**t2, "depth_pred_b1hw": ph.detach().float()})
    assert ph.grad.dtype == torch.bfloat16 and abs(lh.item() - oh.item()) <= 2e-5 * oh.item()


def test_argument_validation(emulated):
    t = make_mvloss_batch(9, 1, 2, 8, 12)
    loss_fn = L.MVDepthLoss(8, 12)
    with pytest.raises(ValueError):
        loss_fn(**{**t, "src_K_bk44": t["src_K_bk44"][:, :1]})
    big = make_mvloss_batch(9, 1, 17, 8, 12)
    with pytest.raises(_native.SrcvError):
        loss_fn(**big)
