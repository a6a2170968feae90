"""Randomised differential test of the host-emulated MVDepthLoss kernels (tests/emu) against the oracle:
random batch / view counts and map sizes, noisy predictions, source views looking away (empty valid
sets: the loss is NaN like the reference's nanmean of nothing), predictions behind a source camera
(NaN terms dropped), zero ground-truth depth.  CPU only.

    python scripts/emu_fuzz_mvloss.py [--cases 100] [--seed 0]
"""
import argparse
import contextlib
import math
import sys
import types
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import mvdepth_oracle as M  # noqa: E402
from simplerecon_b200 import _native, losses as L  # noqa: E402
from simplerecon_b200.synthetic import _axis_angle, make_mvloss_batch  # noqa: E402
from tests import emu  # noqa: E402


def patch():
    _native._lib = emu.load()
    L._require_cuda = lambda t: None
    torch.cuda.device = lambda dev: contextlib.nullcontext()
    torch.cuda.current_stream = lambda dev=None: types.SimpleNamespace(cuda_stream=0)
    real_empty = torch.empty

    def aligned_empty(*size, **kw):
        if kw.get("dtype") is torch.uint8 and len(size) == 1 and isinstance(size[0], int):
            buf = real_empty(size[0] + 256, **kw)
            off = (-buf.data_ptr()) % 256
            return buf[off:off + size[0]]
        return real_empty(*size, **kw)
    torch.empty = aligned_empty


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    patch()
    g = torch.Generator().manual_seed(a.seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    fails = 0
    for i in range(a.cases):
        B, K, H, W = ri(1, 3), [1, 2, 3, 7, 16][ri(0, 4)], ri(3, 40), ri(4, 48)
        t = make_mvloss_batch(ri(0, 10 ** 6), B, K, H, W, pred_noise=[0.02, 0.3, 1.0][ri(0, 2)])
        mode = ri(0, 4)
        if mode == 1:                                 # one source view turned away: its valid set may be empty
            k = ri(0, K - 1)
            R = torch.eye(4)
            R[:3, :3] = _axis_angle(torch.tensor([[0.0, 1.0, 0.0]], dtype=torch.float64), torch.tensor([math.pi * 0.9], dtype=torch.float64))[0].float()
            t["src_cam_T_world_bk44"][:, k] = R @ t["src_cam_T_world_bk44"][:, k]
        elif mode == 2:                               # some predictions far behind the cameras
            m = torch.rand(t["depth_pred_b1hw"].shape, generator=g) < 0.05
            t["depth_pred_b1hw"][m] = -20.0
        elif mode == 3:                               # holes in the ground truth
            t["cur_depth_b1hw"][torch.rand(t["cur_depth_b1hw"].shape, generator=g) < 0.1] = 0.0
        fn = L.MVDepthLoss(H, W)
        p = t["depth_pred_b1hw"].clone().requires_grad_(True)
        loss = fn(**{**t, "depth_pred_b1hw": p})
        loss.backward()
        o32 = M.mv_depth_loss(**t)
        t64 = {k2: v.double() for k2, v in t.items()}
        p64 = t64["depth_pred_b1hw"].clone().requires_grad_(True)
        o64 = M.mv_depth_loss(**{**t64, "depth_pred_b1hw": p64})
        ok = True
        if torch.isnan(o32):
            ok = bool(torch.isnan(loss))
            note = "both NaN (a view without valid pixels)" if ok else f"ours {loss.item()} vs NaN"
        else:
            o64.backward()
            # fp32 validity / nearest-sample decisions can differ from the fp32 oracle at a few pixels of a tiny map
            rel = abs(loss.item() - o32.item()) / max(abs(o32.item()), 1e-12)
            gg, g64 = p.grad.double(), p64.grad
            bad = ((gg - g64).abs() > 1e-4 * g64.abs().max().clamp_min(1e-30)).float().mean().item()
            ok = (rel <= 5e-5 or rel <= 3.0 / (B * H * W)) and bad <= max(4e-3, 2.0 / (B * H * W)) and bool(torch.isfinite(gg).all())
            note = f"loss rel {rel:.1e}, gradient entries off {bad:.4f}"
        fails += (not ok)
        if not ok or i % 10 == 0:
            print(f"[{i}] {'ok  ' if ok else 'FAIL'} B{B} K{K} {H}x{W} mode {mode}: {note}", flush=True)
    print(f"{a.cases} cases, {fails} failures")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
