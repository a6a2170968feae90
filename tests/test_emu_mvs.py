"""CPU: the multi-view consistency kernel (csrc/srcv_mvs.cu) compiled for the host (tests/emu) and
the Python mirror of the reference's process_depth / process_scene, against the oracle."""
import contextlib
import types

import numpy as np
import pytest
import torch

from oracle import mvs_oracle as M
from simplerecon_b200 import _native, point_cloud_fusion as pcf
from simplerecon_b200.synthetic import make_mvs_scene
from tests import emu


@pytest.fixture()
def emulated(monkeypatch):
    lib = emu.load_or_skip()
    monkeypatch.setattr(_native, "_lib", lib)
    monkeypatch.setattr(pcf, "_require_cuda", lambda t: None)
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


def check_against_oracle(scan, sc, ref_idx, z_thresh=0.1, n_thresh=3):
    n = sc["depths"].shape[0]
    src = torch.arange(n) != ref_idx
    inv = (scan.K_inv[ref_idx].cpu(), scan.K_inv[src].cpu(), scan.P_inv[ref_idx].cpu())
    opts, onv, ovalid = M.process_depth_dense(sc["depths"][ref_idx], sc["depths"][src], sc["cam_T_world"][ref_idx],
                                              sc["cam_T_world"][src], sc["K"][ref_idx], sc["K"][src], z_thresh, n_thresh,
                                              inverses=inv)
    pts, nv, valid = scan.consistency(ref_idx, z_thresh, n_thresh)
    pts, nv, valid = pts.cpu(), nv.cpu(), valid.cpu()
    # nearest-sample and threshold decisions are discontinuous: allow a handful of flips, exact elsewhere
    same = nv.long() == onv
    assert same.float().mean().item() > 0.999, f"n_valid differs at {(~same).sum().item()} pixels"
    assert (valid != ovalid).float().mean().item() < 1e-3
    assert 0.05 < ovalid.float().mean().item() < 0.99
    assert (pts[same] - opts[same]).abs().max().item() < 2e-5


@pytest.mark.parametrize("seed,n,hw", [(3, 6, (24, 32)), (4, 37, (18, 26))])      # 37 frames: two shared-memory chunks
def test_consistency_matches_oracle(emulated, seed, n, hw):
    sc = make_mvs_scene(seed=seed, frames=n, height=hw[0], width=hw[1])
    scan = pcf._Scan(sc["depths"], sc["cam_T_world"], sc["K"], torch.device("cpu"))
    for ref_idx in (0, n - 1, n // 2):
        check_against_oracle(scan, sc, ref_idx)


def test_process_depth_and_scene_mirror_the_reference_api(emulated):
    sc = make_mvs_scene(seed=5, frames=5, height=20, width=28)
    d, im, P, K = sc["depths"], sc["images"], sc["cam_T_world"], sc["K"]
    src = torch.arange(5) != 1
    pts, rgb, valid = pcf.process_depth(d[1], im[1], d[src], im[src], P[1], P[src], K[1], K[src], 0.1, 3)
    assert pts.shape[1] == 3 and pts.shape[0] == valid.sum() and rgb.shape == pts.shape and valid.shape == (20, 28)
    opts, _, ovalid = M.process_depth_dense(d[1], d[src], P[1], P[src], K[1], K[src], 0.1, 3)
    assert (valid != ovalid.numpy()).mean() < 2e-3
    fp, fr, av = pcf.process_scene(d, im, P, K, 0.1, 3)
    assert av.shape == (5, 20, 28) and fp.shape[0] == av.sum() and fr.dtype == np.uint8
    assert np.array_equal(av[1], valid)
