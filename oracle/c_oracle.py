"""ctypes front-end of the plain-C oracle restatement (oracle/cv_oracle.c).

TEST INFRASTRUCTURE — never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libcvoracle.so"
_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB.is_file() or LIB.stat().st_mtime < (HERE / "cv_oracle.c").stat().st_mtime:
            subprocess.run(["make", "-C", str(HERE), "-s"], check=True)
        _lib = C.CDLL(str(LIB))
    return _lib


def _np(t, dtype):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(dtype, copy=False))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _planes(min_depth, max_depth, D, B, H, W, depth_planes_bdhw, dtype):
    if depth_planes_bdhw is None:
        tdt = torch.float32 if dtype == np.float32 else torch.float64
        mn = torch.as_tensor(min_depth, dtype=tdt).reshape(())
        mx = torch.as_tensor(max_depth, dtype=tdt).reshape(())
        ramp = torch.linspace(0, 1, D).to(tdt)
        d = torch.exp(torch.log(mn) + torch.log(mx / mn) * ramp)     # reference cost_volume.py:124-127
        return _np(d.view(1, D).expand(B, D), dtype), 0, d.view(1, D, 1, 1).expand(B, D, H, W)
    st = depth_planes_bdhw.stride()
    if st[2] == 0 and st[3] == 0:
        return _np(depth_planes_bdhw[:, :, 0, 0], dtype), 0, depth_planes_bdhw
    return _np(depth_planes_bdhw, dtype), 1, depth_planes_bdhw


def forward_dot(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                num_depth_bins=64, depth_planes_bdhw=None, return_mask=False, double=False):
    lib = load()
    dt = np.float64 if double else np.float32
    B, K, Cc, H, W = src_feats.shape
    D = num_depth_bins if depth_planes_bdhw is None else depth_planes_bdhw.shape[1]
    planes, per_pixel, planes_ret = _planes(min_depth, max_depth, D, B, H, W, depth_planes_bdhw, dt)
    cur, src = _np(cur_feats, dt), _np(src_feats, dt)
    E, Ks, iK = _np(src_extrinsics, dt), _np(src_Ks, dt), _np(cur_invK, dt)
    cost = np.empty((B, D, H, W), dt)
    lowest = np.empty((B, H, W), dt)
    fn = lib.cvo_dot_f64 if double else lib.cvo_dot_f32
    fn(B, K, Cc, H, W, D, _p(cur), _p(src), _p(E), _p(Ks), _p(iK), _p(planes), per_pixel, _p(cost), _p(lowest))
    return torch.from_numpy(cost), torch.from_numpy(lowest), planes_ret, None


def forward_mlp(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                weights, num_depth_bins=64, depth_planes_bdhw=None, return_mask=False, double=False):
    lib = load()
    dt = np.float64 if double else np.float32
    B, K, Cc, H, W = src_feats.shape
    D = num_depth_bins if depth_planes_bdhw is None else depth_planes_bdhw.shape[1]
    planes, per_pixel, planes_ret = _planes(min_depth, max_depth, D, B, H, W, depth_planes_bdhw, dt)
    cur, src = _np(cur_feats, dt), _np(src_feats, dt)
    E, P, Ks, iK = _np(src_extrinsics, dt), _np(src_poses, dt), _np(src_Ks, dt), _np(cur_invK, dt)
    w = [_np(x, dt) for x in weights]
    cost = np.empty((B, D, H, W), dt)
    lowest = np.empty((B, H, W), dt)
    mask = np.zeros((B, H, W), np.uint8)
    fn = lib.cvo_mlp_f64 if double else lib.cvo_mlp_f32
    fn(B, K, Cc, H, W, D, _p(cur), _p(src), _p(E), _p(P), _p(Ks), _p(iK), _p(planes), per_pixel,
       _p(w[0]), _p(w[1]), _p(w[2]), _p(w[3]), _p(w[4]), _p(w[5]), w[0].shape[0], w[2].shape[0],
       _p(cost), _p(lowest), _p(mask) if return_mask else None)
    return (torch.from_numpy(cost), torch.from_numpy(lowest), planes_ret,
            torch.from_numpy(mask).bool() if return_mask else None)
