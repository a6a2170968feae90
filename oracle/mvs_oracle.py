"""CPU restatement of the reference's multi-view depth-consistency point-cloud fusion —
TEST INFRASTRUCTURE (only tests/, smoke() and bench legs may import it).

Follows ``process_depth`` of the reference (tools/torch_point_cloud_fusion.py:12-97; the
3DVNet fuser ``pc_fusion.py:158`` drives through ``process_scene`` :100-118), fp32 torch ops in
the reference's order, without its ``.cuda()`` calls.  Returns the dense per-pixel results
(averaged point, consistent-view count, validity) BEFORE the boolean compaction of :92-95, so the
kernel's outputs can be compared element-wise; ``compact`` applies the compaction.
Pinned against the imported reference in tests/test_mvs_oracle_vs_reference.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def process_depth_dense(ref_depth, src_depths, ref_P, src_Ps, ref_K, src_Ks, z_thresh=0.1,
                        n_consistent_thresh=3, inverses=None):
    """ref_depth (h,w), src_depths (n,h,w), P = world->camera 4x4, K 3x3.
    -> pts_avg (h*w,3), n_valid (h*w) int64, valid (h,w) bool.  `inverses` = (ref_K_inv,
    src_Ks_inv, ref_P_inv) lets a caller share one set of torch.inverse results."""
    n_src = src_depths.shape[0]
    h, w = int(ref_depth.shape[0]), int(ref_depth.shape[1])
    n_pts = h * w
    if inverses is None:
        ref_K_inv, src_Ks_inv, ref_P_inv = torch.inverse(ref_K), torch.inverse(src_Ks), torch.inverse(ref_P)   # :25-27
    else:
        ref_K_inv, src_Ks_inv, ref_P_inv = inverses
    pts_x = np.linspace(0, w - 1, w)
    pts_y = np.linspace(0, h - 1, h)
    pts_xx, pts_yy = np.meshgrid(pts_x, pts_y)
    pts = torch.from_numpy(np.stack((pts_xx, pts_yy, np.ones_like(pts_xx)), axis=0)).float()           # :33
    pts = ref_P_inv[:3, :3] @ (ref_K_inv @ (pts * ref_depth.unsqueeze(0)).view(3, n_pts)) \
        + ref_P_inv[:3, 3, None]                                                                         # :34-35
    pts_reproj = torch.bmm(src_Ps[:, :3, :3], pts.unsqueeze(0).repeat(n_src, 1, 1)) + src_Ps[:, :3, 3, None]   # :51-52
    pts_reproj = torch.bmm(src_Ks, pts_reproj)                                                           # :53
    z_reproj = pts_reproj[:, 2]
    pts_reproj = pts_reproj / z_reproj.unsqueeze(1)                                                      # :55
    valid_z = z_reproj > 1e-4
    valid_x = (pts_reproj[:, 0] >= 0.) & (pts_reproj[:, 0] <= float(w - 1))
    valid_y = (pts_reproj[:, 1] >= 0.) & (pts_reproj[:, 1] <= float(h - 1))
    grid = torch.clone(pts_reproj[:, :2]).transpose(2, 1).view(n_src, n_pts, 1, 2)
    grid[..., 0] = (grid[..., 0] / float(w - 1)) * 2 - 1.0                                               # :62
    grid[..., 1] = (grid[..., 1] / float(h - 1)) * 2 - 1.0
    z_sample = F.grid_sample(src_depths.unsqueeze(1), grid, mode="nearest", align_corners=True,
                             padding_mode="zeros").squeeze(1).squeeze(-1)                               # :64-66
    valid_per_src = (torch.abs(z_reproj - z_sample) < z_thresh) & valid_x & valid_y & valid_z            # :68-71
    n_valid = torch.sum(valid_per_src.int(), dim=0)
    pts_sample = torch.bmm(src_Ks_inv, pts_reproj * z_sample.unsqueeze(1))                               # :75
    pts_sample = torch.bmm(src_Ps[:, :3, :3].transpose(2, 1), pts_sample - src_Ps[:, :3, 3, None])       # :76-77
    valid = n_valid >= n_consistent_thresh                                                               # :83
    pts_avg = pts.clone()
    for i in range(n_src):                                                                               # :87-92
        p = pts_sample[i]
        bad = torch.isnan(p)
        p = torch.where(bad, torch.zeros_like(p), p)
        v = valid_per_src[i] & ~torch.any(bad, dim=0)
        pts_avg = pts_avg + p * v.float().unsqueeze(0)
    pts_avg = pts_avg / (n_valid + 1).float().unsqueeze(0).expand(3, n_pts)                              # :93
    return pts_avg.transpose(1, 0).contiguous(), n_valid.long(), valid.view(h, w)


def compact(pts_avg, valid_hw, ref_image_hw3=None):
    """:95-97: the points (and colours) of the consistent pixels."""
    v = valid_hw.reshape(-1)
    pts = pts_avg[v]
    rgb = ref_image_hw3[valid_hw].view(-1, 3) if ref_image_hw3 is not None else None
    return pts, rgb
