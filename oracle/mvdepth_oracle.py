"""CPU restatement of the reference's multi-view depth regression loss — TEST INFRASTRUCTURE
(only tests/, smoke() and bench legs may import it).

Follows ``MVDepthLoss`` of the reference (losses.py:79-208; Equation 5 of the paper, used by
``experiment_modules/depth_model.py:477-485`` with weight 0.2) with plain torch ops in the
reference's order, the two ScriptModules it calls spelled out:
``BackprojectDepth.forward`` (utils/geometry_utils.py:51-59) and ``Project3D.forward`` (:72-89).
Differentiable (torch autograd), any floating dtype: run it in float64 for gradient checks.
Pinned against the imported reference in tests/test_mvdepth_oracle_vs_reference.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-8   # Project3D(eps), utils/geometry_utils.py:66


def _pix_coords(h: int, w: int, like: torch.Tensor) -> torch.Tensor:
    """pixel centres (x + 0.5, y + 0.5, 1) as (1, 3, h*w)   (utils/geometry_utils.py:34-44)"""
    xx, yy = torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy")
    pix = torch.stack((xx, yy), 0) + 0.5
    pix = torch.cat([pix, torch.ones_like(pix[:1])], 0)
    return pix.flatten(1).unsqueeze(0).to(like)


def _backproject(depth_b1hw, invK_b44, pix_13N):
    cam = torch.matmul(invK_b44[:, :3, :3], pix_13N)                 # :55
    cam = depth_b1hw.flatten(start_dim=2) * cam                      # :56
    return torch.cat([cam, torch.ones_like(cam[:, :1])], 1)          # :57


def _project(points_b4N, K_b44, cam_T_world_b44):
    P = K_b44 @ cam_T_world_b44                                      # :78
    cam = P[:, :3] @ points_b4N                                      # :80
    mask = torch.abs(cam[:, 2:]) > EPS                               # :83
    depth = cam[:, 2:] + EPS                                         # :84
    scale = torch.where(mask, 1.0 / depth, torch.ones_like(depth))   # :85
    return torch.cat([cam[:, :2] * scale, depth], 1)                 # :87-89


def valid_mask(cur_depth_b1hw, src_depth_b1hw, cur_invK_b44, src_K_b44, cur_world_T_cam_b44, src_cam_T_world_b44):
    """losses.py:90-135 -> (valid_mask_b1hw, src_depth_sampled_b1hw)"""
    h, w = cur_depth_b1hw.shape[2:]
    pix = _pix_coords(h, w, cur_depth_b1hw)
    world = cur_world_T_cam_b44 @ _backproject(cur_depth_b1hw, cur_invK_b44, pix)      # :102-103
    cam = _project(world, src_K_b44, src_cam_T_world_b44).view(-1, 3, h, w)            # :106-108
    uv = cam[:, :2].permute(0, 2, 3, 1) / torch.tensor([w, h]).view(1, 1, 1, 2).type_as(cam)   # :112-116
    uv = 2 * uv - 1
    sampled = F.grid_sample(src_depth_b1hw, uv, padding_mode="zeros", mode="nearest", align_corners=False)  # :119-125
    z = cam[:, 2:]
    valid = (z < 1.05 * sampled) & (z > 0) & (sampled > 0)                              # :127-131
    return valid, sampled


def error_for_pair(depth_pred_b1hw, cur_depth_b1hw, src_depth_b1hw, cur_invK_b44, src_K_b44, cur_world_T_cam_b44,
                   src_cam_T_world_b44):
    """losses.py:138-178"""
    h, w = cur_depth_b1hw.shape[2:]
    valid, sampled = valid_mask(cur_depth_b1hw, src_depth_b1hw, cur_invK_b44, src_K_b44, cur_world_T_cam_b44,
                                src_cam_T_world_b44)
    pix = _pix_coords(h, w, depth_pred_b1hw)
    world = cur_world_T_cam_b44 @ _backproject(depth_pred_b1hw, cur_invK_b44, pix)      # :158-159
    z_pred = _project(world, src_K_b44, src_cam_T_world_b44).view(-1, 3, h, w)[:, 2:]   # :161-166
    diff = torch.abs(torch.log(sampled) - torch.log(z_pred)).masked_select(valid)       # :168-171
    return diff.nanmean()                                                               # :173


def mv_depth_loss(depth_pred_b1hw, cur_depth_b1hw, src_depth_bk1hw, cur_invK_b44, src_K_bk44, cur_world_T_cam_b44,
                  src_cam_T_world_bk44):
    """losses.py:180-208: mean over the source views of the per-view (whole-batch) nanmean."""
    k = src_depth_bk1hw.shape[1]
    loss = 0
    for i in range(k):
        loss = loss + error_for_pair(depth_pred_b1hw, cur_depth_b1hw, src_depth_bk1hw[:, i], cur_invK_b44,
                                     src_K_bk44[:, i], cur_world_T_cam_b44, src_cam_T_world_bk44[:, i])
    return loss / k
