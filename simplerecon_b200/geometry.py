"""Geometry helper modules kept for interface / ``state_dict`` parity.

``CostVolumeManager`` in the reference owns a ``BackprojectDepth`` (buffer
``pix_coords_13N``) and a ``Project3D`` (buffer ``eps``); Lightning loads
checkpoints strictly, so a drop-in has to expose the same buffers
(reference utils/geometry_utils.py:22-89).  The fused kernels fold this arithmetic
into the sweep (csrc/srcv_common.cuh); these modules are the stand-alone forms.
"""
import torch
from torch import Tensor, nn


class BackprojectDepth(nn.Module):
    """Pixel -> camera-space points at given depths (reference
    utils/geometry_utils.py:22-59; +0.5 pixel centres, xy order)."""

    def __init__(self, height: int, width: int):
        super().__init__()
        self.height, self.width = height, width
        v, u = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
        pix = torch.stack([u + 0.5, v + 0.5, torch.ones_like(u, dtype=torch.float32)], 0)
        self.register_buffer("pix_coords_13N", pix.flatten(1).unsqueeze(0).float())

    def forward(self, depth_b1hw: Tensor, invK_b44: Tensor) -> Tensor:
        cam = torch.matmul(invK_b44[:, :3, :3], self.pix_coords_13N)
        cam = depth_b1hw.flatten(start_dim=2) * cam
        return torch.cat([cam, torch.ones_like(cam[:, :1])], 1)


class Project3D(nn.Module):
    """Camera-space points -> pixel coordinates + depth with the reference's
    guarded divide (reference utils/geometry_utils.py:62-89)."""

    def __init__(self, eps: float = 1e-8):
        super().__init__()
        self.register_buffer("eps", torch.tensor(eps).view(1, 1, 1))

    def forward(self, points_b4N: Tensor, K_b44: Tensor, cam_T_world_b44: Tensor) -> Tensor:
        P = K_b44 @ cam_T_world_b44
        cam = P[:, :3] @ points_b4N
        z = cam[:, 2:]
        depth = z + self.eps
        scale = torch.where(z.abs() > self.eps, 1.0 / depth, torch.ones_like(depth))
        return torch.cat([cam[:, :2] * scale, depth], 1)


def pose_distance(pose_b44: Tensor):
    """DVMVS pose distance (reference utils/geometry_utils.py:178-191)."""
    R, t = pose_b44[:, :3, :3], pose_b44[:, :3, 3]
    tr = R.diagonal(dim1=-1, dim2=-2).sum(-1)
    r_meas = torch.sqrt(2 * (1 - torch.minimum(torch.full_like(tr, 3.0), tr) / 3))
    t_meas = torch.norm(t, dim=1)
    return torch.sqrt(t_meas ** 2 + r_meas ** 2), r_meas, t_meas
