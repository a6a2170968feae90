"""simplerecon_b200 — B200-native plane-sweep cost volume behind SimpleRecon's
``CostVolumeManager`` / ``FeatureVolumeManager`` API.

Only the hot path named by BASELINE.json is here: the cost-volume build of the
reference's ``modules/cost_volume.py``, as hand-written sm_100a kernels in
``csrc/`` behind a C ABI (``include/srcv_b200.h``), plus the Python mirror of the
reference's manager classes that binds it.  Encoders, decoder, datasets and
training stay the reference's own PyTorch code.
"""
from .cost_volume import (CostVolumeManager, FastFeatureVolumeManager, FeatureVolumeManager,
                          instance_norm_to_chunk_planar)
from .geometry import BackprojectDepth, Project3D, pose_distance
from .install import install, uninstall
from .networks import MLP
from . import torch_ops  # noqa: E402,F401  registers torch.ops.b200cv.{dot_forward,dot_backward,mlp_forward}

__all__ = [
    "CostVolumeManager", "FeatureVolumeManager", "FastFeatureVolumeManager", "MLP",
    "BackprojectDepth", "Project3D", "pose_distance", "install", "uninstall",
]
__version__ = "0.1.0"
from .tsdf import TSDF, TSDFFuser  # noqa: E402,F401  (reference tools/tsdf.py)
from . import point_cloud_fusion  # noqa: E402,F401  (reference tools/torch_point_cloud_fusion.py)
from .losses import MVDepthLoss  # noqa: E402,F401  (reference losses.py:79-208)
