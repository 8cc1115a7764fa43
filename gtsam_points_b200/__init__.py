"""gtsam_points_b200 -- B200-native (sm_100a) scan-matching linearization behind gtsam_points' factor interface.

Only the hot path of koide3/gtsam_points is implemented here: voxel-hash / kd-tree
correspondence search, SE(3) transform, Mahalanobis residual and Jacobians, and the reduction into each factor's
H and b.  The compute lives in lib/libb2points.so (hand-written CUDA, C ABI in include/b2points.h); this package is
the thin host-side mirror of the reference's interface.  There is no CPU fallback.
"""
from . import capi  # noqa: F401
from .capi import B2Error  # noqa: F401
from .factors import (  # noqa: F401
    HessianFactor,
    IntegratedGICPFactor,
    IntegratedICPFactor,
    IntegratedPointToPlaneICPFactor,
    IntegratedMatchingCostFactor,
    IntegratedVGICPFactor,
    IntegratedVGICPFactorGPU,
    NonlinearFactorSetGPU,
    pose_inverse,
)
from .types import Context, GaussianVoxelMapGPU, KdTree, PointCloud, default_context, estimate_covariances, merge_frames_gpu, overlap_gpu  # noqa: F401

__version__ = "0.1.0"
