"""NormalEstimation -- reference normal_estimation.py:3-22: tensor estimate_normals(max_nn=50, radius=0.05) (:20).
orient_normals_consistent_tangent_plane(100) (:21) is a sequential MST propagation that only flips signs; it is a "next"
row of the scope table and is NOT applied here -- normals carry the eigen-solver's sign."""
import numpy as np

from . import ops
from .context import parse_device
from .geometry import as_cloud


class NormalEstimation:
    def __init__(self, device="CUDA:0"):
        self.device = parse_device(device)

    def estimate_normals(self, pcd):
        pcd = as_cloud(pcd, self.device).clone()
        if not pcd.has_points():
            return pcd
        pts32 = np.asarray(pcd.points).astype(np.float32)  # from_legacy(..., Float32)
        nrm = ops.estimate_normals_tensor(pts32, 50, 0.05, device=self.device)
        pcd.points = pts32.astype(np.float64)  # to_legacy of the float32 tensor cloud
        pcd.normals = nrm.astype(np.float64)
        return pcd
