"""PointCloudProcessingWithCUDA -- reference pointcloud_processing.py:4-44: load PLY, tensor voxel_down_sample (:27),
remove_statistical_outlier(30, 1.2) + select_by_index (:35-36), remove_radius_outlier(16, 0.01) (:39)."""
import numpy as np

from . import ops, plyio
from .context import parse_device
from .geometry import PointCloud, as_cloud


class PointCloudProcessingWithCUDA:
    def __init__(self, device="CUDA:0", downsample_voxel_size=0.0025):
        self.device = parse_device(device)
        self.downsample_voxel_size = downsample_voxel_size

    def process_point_cloud(self, filename):
        pcd = filename if not isinstance(filename, (str, bytes)) and hasattr(filename, "points") else plyio.read_point_cloud(filename, device=self.device)
        pcd = as_cloud(pcd, self.device)
        return self.process(pcd)

    def process(self, pcd):
        """The same chain on an in-memory cloud."""
        if not pcd.has_points():
            return PointCloud(device=self.device)
        # from_legacy(Float32) -> tensor voxel_down_sample -> to_legacy: colours ride as the float32 attribute
        pts32 = np.asarray(pcd.points).astype(np.float32)
        col32 = np.asarray(pcd.colors).astype(np.float32) if pcd.has_colors() else None
        r = ops.voxel_down_sample_tensor(pts32, self.downsample_voxel_size, attr=col32, device=self.device)
        down = PointCloud(r["points"].astype(np.float64), device=self.device)
        if r["attr"] is not None:
            down.colors = r["attr"].astype(np.float64)
        cl, ind = down.remove_statistical_outlier(nb_neighbors=30, std_ratio=1.2)
        inlier = down.select_by_index(ind)
        inlier, ind = inlier.remove_radius_outlier(nb_points=16, radius=0.01)
        return inlier
