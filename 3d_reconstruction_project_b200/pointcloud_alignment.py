"""PointCloudAlignment -- reference pointcloud_alignment.py:5-43. Same signature and progress prints; the default is the
reference's point-to-point ICP. ``method`` additionally selects the estimators the reference uses elsewhere on this path:
"point_to_plane" (test/mini1.py:293-296) and "gicp" (test/GICP1.py:99-102)."""
import numpy as np

from . import registration as reg
from .geometry import KDTreeSearchParamHybrid, as_cloud


class PointCloudAlignment:
    def align_point_clouds(self, source, target, threshold=0.02, voxel_size=0.01, max_iter=100, method="point_to_point"):
        source, target = as_cloud(source), as_cloud(target)
        print("Downsampling point clouds using voxel size:", voxel_size)
        source = source.voxel_down_sample(voxel_size=voxel_size)
        target = target.voxel_down_sample(voxel_size=voxel_size)

        print("Estimating normals on CPU...")  # message kept verbatim; the estimation runs on the GPU here
        source.estimate_normals(search_param=KDTreeSearchParamHybrid(radius=voxel_size * 2, max_nn=30))
        target.estimate_normals(search_param=KDTreeSearchParamHybrid(radius=voxel_size * 2, max_nn=30))

        trans_init = np.eye(4)
        print("Performing ICP alignment using CUDA...")
        criteria = reg.ICPConvergenceCriteria(relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=max_iter)
        if method == "point_to_point":
            result = reg.registration_icp(source, target, threshold, trans_init, reg.TransformationEstimationPointToPoint(), criteria)
        elif method == "point_to_plane":
            result = reg.registration_icp(source, target, threshold, trans_init, reg.TransformationEstimationPointToPlane(), criteria)
        elif method == "gicp":
            result = reg.registration_generalized_icp(source, target, threshold, trans_init, reg.TransformationEstimationForGeneralizedICP(), criteria)
        else:
            raise ValueError(f"unknown method {method!r}")
        self.last_result = result
        source.transform(result.transformation)
        return source
