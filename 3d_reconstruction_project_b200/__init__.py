"""b200recon -- B200-native point-cloud front end + registration (depth -> cloud -> voxel -> normals -> ICP/GICP).

Drop-in for the hot path of aagsi/3D_Reconstruction_Project: the five classes ``main.py`` imports keep their names and
signatures (``RealSensePipeline``, ``PointCloudCapture``, ``PointCloudAlignment``, ``PointCloudProcessingWithCUDA``,
``NormalEstimation``); all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of ``libb200recon.so``
(include/b200recon.h). The directory name starts with a digit, so import it as ``b200recon`` (alias package) or with
``importlib.import_module("3d_reconstruction_project_b200")``.
"""
from . import _native  # noqa: F401  (ctypes signatures; loading is lazy so CPU-only tooling can import the package)
from .geometry import PointCloud, Vector3dVector  # noqa: F401
from .realsense_pipeline import RealSensePipeline, ReplayPipeline  # noqa: F401
from .pointcloud_capture import PointCloudCapture  # noqa: F401
from .pointcloud_alignment import PointCloudAlignment  # noqa: F401
from .pointcloud_processing import PointCloudProcessingWithCUDA  # noqa: F401
from .normal_estimation import NormalEstimation  # noqa: F401

__all__ = ["PointCloud", "Vector3dVector", "RealSensePipeline", "ReplayPipeline", "PointCloudCapture", "PointCloudAlignment",
           "PointCloudProcessingWithCUDA", "NormalEstimation"]
