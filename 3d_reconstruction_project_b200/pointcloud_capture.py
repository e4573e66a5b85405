"""PointCloudCapture -- reference pointcloud_capture.py:5-55 with the arithmetic on the GPU:
rs.pointcloud().calculate (:35) -> deprojection kernel; colours/255 (:39) fused into the same pass;
from_legacy(Float32) + voxel_down_sample (:47-50) -> radix-sorted voxel reduction; to_legacy (:53) -> float64 host arrays."""
import numpy as np

from . import ops
from .context import parse_device
from .geometry import PointCloud
from .realsense_pipeline import Intrinsics


def _frame_intrinsics(depth_frame, pipeline):
    """Depth-stream intrinsics: replay frames carry them; real rs frames expose profile.as_video_stream_profile().intrinsics."""
    intr = getattr(depth_frame, "intrinsics", None)
    if intr is None and hasattr(depth_frame, "profile"):
        intr = depth_frame.profile.as_video_stream_profile().intrinsics
    if intr is None:
        intr = getattr(pipeline, "intrinsics", None)
    if intr is None:
        raise RuntimeError("depth frame carries no intrinsics")
    return Intrinsics(intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy)


def _frame_units(depth_frame, pipeline):
    units = depth_frame.get_units() if hasattr(depth_frame, "get_units") else None
    if not units:
        units = getattr(pipeline, "depth_scale", None) or 0.001
    return float(units)


class PointCloudCapture:
    def __init__(self, device="CUDA:0", voxel_size=0.01):
        self.device = parse_device(device)
        self.voxel_size = voxel_size

    def capture_point_cloud(self, pipeline):
        """One frame -> down-sampled legacy-style cloud (float64 points + colours) or None when a frame is missing."""
        frames = pipeline.wait_for_frames()
        depth_frame = frames.get_depth_frame()
        color_frame = frames.get_color_frame()
        if not depth_frame or not color_frame:
            return None
        depth_image = np.asanyarray(depth_frame.get_data())
        color_image = np.asanyarray(color_frame.get_data())
        intr = _frame_intrinsics(depth_frame, pipeline)
        scale = _frame_units(depth_frame, pipeline)
        # raw BGR raster as colours, exactly like the reference (no depth/colour alignment, pointcloud_capture.py:36-39)
        vtx, col = ops.deproject_z16(depth_image, intr.fx, intr.fy, intr.ppx, intr.ppy, scale, color_bgr=color_image.reshape(depth_image.shape + (3,)),
                                     device=self.device, as_tensor=True)
        r = ops.voxel_down_sample_tensor(vtx, self.voxel_size, attr=col, device=self.device)
        pcd = PointCloud(device=self.device)
        pcd.points = r["points"].astype(np.float64)
        pcd.colors = r["attr"].astype(np.float64)
        return pcd
