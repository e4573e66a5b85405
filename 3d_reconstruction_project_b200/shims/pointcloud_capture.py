"""Shim: the reference's `pointcloud_capture` module name resolving to b200recon's class (same name, same signature)."""
from b200recon.pointcloud_capture import PointCloudCapture  # noqa: F401
