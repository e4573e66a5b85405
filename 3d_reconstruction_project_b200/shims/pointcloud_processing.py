"""Shim: the reference's `pointcloud_processing` module name resolving to b200recon's class (same name, same signature)."""
from b200recon.pointcloud_processing import PointCloudProcessingWithCUDA  # noqa: F401
