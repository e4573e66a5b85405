"""Shim: the reference's `pointcloud_alignment` module name resolving to b200recon's class (same name, same signature)."""
from b200recon.pointcloud_alignment import PointCloudAlignment  # noqa: F401
