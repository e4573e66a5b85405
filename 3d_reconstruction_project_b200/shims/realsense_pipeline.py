"""Shim: the reference's `realsense_pipeline` module name resolving to b200recon's class (same name, same signature)."""
from b200recon.realsense_pipeline import RealSensePipeline  # noqa: F401
