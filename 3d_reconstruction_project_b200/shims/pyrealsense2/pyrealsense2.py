"""`import pyrealsense2.pyrealsense2 as rs` (the source-build layout the reference uses, realsense_pipeline.py:1)."""
from . import config, format, pipeline, pointcloud, stream  # noqa: F401
