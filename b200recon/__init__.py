"""Importable alias of the product package (its directory name, 3d_reconstruction_project_b200, starts with a digit)."""
import importlib
import sys

_pkg = importlib.import_module("3d_reconstruction_project_b200")
sys.modules[__name__] = _pkg
for _name in ("ops", "geometry", "registration", "plyio", "context", "_native", "realsense_pipeline", "pointcloud_capture",
              "pointcloud_alignment", "pointcloud_processing", "normal_estimation", "distributed"):
    try:
        sys.modules[__name__ + "." + _name] = importlib.import_module("3d_reconstruction_project_b200." + _name)
    except ImportError:
        pass
