"""Shim: the reference's `normal_estimation` module name resolving to b200recon's class (same name, same signature)."""
from b200recon.normal_estimation import NormalEstimation  # noqa: F401
