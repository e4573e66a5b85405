"""Fake `pycuda` (the reference only uses mem_alloc + memcpy_htod, realsense_pipeline.py:66-69)."""
