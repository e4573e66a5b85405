"""Fake `pyrealsense2`: `pipeline()` replays the frames of the .npz named by $B3D_REPLAY (see shims/README.md)."""
import os
import time
import types

import numpy as np

from b200recon.realsense_pipeline import Intrinsics, ReplayPipeline, _ReplayFrame, _ReplayFrameset

stream = types.SimpleNamespace(depth="depth", color="color")
format = types.SimpleNamespace(z16="z16", bgr8="bgr8")


class config:
    def enable_stream(self, *args):
        pass


class _EndlessReplay(ReplayPipeline):
    """After the recorded frames: empty framesets (the reference's loop prints "No valid point cloud captured"), one every
    50 ms, for as long as the caller keeps polling -- nothing is stored per poll."""

    def wait_for_frames(self, timeout_ms=5000):
        if self._i >= len(self._frames):
            time.sleep(0.05)
            return _ReplayFrameset(_ReplayFrame(None, self.intrinsics, self.depth_scale), _ReplayFrame(None))
        return super().wait_for_frames(timeout_ms)


def pipeline():
    path = os.environ.get("B3D_REPLAY")
    if not path:
        raise RuntimeError("fake pyrealsense2: set B3D_REPLAY to an .npz with depth / color / intrinsics / depth_scale")
    d = np.load(path)
    fx, fy, ppx, ppy = [float(v) for v in d["intrinsics"]]
    h, w = d["depth"].shape[1:]
    frames = [(d["depth"][i], d["color"][i]) for i in range(len(d["depth"]))]
    return _EndlessReplay(frames, Intrinsics(w, h, fx, fy, ppx, ppy), float(d["depth_scale"]))


class pointcloud:
    """rs.pointcloud(): the deprojection itself lives in b200recon.PointCloudCapture (b3d_deproject_z16)."""
