"""RealSensePipeline -- same class / method names as the reference's realsense_pipeline.py:6-85.

The camera itself (librealsense over USB) is outside the hot path. ``start_pipeline`` uses ``pyrealsense2`` when it is
importable; otherwise a frame source must be injected (``RealSensePipeline(source=ReplayPipeline(...))``), e.g. the
replay of recorded / synthetic depth+colour rasters used by the tests and benchmarks. ``transfer_to_cuda`` is the
reference's host->device staging (realsense_pipeline.py:58-71, pycuda mem_alloc + memcpy_htod): here a pinned-memory
upload into a torch CUDA tensor, which is what the deprojection kernel consumes.
"""
import numpy as np
import torch


class Intrinsics:
    """rs.intrinsics fields used by rs2_deproject_pixel_to_point."""

    def __init__(self, width, height, fx, fy, ppx, ppy):
        self.width, self.height, self.fx, self.fy, self.ppx, self.ppy = int(width), int(height), float(fx), float(fy), float(ppx), float(ppy)


class _ReplayFrame:
    def __init__(self, data, intrinsics=None, units=None):
        self._data, self.intrinsics, self._units = data, intrinsics, units

    def get_data(self):
        return self._data

    def get_units(self):
        return self._units

    def __bool__(self):
        return self._data is not None


class _ReplayFrameset:
    def __init__(self, depth, color):
        self._d, self._c = depth, color

    def get_depth_frame(self):
        return self._d

    def get_color_frame(self):
        return self._c


class ReplayPipeline:
    """Stands in for rs.pipeline(): ``wait_for_frames()`` yields (depth u16 [H,W], colour u8 [H,W,3] BGR) pairs from a list
    or iterator. ``loop=False`` raises RuntimeError("Frame didn't arrive within 5000") at the end, like a stalled camera."""

    def __init__(self, frames, intrinsics, depth_scale=0.001, loop=False):
        self._frames = list(frames)
        self._i = 0
        self.intrinsics = intrinsics
        self.depth_scale = float(depth_scale)
        self.loop = loop
        self.started = False

    def start(self, config=None):
        self.started = True
        return self

    def stop(self):
        self.started = False

    def wait_for_frames(self, timeout_ms=5000):
        if self._i >= len(self._frames):
            if not self.loop or not self._frames:
                raise RuntimeError(f"Frame didn't arrive within {timeout_ms}")
            self._i = 0
        depth, color = self._frames[self._i]
        self._i += 1
        d = _ReplayFrame(None if depth is None else np.ascontiguousarray(depth, dtype=np.uint16), self.intrinsics, self.depth_scale)
        c = _ReplayFrame(None if color is None else np.ascontiguousarray(color, dtype=np.uint8))
        return _ReplayFrameset(d, c)


class RealSensePipeline:
    def __init__(self, source=None):
        """Initializes the RealSense camera pipeline (realsense_pipeline.py:7-13). ``source``: optional frame source used
        instead of a physical camera."""
        self.pipeline = None
        self.depth_frame = None
        self.color_frame = None
        self._source = source

    def start_pipeline(self):
        """Starts the colour + depth streams, z16 + bgr8 640x480 @ 15 fps (realsense_pipeline.py:15-31)."""
        if self._source is not None:
            self.pipeline = self._source
            self.pipeline.start(None)
            return
        try:
            import pyrealsense2.pyrealsense2 as rs
        except ImportError:
            try:
                import pyrealsense2 as rs
            except ImportError as e:
                raise RuntimeError("pyrealsense2 is not installed and no frame source was injected: "
                                   "use RealSensePipeline(source=ReplayPipeline(...))") from e
        self.pipeline = rs.pipeline()
        config = rs.config()
        config.enable_stream(rs.stream.depth, 640, 480, rs.format.z16, 15)
        config.enable_stream(rs.stream.color, 640, 480, rs.format.bgr8, 15)
        try:
            self.pipeline.start(config)
        except RuntimeError as e:
            # The reference prints the error, asks the pipeline that just failed to start for its active profile (which raises
            # again), and exit(1)s the whole process from inside this class (realsense_pipeline.py:24-31). Same message, but the
            # caller gets an exception it can handle and no call is made on a pipeline that never started.
            print(f"Failed to start pipeline: {e}")
            self.pipeline = None
            raise RuntimeError(f"Failed to start pipeline: {e}") from e

    def stop_pipeline(self):
        self.pipeline.stop()

    def get_frames(self):
        """One frame of depth + colour as numpy arrays (realsense_pipeline.py:39-56)."""
        frames = self.pipeline.wait_for_frames()
        depth_frame = frames.get_depth_frame()
        color_frame = frames.get_color_frame()
        if not depth_frame or not color_frame:
            raise RuntimeError("Failed to capture frames")
        self.depth_frame, self.color_frame = depth_frame, color_frame
        return np.asanyarray(depth_frame.get_data()), np.asanyarray(color_frame.get_data())

    def transfer_to_cuda(self, np_array, device=0):
        """Host -> device copy through pinned memory (realsense_pipeline.py:58-71). Returns a CUDA torch tensor."""
        if not torch.cuda.is_available():
            raise RuntimeError("transfer_to_cuda needs a CUDA device; there is no CPU fallback")
        a = np.ascontiguousarray(np_array)
        if a.dtype == np.uint16:
            a = a.view(np.int16)  # same bits; torch's uint16 support is partial
        host = torch.from_numpy(a).pin_memory()
        return host.to(torch.device("cuda", device), non_blocking=True)

    def process_frames_with_cuda(self):
        depth_image, color_image = self.get_frames()
        depth_cuda = self.transfer_to_cuda(depth_image)
        color_cuda = self.transfer_to_cuda(color_image)
        print("Depth and color images have been transferred to CUDA.")
        return depth_cuda, color_cuda
