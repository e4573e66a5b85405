"""Per-(device, thread) native contexts. PyTorch supplies the device memory and the stream; the kernels are ours."""
import ctypes as C
import threading

import numpy as np
import torch

from . import _native as N

_tls = threading.local()


class Context:
    """Owns a b3d_ctx bound to ``torch.cuda.current_stream(device)``; not thread-safe (one per thread, like the C ABI)."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("b200recon needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", int(device))
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream(self.device)
            h = C.c_void_p()
            N.check(N.lib().b3d_ctx_create(self.device.index, C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            N.lib().b3d_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        N.check(N.lib().b3d_ctx_synchronize(self.handle))

    @property
    def launches(self):
        return int(N.lib().b3d_ctx_launch_count(self.handle))

    def profile(self, enable=True):
        """Per-kernel CUDA-event timing on/off (clears earlier records)."""
        N.check(N.lib().b3d_ctx_profile(self.handle, int(bool(enable))))

    def profile_report(self):
        """{kernel name: (launches, total_ms, declared bytes)} since profiling was enabled; clears the records."""
        need = N.lib().b3d_ctx_profile_report(self.handle, None, 0)
        if need < 0:
            raise RuntimeError("profile report failed")
        buf = C.create_string_buffer(int(need) + 16)
        N.lib().b3d_ctx_profile_report(self.handle, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            f = line.split("\t")
            out[f[0]] = (int(f[1]), float(f[2]), int(f[3]) if len(f) > 3 else 0)
        return out

    # ---- torch-backed buffers ----------------------------------------------------------------------------------
    def to_device(self, a, dtype=None):
        """numpy / torch -> contiguous CUDA tensor on this context's device (no copy if already there)."""
        if isinstance(a, torch.Tensor):
            t = a.to(self.device)
            if dtype is not None:
                t = t.to(dtype)
            return t.contiguous()
        arr = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=_np_dtype(dtype)))
        return torch.from_numpy(arr).to(self.device, non_blocking=False)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)


def _np_dtype(dt):
    return {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64, torch.uint8: np.uint8,
            torch.int16: np.int16, torch.uint16: np.uint16}.get(dt, dt)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def get_context(device=0):
    """The calling thread's context for ``device`` (the reference calls from a scan thread, main.py:56-61)."""
    d = int(torch.device(device).index or 0) if not isinstance(device, int) else device
    pool = getattr(_tls, "pool", None)
    if pool is None:
        pool = _tls.pool = {}
    if not torch.cuda.is_available():
        raise RuntimeError("b200recon needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    # one context per (device, current torch stream): a caller inside `with torch.cuda.stream(s)` gets a context whose
    # kernels run on s, like the torch allocations and copies around the call
    key = (d, int(torch.cuda.current_stream(d).cuda_stream))
    ctx = pool.get(key)
    if ctx is None or ctx.handle is None:
        ctx = pool[key] = Context(d)
    return ctx


def parse_device(device):
    """'CUDA:0' (Open3D spelling, pointcloud_capture.py:14), 'cuda:1', 1, torch.device -> CUDA index."""
    if isinstance(device, int):
        return device
    if isinstance(device, torch.device):
        return device.index or 0
    s = str(device).strip().lower()
    if s.startswith("cuda"):
        return int(s.split(":")[1]) if ":" in s else 0
    if s.startswith("cpu"):
        raise RuntimeError("b200recon has no CPU path; pass device='CUDA:0'")
    return int(s)
