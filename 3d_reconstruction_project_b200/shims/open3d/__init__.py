"""Minimal `open3d` namespace over b200recon: exactly the attributes the reference's main.py / hot-path modules touch."""
import types

from b200recon import geometry as _g, plyio as _io, registration as _reg

__version__ = "0.18.0-b200recon-shim"

geometry = types.SimpleNamespace(PointCloud=_g.PointCloud, KDTreeSearchParamHybrid=_g.KDTreeSearchParamHybrid,
                                 KDTreeSearchParamKNN=_g.KDTreeSearchParamKNN, KDTreeSearchParamRadius=_g.KDTreeSearchParamRadius)
utility = types.SimpleNamespace(Vector3dVector=_g.Vector3dVector)
io = types.SimpleNamespace(read_point_cloud=_io.read_point_cloud, write_point_cloud=_io.write_point_cloud)
pipelines = types.SimpleNamespace(registration=_reg)


class _Device:
    def __init__(self, name="CUDA:0"):
        self.name = str(name)

    def __repr__(self):
        return self.name


def _cuda_available():
    import torch
    return torch.cuda.is_available()


core = types.SimpleNamespace(Device=_Device, Dtype=types.SimpleNamespace(Float32="Float32", Float64="Float64"),
                             cuda=types.SimpleNamespace(is_available=_cuda_available))
