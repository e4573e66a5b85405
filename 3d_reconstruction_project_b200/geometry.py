"""Duck-typed stand-ins for the Open3D objects the reference's hot path exchanges.

``PointCloud`` mirrors ``o3d.geometry.PointCloud`` (legacy, float64): ``points`` / ``colors`` / ``normals`` are
assignable, ``len()``-able and ``np.asarray()``-able; the methods the reference calls on the path
(``voxel_down_sample``, ``estimate_normals``, ``remove_statistical_outlier``, ``remove_radius_outlier``,
``select_by_index``, ``transform``, ``+=``) run the CUDA kernels of libb200recon.so. Open3D itself is not installed
in this image; real ``o3d.geometry.PointCloud`` objects are accepted wherever a cloud is taken (``as_cloud``).
"""
import copy

import numpy as np

from . import ops


class Vector3dVector:
    """o3d.utility.Vector3dVector: a float64 [N,3] array behind the sequence protocol (pointcloud_capture.py:43-44)."""

    __slots__ = ("_a",)

    def __init__(self, data=None):
        if data is None:
            a = np.zeros((0, 3), np.float64)
        elif isinstance(data, Vector3dVector):
            a = data._a.copy()
        else:
            a = np.array(data, dtype=np.float64, order="C", copy=True)
            if a.size == 0:
                a = a.reshape(0, 3)
            if a.ndim != 2 or a.shape[1] != 3:
                raise RuntimeError(f"Vector3dVector expects an [N,3] array, got shape {a.shape}")
        self._a = a

    def __len__(self):
        return self._a.shape[0]

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __getitem__(self, i):
        return self._a[i]

    def __iter__(self):
        return iter(self._a)

    def __repr__(self):
        return f"Vector3dVector with {len(self)} elements."


class KDTreeSearchParamHybrid:
    def __init__(self, radius, max_nn):
        self.radius, self.max_nn = float(radius), int(max_nn)


class KDTreeSearchParamKNN:
    def __init__(self, knn=30):
        self.knn = int(knn)


class KDTreeSearchParamRadius:
    def __init__(self, radius):
        self.radius = float(radius)


def _vec(v):
    return v if isinstance(v, Vector3dVector) else Vector3dVector(v)


class PointCloud:
    """Legacy-style point cloud (float64 host arrays); all geometry work is done on the GPU."""

    def __init__(self, points=None, device=0):
        self._points = _vec(points)
        self._colors = Vector3dVector()
        self._normals = Vector3dVector()
        self._covariances = None  # [N,3,3] float64 (GICP), or None
        self.device = device

    # attribute protocol used by main.py:39-49
    points = property(lambda s: s._points, lambda s, v: setattr(s, "_points", _vec(v)))
    colors = property(lambda s: s._colors, lambda s, v: setattr(s, "_colors", _vec(v)))
    normals = property(lambda s: s._normals, lambda s, v: setattr(s, "_normals", _vec(v)))

    @property
    def covariances(self):
        return self._covariances

    @covariances.setter
    def covariances(self, v):
        self._covariances = None if v is None else np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3, 3)

    def has_points(self):
        return len(self._points) > 0

    def has_colors(self):
        return len(self._points) > 0 and len(self._colors) == len(self._points)

    def has_normals(self):
        return len(self._points) > 0 and len(self._normals) == len(self._points)

    def has_covariances(self):
        return self._covariances is not None and len(self._points) > 0 and len(self._covariances) == len(self._points)

    def is_empty(self):
        return not self.has_points()

    def __repr__(self):
        return f"PointCloud with {len(self._points)} points."

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        c = PointCloud(device=self.device)
        c._points, c._colors, c._normals = Vector3dVector(self._points), Vector3dVector(self._colors), Vector3dVector(self._normals)
        c._covariances = None if self._covariances is None else self._covariances.copy()
        return c

    # ---- operators -----------------------------------------------------------------------------------------
    def __iadd__(self, other):
        """combined_pcd += aligned (main.py:49): Open3D keeps an attribute only if both sides carry it (or self is empty)."""
        other = as_cloud(other)
        n_self = len(self._points)
        keep_colors = (n_self == 0 or self.has_colors()) and other.has_colors()
        keep_normals = (n_self == 0 or self.has_normals()) and other.has_normals()
        keep_cov = (n_self == 0 or self.has_covariances()) and other.has_covariances()
        cat = lambda a, b: np.concatenate([np.asarray(a).reshape(-1, 3), np.asarray(b).reshape(-1, 3)], axis=0)
        new_colors = cat(self._colors, other._colors) if keep_colors else None
        new_normals = cat(self._normals, other._normals) if keep_normals else None
        new_cov = (np.concatenate([self._covariances if n_self else np.zeros((0, 3, 3)), other._covariances], axis=0) if keep_cov else None)
        self._points = Vector3dVector(cat(self._points, other._points))
        self._colors = Vector3dVector(new_colors)
        self._normals = Vector3dVector(new_normals)
        self._covariances = new_cov
        return self

    def __add__(self, other):
        c = self.clone()
        c += other
        return c

    # ---- K2 ------------------------------------------------------------------------------------------------
    def voxel_down_sample(self, voxel_size):
        """pointcloud_alignment.py:22-23 (legacy semantics). Output order: ascending voxel index."""
        out = PointCloud(device=self.device)
        if voxel_size <= 0:
            raise RuntimeError("voxel_size <= 0.")
        if not self.has_points():
            return out
        r = ops.voxel_down_sample_legacy(np.asarray(self._points), voxel_size, colors=np.asarray(self._colors) if self.has_colors() else None,
                                         normals=np.asarray(self._normals) if self.has_normals() else None, device=self.device)
        out._points = Vector3dVector(r["points"])
        if r["colors"] is not None:
            out._colors = Vector3dVector(r["colors"])
        if r["normals"] is not None:
            out._normals = Vector3dVector(r["normals"])
        return out

    # ---- K3 ------------------------------------------------------------------------------------------------
    def estimate_normals(self, search_param=None, fast_normal_computation=True):
        """pointcloud_alignment.py:27-28, test/GICP1.py:77. In place; existing normals only steer the sign."""
        if search_param is None:
            search_param = KDTreeSearchParamKNN(30)
        if isinstance(search_param, KDTreeSearchParamHybrid):
            k, r = search_param.max_nn, search_param.radius
        elif isinstance(search_param, KDTreeSearchParamKNN):
            k, r = search_param.knn, 0.0
        else:
            raise RuntimeError("estimate_normals: only KDTreeSearchParamHybrid / KDTreeSearchParamKNN are supported on this path")
        if not self.has_points():
            return self
        prior = np.asarray(self._normals) if self.has_normals() else None
        self._normals = Vector3dVector(ops.estimate_normals_legacy(np.asarray(self._points), k, r, prior=prior, device=self.device))
        return self

    def orient_normals_consistent_tangent_plane(self, k, lambda_penalty=0.0, cos_alpha_tol=1.0):
        """normal_estimation.py:21 (k = 100). In place. Only the published algorithm (lambda = 0, cos_alpha_tol = 1) is built."""
        if lambda_penalty != 0.0 or cos_alpha_tol != 1.0:
            raise RuntimeError("orient_normals_consistent_tangent_plane: lambda / cos_alpha_tol variants are not supported on this path")
        if not self.has_normals():
            raise RuntimeError("No normals in the PointCloud. Call EstimateNormals() first.")
        nrm, _ = ops.orient_normals_consistent_tangent_plane(np.asarray(self._points), np.asarray(self._normals), int(k), device=self.device)
        self._normals = Vector3dVector(nrm)
        return self

    def estimate_covariances_from_normals(self, eps=1e-3):
        """What registration_generalized_icp does first (test/GICP1.py:99-102): normals (KNN 20 if absent) -> C = R diag(eps,1,1) R^T."""
        if not self.has_normals():
            self.estimate_normals(KDTreeSearchParamKNN(20))
        self._covariances = ops.covariances_from_normals(np.asarray(self._normals), eps, device=self.device)
        return self

    # ---- outlier filters -----------------------------------------------------------------------------------
    def select_by_index(self, indices, invert=False):
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        n = len(self._points)
        if invert:
            mask = np.ones(n, bool)
            mask[idx] = False
            idx = np.nonzero(mask)[0]
        out = PointCloud(device=self.device)
        if len(idx) == 0 or n == 0:
            return out
        out._points = Vector3dVector(ops.select_rows(np.asarray(self._points), idx, device=self.device))
        if self.has_colors():
            out._colors = Vector3dVector(ops.select_rows(np.asarray(self._colors), idx, device=self.device))
        if self.has_normals():
            out._normals = Vector3dVector(ops.select_rows(np.asarray(self._normals), idx, device=self.device))
        if self.has_covariances():
            out._covariances = ops.select_rows(self._covariances.reshape(n, 9), idx, device=self.device).reshape(-1, 3, 3)
        return out

    def remove_statistical_outlier(self, nb_neighbors, std_ratio, print_progress=False):
        """pointcloud_processing.py:35-36 -> (cloud, ascending index list)"""
        if nb_neighbors < 1 or std_ratio <= 0:
            raise RuntimeError("Illegal input parameters, the number of neighbors and standard deviation ratio must be positive.")
        if not self.has_points():
            return PointCloud(device=self.device), []
        _, idx = ops.remove_statistical_outlier(np.asarray(self._points), nb_neighbors, std_ratio, device=self.device)
        return self.select_by_index(idx), idx.tolist()

    def remove_radius_outlier(self, nb_points, radius, print_progress=False):
        """pointcloud_processing.py:39 -> (cloud, ascending index list)"""
        if nb_points < 1 or radius <= 0:
            raise RuntimeError("Illegal input parameters, number of points and radius must be positive.")
        if not self.has_points():
            return PointCloud(device=self.device), []
        _, idx = ops.remove_radius_outlier(np.asarray(self._points), nb_points, radius, device=self.device)
        return self.select_by_index(idx), idx.tolist()

    # ---- a12 -----------------------------------------------------------------------------------------------
    def transform(self, T):
        """pointcloud_alignment.py:42; in place, returns self."""
        if not self.has_points():
            return self
        p, n, c = ops.transform(T, np.asarray(self._points), np.asarray(self._normals) if self.has_normals() else None,
                                self._covariances.reshape(-1, 9) if self.has_covariances() else None, device=self.device)
        self._points = Vector3dVector(p)
        if n is not None:
            self._normals = Vector3dVector(n)
        if c is not None:
            self._covariances = c.reshape(-1, 3, 3)
        return self

    def get_min_bound(self):
        return np.asarray(self._points).min(axis=0) if self.has_points() else np.zeros(3)

    def get_max_bound(self):
        return np.asarray(self._points).max(axis=0) if self.has_points() else np.zeros(3)


def as_cloud(obj, device=0):
    """Accepts our PointCloud or anything with Open3D's attribute protocol (a real o3d.geometry.PointCloud)."""
    if isinstance(obj, PointCloud):
        return obj
    if hasattr(obj, "points"):
        c = PointCloud(np.asarray(obj.points), device=device)
        if hasattr(obj, "colors") and len(obj.colors) == len(c.points):
            c.colors = np.asarray(obj.colors)
        if hasattr(obj, "normals") and len(obj.normals) == len(c.points):
            c.normals = np.asarray(obj.normals)
        return c
    raise TypeError(f"expected a PointCloud-like object, got {type(obj).__name__}")
