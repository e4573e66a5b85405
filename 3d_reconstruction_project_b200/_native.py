"""ctypes binding of libb200recon.so (the C ABI declared in include/b200recon.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the first call raises.
The library is built in-tree by ``__graft_entry__.build()`` (``make -C 3d_reconstruction_project_b200/csrc``).
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# B3D_LIB: alternative build of the same library (A/B experiments); default is the in-tree build
LIB_PATH = os.environ.get("B3D_LIB") or os.path.join(_HERE, "libb200recon.so")

OK = 0
E_INVALID, E_CUDA, E_RANGE, E_NOMEM, E_STATE = -1, -2, -3, -4, -5
ICP_POINT_TO_POINT, ICP_POINT_TO_PLANE, ICP_GENERALIZED = 0, 1, 2


class NativeError(RuntimeError):
    """Raised for every non-zero return code of the C ABI (Open3D raises RuntimeError for the same conditions)."""

    def __init__(self, code, text):
        super().__init__(text)
        self.code = code


class IcpResult(C.Structure):
    _fields_ = [("transformation", C.c_double * 16), ("fitness", C.c_double), ("inlier_rmse", C.c_double),
                ("iterations", C.c_int32), ("converged", C.c_int32), ("n_correspondences", C.c_int64)]


class RansacResult(C.Structure):
    _fields_ = [("transformation", C.c_double * 16), ("fitness", C.c_double), ("inlier_rmse", C.c_double),
                ("n_correspondences", C.c_int64), ("iterations", C.c_int64), ("validated", C.c_int64)]


class FgrOption(C.Structure):
    _fields_ = [("division_factor", C.c_double), ("use_absolute_scale", C.c_int), ("decrease_mu", C.c_int),
                ("maximum_correspondence_distance", C.c_double), ("iteration_number", C.c_int), ("tuple_scale", C.c_double),
                ("maximum_tuple_count", C.c_int), ("tuple_test", C.c_int)]


class PairParams(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("fx", C.c_float), ("fy", C.c_float), ("ppx", C.c_float), ("ppy", C.c_float),
                ("depth_scale", C.c_float), ("voxel_size", C.c_float), ("normals_max_nn", C.c_int), ("normals_radius", C.c_double),
                ("icp_kind", C.c_int), ("icp_max_dist", C.c_double), ("icp_rel_fitness", C.c_double), ("icp_rel_rmse", C.c_double),
                ("icp_max_iter", C.c_int)]


class DisparityParams(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("Q", C.c_double * 16), ("min_disp16", C.c_int), ("voxel_size", C.c_float),
                ("normals_max_nn", C.c_int), ("normals_radius", C.c_double), ("icp_kind", C.c_int), ("icp_max_dist", C.c_double),
                ("icp_rel_fitness", C.c_double), ("icp_rel_rmse", C.c_double), ("icp_max_iter", C.c_int)]


class PairResult(C.Structure):
    _fields_ = [("icp", IcpResult), ("n_raw", C.c_int64), ("m_source", C.c_int64), ("m_target", C.c_int64)]


_vp, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/b200recon.h declares
SIGNATURES = {
    "b3d_version": (_i, []),
    "b3d_last_error": (C.c_char_p, []),
    "b3d_ctx_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "b3d_ctx_destroy": (_i, [_vp]),
    "b3d_ctx_synchronize": (_i, [_vp]),
    "b3d_ctx_launch_count": (_i64, [_vp]),
    "b3d_ctx_profile": (_i, [_vp, _i]),
    "b3d_ctx_profile_report": (_i64, [_vp, C.c_char_p, _i64]),
    "b3d_deproject_z16": (_i, [_vp, _vp, _i, _i, _f, _f, _f, _f, _f, _vp]),
    "b3d_deproject_z16_color": (_i, [_vp, _vp, _vp, _i, _i, _f, _f, _f, _f, _f, _vp, _vp]),
    "b3d_deproject_rgbd": (_i, [_vp, _vp, _vp, _i, _i, _d, _d, _d, _d, _f, _f, _i, _vp, _vp, _pi64]),
    "b3d_reproject_disparity": (_i, [_vp, _vp, _i, _i, C.POINTER(_d), _vp]),
    "b3d_reproject_disparity_valid": (_i, [_vp, _vp, _i, _i, C.POINTER(_d), _i, _vp, _pi64]),
    "b3d_voxel_downsample_legacy": (_i, [_vp, _vp, _vp, _vp, _i64, _d, _vp, _vp, _vp, _vp, _vp, _pi64]),
    "b3d_voxel_downsample_tensor": (_i, [_vp, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _pi64]),
    "b3d_grid_build": (_i, [_vp, _vp, _i64, _i, _d, _i, _d, C.POINTER(_vp)]),
    "b3d_grid_destroy": (_i, [_vp, _vp]),
    "b3d_grid_info": (_i, [_vp, _pi64, _pi64, C.POINTER(_d)]),
    "b3d_knn_hybrid": (_i, [_vp, _vp, _vp, _i64, _i, _d, _vp, _vp, _vp]),
    "b3d_estimate_normals_legacy": (_i, [_vp, _vp, _i64, _i, _d, _vp, _vp]),
    "b3d_estimate_normals_tensor": (_i, [_vp, _vp, _i64, _i, _f, _vp]),
    "b3d_covariances_from_normals": (_i, [_vp, _vp, _i64, _d, _vp]),
    "b3d_compute_fpfh": (_i, [_vp, _vp, _vp, _i64, _i, _d, _vp]),
    "b3d_orient_normals_consistent_tangent_plane": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "b3d_match_features": (_i, [_vp, _vp, _i64, _vp, _i64, _i, _vp]),
    "b3d_fgr_feature_matching": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _i, _vp, C.c_uint64, _vp, _vp]),
    "b3d_ransac_correspondence": (_i, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _d, _i, _d, _d, _i64, _d, C.c_uint64, _vp]),
    "b3d_statistical_outlier": (_i, [_vp, _vp, _i64, _i, _d, _vp, _vp, _pi64]),
    "b3d_radius_outlier": (_i, [_vp, _vp, _i64, _i, _d, _vp, _vp, _pi64]),
    "b3d_gather_rows_f64": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "b3d_transform_f64": (_i, [_vp, C.POINTER(_d), _vp, _i64, _vp, _vp]),
    "b3d_icp_correspondences": (_i, [_vp, _vp, _i64, _vp, _i64, C.POINTER(_d), _d, _vp, C.POINTER(_d)]),
    "b3d_information_matrix": (_i, [_vp, _vp, _i64, _vp, _i64, C.POINTER(_d), _d, C.POINTER(_d)]),
    "b3d_icp": (_i, [_vp, _i, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _d, C.POINTER(_d), _d, _d, _i, C.POINTER(IcpResult), _vp]),
    "b3d_icp_batch": (_i, [_vp, _i, _i, _vp, _pi64, _vp, _vp, _pi64, _vp, _vp, _d, C.POINTER(_d), _d, _d, _i, C.POINTER(IcpResult), _vp]),
    "b3d_icp_begin": (_i, [_vp, _i, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _d, C.POINTER(_d), _d, _d, _i, C.POINTER(_vp)]),
    "b3d_icp_accumulate": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "b3d_icp_update": (_i, [_vp, _vp, C.POINTER(_i)]),
    "b3d_icp_set_peers": (_i, [_vp, _vp, _i, _i, C.POINTER(_vp)]),
    "b3d_icp_pass_peers": (_i, [_vp, _vp, C.POINTER(_i)]),
    "b3d_icp_finish": (_i, [_vp, _vp, C.POINTER(IcpResult), _vp]),
    "b3d_register_depth_pair": (_i, [_vp, C.POINTER(PairParams), _vp, _vp, _i, C.POINTER(PairResult)]),
    "b3d_register_depth_pairs": (_i, [_vp, C.POINTER(PairParams), _vp, _vp, _i, _i, C.POINTER(PairResult)]),
    "b3d_register_disparity_pairs": (_i, [_vp, C.POINTER(DisparityParams), _vp, _vp, _i, _i, C.POINTER(PairResult)]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Loads libb200recon.so once. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                      "(make -C 3d_reconstruction_project_b200/csrc). There is no CPU fallback.")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def check(rc):
    if rc != OK:
        text = lib().b3d_last_error()
        raise NativeError(rc, (text or b"").decode("utf-8", "replace") or f"libb200recon error {rc}")
    return rc
