"""CPU oracle bindings (ctypes over oracle/libb3d_oracle.so). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. The product package (3d_reconstruction_project_b200 / b200recon) never does.
Every wrapper cites the reference call site it restates in oracle/b3d_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb3d_oracle.so")
_SRC = os.path.join(_HERE, "b3d_oracle.cpp")


def build(force=False):
    """Compile the oracle with g++ (seconds). Rebuilds when the source is newer than the .so."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libb3d_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_deproject_rgbd.restype = C.c_int64
        _lib.orc_voxel_legacy.restype = C.c_int64
        _lib.orc_voxel_tensor.restype = C.c_int64
        _lib.orc_statistical_outlier.restype = C.c_int64
        _lib.orc_radius_outlier.restype = C.c_int64
        _lib.orc_correspondences.restype = C.c_int64
        _lib.orc_icp.restype = C.c_int
        _lib.orc_information_matrix.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def deproject_z16(depth, fx, fy, ppx, ppy, depth_scale=0.001):
    depth = _c(depth, np.uint16)
    h, w = depth.shape
    out = np.empty((h * w, 3), np.float32)
    lib().orc_deproject_z16(_p(depth), w, h, C.c_float(fx), C.c_float(fy), C.c_float(ppx), C.c_float(ppy), C.c_float(depth_scale), _p(out))
    return out


def deproject_rgbd(depth, color, fx, fy, cx, cy, depth_scale=1000.0, depth_trunc=3.0, flip=True):
    depth = _c(depth, np.uint16)
    color = _c(color, np.uint8)
    h, w = depth.shape
    xyz = np.empty((h * w, 3), np.float64)
    rgb = np.empty((h * w, 3), np.float64) if color is not None else None
    n = lib().orc_deproject_rgbd(_p(depth), _p(color), w, h, C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                                 C.c_float(np.float32(depth_scale)), C.c_float(np.float32(depth_trunc)), int(bool(flip)), _p(xyz), _p(rgb))
    return xyz[:n].copy(), (rgb[:n].copy() if rgb is not None else None)


def reproject_disparity(disp16, Q):
    disp16 = _c(disp16, np.int16)
    Q = _c(Q, np.float64)
    h, w = disp16.shape
    out = np.empty((h, w, 3), np.float32)
    lib().orc_reproject_disparity(_p(disp16), w, h, _p(Q), _p(out))
    return out


def voxel_legacy(xyz, voxel_size, colors=None, normals=None):
    xyz = _c(xyz, np.float64)
    colors = _c(colors, np.float64)
    normals = _c(normals, np.float64)
    n = len(xyz)
    o_xyz = np.empty((max(n, 1), 3), np.float64)
    o_col = np.empty((max(n, 1), 3), np.float64) if colors is not None else None
    o_nrm = np.empty((max(n, 1), 3), np.float64) if normals is not None else None
    o_idx = np.empty((max(n, 1), 3), np.int32)
    m = lib().orc_voxel_legacy(_p(xyz), _p(colors), _p(normals), C.c_int64(n), C.c_double(voxel_size), _p(o_xyz), _p(o_col), _p(o_nrm), _p(o_idx))
    if m == -1:
        raise RuntimeError("voxel_size <= 0.")
    if m == -2:
        raise RuntimeError("voxel_size is too small.")
    return dict(points=o_xyz[:m].copy(), colors=None if o_col is None else o_col[:m].copy(),
                normals=None if o_nrm is None else o_nrm[:m].copy(), index=o_idx[:m].copy())


def voxel_tensor(xyz, voxel_size, attr=None):
    xyz = _c(xyz, np.float32)
    attr = _c(attr, np.float32)
    n = len(xyz)
    o_xyz = np.empty((max(n, 1), 3), np.float32)
    o_att = np.empty((max(n, 1), 3), np.float32) if attr is not None else None
    o_idx = np.empty((max(n, 1), 3), np.int64)
    m = lib().orc_voxel_tensor(_p(xyz), _p(attr), C.c_int64(n), C.c_float(np.float32(voxel_size)), _p(o_xyz), _p(o_att), _p(o_idx))
    if m == -1:
        raise RuntimeError("voxel_size must be positive.")
    return dict(points=o_xyz[:m].copy(), attr=None if o_att is None else o_att[:m].copy(), index=o_idx[:m].copy())


def knn(points, queries, k, radius=0.0):
    f64 = points.dtype == np.float64
    dt = np.float64 if f64 else np.float32
    points = _c(points, dt)
    queries = _c(queries, dt)
    nq = len(queries)
    idx = np.empty((nq, k), np.int32)
    d2 = np.empty((nq, k), dt)
    cnt = np.empty(nq, np.int32)
    if f64:
        lib().orc_knn_f64(_p(points), C.c_int64(len(points)), _p(queries), C.c_int64(nq), int(k), C.c_double(radius), _p(idx), _p(d2), _p(cnt))
    else:
        lib().orc_knn_f32(_p(points), C.c_int64(len(points)), _p(queries), C.c_int64(nq), int(k), C.c_float(radius), _p(idx), _p(d2), _p(cnt))
    return idx, d2, cnt


def normals_legacy(points, max_nn, radius, prior=None):
    points = _c(points, np.float64)
    prior = _c(prior, np.float64)
    out = np.empty_like(points)
    lib().orc_normals_legacy(_p(points), C.c_int64(len(points)), int(max_nn), C.c_double(radius), _p(prior), _p(out))
    return out


def normals_tensor(points, max_nn, radius):
    points = _c(points, np.float32)
    out = np.empty_like(points)
    lib().orc_normals_tensor(_p(points), C.c_int64(len(points)), int(max_nn), C.c_float(radius), _p(out))
    return out


def statistical_outlier(points, nb_neighbors, std_ratio):
    points = _c(points, np.float64)
    keep = np.zeros(len(points), np.uint8)
    avg = np.empty(len(points), np.float64)
    r = lib().orc_statistical_outlier(_p(points), C.c_int64(len(points)), int(nb_neighbors), C.c_double(std_ratio), _p(keep), _p(avg))
    if r < 0:
        raise RuntimeError("Illegal input parameters, the number of neighbors and standard deviation ratio must be positive.")
    return keep.astype(bool), avg


def radius_outlier(points, nb_points, radius):
    points = _c(points, np.float64)
    keep = np.zeros(len(points), np.uint8)
    r = lib().orc_radius_outlier(_p(points), C.c_int64(len(points)), int(nb_points), C.c_double(radius), _p(keep))
    if r < 0:
        raise RuntimeError("Illegal input parameters, number of points and radius must be positive.")
    return keep.astype(bool)


def covariances_from_normals(normals, eps=1e-3):
    normals = _c(normals, np.float64)
    cov = np.empty((len(normals), 3, 3), np.float64)
    lib().orc_covariances_from_normals(_p(normals), C.c_int64(len(normals)), C.c_double(eps), _p(cov))
    return cov


def transform(T, points, normals=None, cov=None):
    T = _c(T, np.float64)
    points = np.array(points, dtype=np.float64, order="C", copy=True)
    normals = None if normals is None else np.array(normals, dtype=np.float64, order="C", copy=True)
    cov = None if cov is None else np.array(cov, dtype=np.float64, order="C", copy=True)
    lib().orc_transform(_p(T), _p(points), C.c_int64(len(points)), _p(normals), _p(cov))
    return points, normals, cov


def correspondences(src, tgt, T, dmax):
    src = _c(src, np.float64)
    tgt = _c(tgt, np.float64)
    T = _c(np.eye(4) if T is None else T, np.float64)
    corr = np.empty(len(src), np.int32)
    s = C.c_double(0)
    n = lib().orc_correspondences(_p(src), C.c_int64(len(src)), _p(tgt), C.c_int64(len(tgt)), _p(T), C.c_double(dmax), _p(corr), C.byref(s))
    return corr, int(n), s.value


def information_matrix(src, tgt, dmax, T=None):
    src = _c(src, np.float64)
    tgt = _c(tgt, np.float64)
    T = _c(np.eye(4) if T is None else T, np.float64)
    out = np.empty((6, 6), np.float64)
    lib().orc_information_matrix(_p(src), C.c_int64(len(src)), _p(tgt), C.c_int64(len(tgt)), _p(T), C.c_double(dmax), _p(out))
    return out


def fpfh(points, normals, max_nn, radius):
    points = _c(points, np.float64)
    normals = _c(normals, np.float64)
    out = np.empty((len(points), 33), np.float64)
    lib().orc_fpfh(_p(points), _p(normals), C.c_int64(len(points)), int(max_nn), C.c_double(radius), _p(out))
    return out


def orient_normals(points, normals, k=100):
    """PointCloud.orient_normals_consistent_tangent_plane(k) -- normal_estimation.py:21. -> (oriented normals, flipped mask)"""
    points = _c(points, np.float64)
    out = np.array(normals, dtype=np.float64, order="C", copy=True)
    flipped = np.zeros(len(points), np.uint8)
    rc = lib().orc_orient_normals(_p(points), _p(out), C.c_int64(len(points)), int(k), _p(flipped))
    if rc != 0:
        raise RuntimeError("Not enough points to create a tetrahedral mesh.")
    return out, flipped.astype(bool)


def match_features(fa, fb):
    """Nearest row of fb for every row of fa (float64 squared L2, ties to the smallest index)."""
    fa, fb = _c(fa, np.float64), _c(fb, np.float64)
    out = np.empty(len(fa), np.int32)
    lib().orc_match_features(_p(fa), C.c_int64(len(fa)), _p(fb), C.c_int64(len(fb)), int(fa.shape[1]), _p(out))
    return out


def ransac(src, tgt, corres, dmax, ransac_n=3, edge_similarity=0.0, checker_distance=0.0, max_iteration=100000, confidence=0.999, seed=0):
    """registration_ransac_based_on_correspondence, single-threaded bookkeeping, counter-based picks (test/mini1.py:269-281)."""
    src, tgt = _c(src, np.float64), _c(tgt, np.float64)
    corres = np.ascontiguousarray(np.asarray(corres).reshape(-1, 2), dtype=np.int32)
    T = np.empty(16, np.float64)
    st = np.zeros(5, np.float64)
    lib().orc_ransac(_p(src), C.c_int64(len(src)), _p(tgt), C.c_int64(len(tgt)), _p(corres), C.c_int64(len(corres)), C.c_double(dmax), int(ransac_n),
                     C.c_double(edge_similarity), C.c_double(checker_distance), C.c_int64(max_iteration), C.c_double(confidence),
                     C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), _p(T), _p(st))
    return dict(transformation=T.reshape(4, 4), fitness=st[0], inlier_rmse=st[1], n_corr=int(st[2]), iterations=int(st[3]), validated=int(st[4]))


def fgr(src, tgt, feat_src, feat_tgt, division_factor=1.4, use_absolute_scale=False, decrease_mu=True, maximum_correspondence_distance=0.025,
        iteration_number=64, tuple_scale=0.95, maximum_tuple_count=1000, tuple_test=True, seed=0):
    """registration_fgr_based_on_feature_matching (test/check6.py:236-240). Features [n, dim]. -> (T source->target, matches used)"""
    src, tgt, fs, ft = _c(src, np.float64), _c(tgt, np.float64), _c(feat_src, np.float64), _c(feat_tgt, np.float64)
    opt = np.array([division_factor, float(use_absolute_scale), float(decrease_mu), maximum_correspondence_distance, iteration_number, tuple_scale,
                    maximum_tuple_count, float(tuple_test)], np.float64)
    T = np.empty(16, np.float64)
    lib().orc_fgr.restype = C.c_int64
    n = lib().orc_fgr(_p(src), C.c_int64(len(src)), _p(tgt), C.c_int64(len(tgt)), _p(fs), _p(ft), int(fs.shape[1]), _p(opt),
                      C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), _p(T))
    return T.reshape(4, 4), int(n)


P2P, P2L, GICP = 0, 1, 2


def icp(kind, src, tgt, dmax, T0=None, tgt_normals=None, src_cov=None, tgt_cov=None, rel_fitness=1e-6, rel_rmse=1e-6, max_iter=30):
    src = _c(src, np.float64)
    tgt = _c(tgt, np.float64)
    tgt_normals = _c(tgt_normals, np.float64)
    src_cov = _c(src_cov, np.float64)
    tgt_cov = _c(tgt_cov, np.float64)
    T0 = _c(np.eye(4) if T0 is None else T0, np.float64)
    T = np.empty((4, 4), np.float64)
    stats = np.empty(4, np.float64)
    corr = np.empty(max(len(src), 1), np.int32)
    r = lib().orc_icp(int(kind), _p(src), C.c_int64(len(src)), _p(src_cov), _p(tgt), C.c_int64(len(tgt)), _p(tgt_normals), _p(tgt_cov),
                      C.c_double(dmax), _p(T0), C.c_double(rel_fitness), C.c_double(rel_rmse), int(max_iter), _p(T), _p(stats), _p(corr))
    if r == -1:
        raise RuntimeError("Invalid max_correspondence_distance.")
    if r == -2:
        raise RuntimeError("TransformationEstimationPointToPlane and TransformationEstimationColoredICP require pre-computed normal vectors for target PointCloud.")
    if r == -3:
        raise RuntimeError("GeneralizedICP requires covariances.")
    return dict(transformation=T, fitness=stats[0], inlier_rmse=stats[1], iterations=int(stats[2]), n_corr=int(stats[3]), corr=corr[:len(src)])
