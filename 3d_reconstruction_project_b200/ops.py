"""Functional API over the C ABI: numpy (host) in, numpy out; every call runs the CUDA kernels of libb200recon.so.

Each function names the reference call site it stands in for (paths under the reference root). Inputs may also be
CUDA torch tensors (then no upload happens); outputs are numpy unless ``as_tensor=True``.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from .context import get_context, ptr


def _out(t, as_tensor):
    return t if as_tensor else t.cpu().numpy()


def _dev(ctx, a, dtype):
    return None if a is None else ctx.to_device(a, dtype)


def _check_n3(t, name):
    if t is not None and (t.dim() != 2 or t.shape[1] != 3):
        raise ValueError(f"{name} must have shape [N, 3], got {tuple(t.shape)}")


# ---- K1 ------------------------------------------------------------------------------------------------------------
def deproject_z16(depth, fx, fy, ppx, ppy, depth_scale=0.001, color_bgr=None, device=0, as_tensor=False):
    """rs.pointcloud().calculate(depth_frame).get_vertices() -- pointcloud_capture.py:35,38 (+ colours/255, :39)."""
    ctx = get_context(device)
    d = ctx.to_device(np.ascontiguousarray(depth, dtype=np.uint16).view(np.int16) if not isinstance(depth, torch.Tensor) else depth)
    h, w = d.shape
    xyz = ctx.empty((h * w, 3), torch.float32)
    if color_bgr is None:
        N.check(N.lib().b3d_deproject_z16(ctx.handle, ptr(d), w, h, fx, fy, ppx, ppy, depth_scale, ptr(xyz)))
        return _out(xyz, as_tensor)
    c = ctx.to_device(color_bgr, torch.uint8)
    rgb = ctx.empty((h * w, 3), torch.float32)
    N.check(N.lib().b3d_deproject_z16_color(ctx.handle, ptr(d), ptr(c), w, h, fx, fy, ppx, ppy, depth_scale, ptr(xyz), ptr(rgb)))
    return _out(xyz, as_tensor), _out(rgb, as_tensor)


def deproject_rgbd(depth, color, fx, fy, cx, cy, depth_scale=1000.0, depth_trunc=3.0, flip=True, device=0, as_tensor=False):
    """RGBDImage.create_from_color_and_depth + PointCloud.create_from_rgbd_image (+ flip) -- test/check84.py:155-178."""
    ctx = get_context(device)
    d = ctx.to_device(np.ascontiguousarray(depth, dtype=np.uint16).view(np.int16) if not isinstance(depth, torch.Tensor) else depth)
    h, w = d.shape
    c = _dev(ctx, color, torch.uint8)
    xyz = ctx.empty((h * w, 3), torch.float64)
    rgb = ctx.empty((h * w, 3), torch.float64) if c is not None else None
    n = C.c_int64(0)
    N.check(N.lib().b3d_deproject_rgbd(ctx.handle, ptr(d), ptr(c), w, h, fx, fy, cx, cy, float(np.float32(depth_scale)), float(np.float32(depth_trunc)),
                                       int(bool(flip)), ptr(xyz), ptr(rgb), C.byref(n)))
    m = n.value
    return _out(xyz[:m], as_tensor), (None if rgb is None else _out(rgb[:m], as_tensor))


def reproject_disparity(disp16, Q, device=0, as_tensor=False):
    """cv2.reprojectImageTo3D(disp16 / 16, Q) for the Q loaded at Calib_depth/depth4.py:98."""
    ctx = get_context(device)
    d = ctx.to_device(disp16, torch.int16)
    h, w = d.shape
    Qh = (C.c_double * 16)(*np.asarray(Q, dtype=np.float64).reshape(16))
    xyz = ctx.empty((h, w, 3), torch.float32)
    N.check(N.lib().b3d_reproject_disparity(ctx.handle, ptr(d), w, h, Qh, ptr(xyz)))
    return _out(xyz, as_tensor)


def reproject_disparity_valid(disp16, Q, min_disp16=16, device=0, as_tensor=False):
    """Valid pixels only (disparity >= min_disp16; SGBM marks invalid pixels with (minDisparity - 1) * 16), raster order."""
    ctx = get_context(device)
    d = ctx.to_device(disp16, torch.int16)
    h, w = d.shape
    Qh = (C.c_double * 16)(*np.asarray(Q, dtype=np.float64).reshape(16))
    xyz = ctx.empty((h * w, 3), torch.float32)
    n = C.c_int64(0)
    N.check(N.lib().b3d_reproject_disparity_valid(ctx.handle, ptr(d), w, h, Qh, int(min_disp16), ptr(xyz), C.byref(n)))
    return _out(xyz[:n.value], as_tensor)


# ---- K2 ------------------------------------------------------------------------------------------------------------
def voxel_down_sample_legacy(points, voxel_size, colors=None, normals=None, device=0, as_tensor=False):
    """o3d.geometry.PointCloud.voxel_down_sample -- pointcloud_alignment.py:22-23. Returns dict(points, colors, normals,
    index [M,3] int32, count [M])."""
    ctx = get_context(device)
    p = ctx.to_device(points, torch.float64)
    _check_n3(p, "points")
    c, nr = _dev(ctx, colors, torch.float64), _dev(ctx, normals, torch.float64)
    n = p.shape[0]
    cap = max(n, 1)
    o_p = ctx.empty((cap, 3), torch.float64)
    o_c = ctx.empty((cap, 3), torch.float64) if c is not None else None
    o_n = ctx.empty((cap, 3), torch.float64) if nr is not None else None
    o_i = ctx.empty((cap, 3), torch.int32)
    o_k = ctx.empty((cap,), torch.int32)
    m = C.c_int64(0)
    N.check(N.lib().b3d_voxel_downsample_legacy(ctx.handle, ptr(p), ptr(c), ptr(nr), n, float(voxel_size), ptr(o_p), ptr(o_c), ptr(o_n), ptr(o_i), ptr(o_k),
                                                C.byref(m)))
    m = m.value
    f = lambda t: None if t is None else _out(t[:m], as_tensor)
    return dict(points=f(o_p), colors=f(o_c), normals=f(o_n), index=f(o_i), count=f(o_k))


def voxel_down_sample_tensor(points, voxel_size, attr=None, device=0, as_tensor=False):
    """o3d.t.geometry.PointCloud.voxel_down_sample -- pointcloud_capture.py:50, pointcloud_processing.py:27."""
    ctx = get_context(device)
    p = ctx.to_device(points, torch.float32)
    _check_n3(p, "points")
    a = _dev(ctx, attr, torch.float32)
    n = p.shape[0]
    cap = max(n, 1)
    o_p = ctx.empty((cap, 3), torch.float32)
    o_a = ctx.empty((cap, 3), torch.float32) if a is not None else None
    o_i = ctx.empty((cap, 3), torch.int64)
    o_k = ctx.empty((cap,), torch.int32)
    m = C.c_int64(0)
    N.check(N.lib().b3d_voxel_downsample_tensor(ctx.handle, ptr(p), ptr(a), n, float(np.float32(voxel_size)), ptr(o_p), ptr(o_a), ptr(o_i), ptr(o_k),
                                                C.byref(m)))
    m = m.value
    f = lambda t: None if t is None else _out(t[:m], as_tensor)
    return dict(points=f(o_p), attr=f(o_a), index=f(o_i), count=f(o_k))


def knn(points, queries, k, radius=0.0, device=0, as_tensor=False):
    """KDTreeFlann.search_hybrid_vector_3d / search_knn_vector_3d: (idx [nq,k] -1 padded, d2 [nq,k], cnt [nq])."""
    ctx = get_context(device)
    f64 = (points.dtype == np.float64) if not isinstance(points, torch.Tensor) else points.dtype == torch.float64
    dt = torch.float64 if f64 else torch.float32
    p, q = ctx.to_device(points, dt), ctx.to_device(queries, dt)
    g = C.c_void_p()
    N.check(N.lib().b3d_grid_build(ctx.handle, ptr(p), p.shape[0], int(f64), 0.0, int(k), float(radius), C.byref(g)))
    try:
        nq = q.shape[0]
        idx = ctx.empty((nq, k), torch.int32)
        d2 = ctx.empty((nq, k), dt)
        cnt = ctx.empty((nq,), torch.int32)
        N.check(N.lib().b3d_knn_hybrid(ctx.handle, g, ptr(q), nq, int(k), float(radius), ptr(idx), ptr(d2), ptr(cnt)))
        ctx.synchronize()
    finally:
        N.lib().b3d_grid_destroy(ctx.handle, g)
    return _out(idx, as_tensor), _out(d2, as_tensor), _out(cnt, as_tensor)


# ---- K3 ------------------------------------------------------------------------------------------------------------
def estimate_normals_legacy(points, max_nn, radius, prior=None, device=0, as_tensor=False):
    """PointCloud.estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) -- pointcloud_alignment.py:27-28. radius<=0: KNN."""
    ctx = get_context(device)
    p = ctx.to_device(points, torch.float64)
    _check_n3(p, "points")
    pr = _dev(ctx, prior, torch.float64)
    out = ctx.empty(tuple(p.shape), torch.float64)
    N.check(N.lib().b3d_estimate_normals_legacy(ctx.handle, ptr(p), p.shape[0], int(max_nn), float(radius), ptr(pr), ptr(out)))
    return _out(out, as_tensor)


def estimate_normals_tensor(points, max_nn, radius, device=0, as_tensor=False):
    """t.PointCloud.estimate_normals(max_nn, radius) -- normal_estimation.py:20."""
    ctx = get_context(device)
    p = ctx.to_device(points, torch.float32)
    _check_n3(p, "points")
    out = ctx.empty(tuple(p.shape), torch.float32)
    N.check(N.lib().b3d_estimate_normals_tensor(ctx.handle, ptr(p), p.shape[0], int(max_nn), float(np.float32(radius)), ptr(out)))
    return _out(out, as_tensor)


def covariances_from_normals(normals, eps=1e-3, device=0, as_tensor=False):
    """GICP covariances C = R diag(eps,1,1) R^T (inside registration_generalized_icp, test/GICP1.py:99-102)."""
    ctx = get_context(device)
    nr = ctx.to_device(normals, torch.float64)
    cov = ctx.empty((nr.shape[0], 3, 3), torch.float64)
    N.check(N.lib().b3d_covariances_from_normals(ctx.handle, ptr(nr), nr.shape[0], float(eps), ptr(cov)))
    return _out(cov, as_tensor)


def orient_normals_consistent_tangent_plane(points, normals, k=100, device=0, as_tensor=False):
    """PointCloud.orient_normals_consistent_tangent_plane(k) -- normal_estimation.py:21. -> (oriented normals f64 [n,3], flipped bool [n])"""
    ctx = get_context(device)
    p = ctx.to_device(points, torch.float64)
    _check_n3(p, "points")
    if normals is None:
        raise RuntimeError("No normals in the PointCloud. Call EstimateNormals() first.")
    nr = ctx.to_device(normals, torch.float64).clone()
    _check_n3(nr, "normals")
    fl = ctx.empty((max(p.shape[0], 1),), torch.uint8)
    N.check(N.lib().b3d_orient_normals_consistent_tangent_plane(ctx.handle, ptr(p), ptr(nr), p.shape[0], int(k), ptr(fl)))
    fl = fl[:p.shape[0]].bool()
    return _out(nr, as_tensor), _out(fl, as_tensor)


def compute_fpfh(points, normals, max_nn, radius, device=0, as_tensor=False):
    """o3d.pipelines.registration.compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius, max_nn)) -- test/mini1.py:244-250.
    -> [N, 33] float64 (the transpose of Open3D's Feature.data)."""
    ctx = get_context(device)
    p, nr = ctx.to_device(points, torch.float64), ctx.to_device(normals, torch.float64)
    _check_n3(p, "points")
    out = ctx.empty((p.shape[0], 33), torch.float64)
    N.check(N.lib().b3d_compute_fpfh(ctx.handle, ptr(p), ptr(nr), p.shape[0], int(max_nn), float(radius), ptr(out)))
    return _out(out, as_tensor)


def match_features(feat_a, feat_b, device=0, as_tensor=False):
    """Nearest feature of feat_b ([nb, dim]) for every row of feat_a ([na, dim]): the k-d tree search on FPFH features inside
    registration_ransac_based_on_feature_matching -- test/mini1.py:269. -> int32 [na]"""
    ctx = get_context(device)
    a, b = ctx.to_device(feat_a, torch.float64), ctx.to_device(feat_b, torch.float64)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("features must be [n, dim] arrays of the same dimension")
    out = ctx.empty((max(a.shape[0], 1),), torch.int32)
    N.check(N.lib().b3d_match_features(ctx.handle, ptr(a), a.shape[0], ptr(b), b.shape[0], int(a.shape[1]), ptr(out)))
    return _out(out[:a.shape[0]], as_tensor)


def ransac_correspondence(src, tgt, corres, max_dist, ransac_n=3, edge_similarity=0.0, checker_distance=0.0, max_iteration=100000,
                          confidence=0.999, seed=0, device=0):
    """registration_ransac_based_on_correspondence (point-to-point estimation) -- the loop behind test/mini1.py:269-281.
    corres: int [nc, 2] (source index, target index)."""
    ctx = get_context(device)
    s, t = ctx.to_device(src, torch.float64), ctx.to_device(tgt, torch.float64)
    _check_n3(s, "source points")
    _check_n3(t, "target points")
    c = ctx.to_device(np.ascontiguousarray(np.asarray(corres).reshape(-1, 2), dtype=np.int32) if not isinstance(corres, torch.Tensor) else corres,
                      torch.int32)
    r = N.RansacResult()
    N.check(N.lib().b3d_ransac_correspondence(ctx.handle, ptr(s), s.shape[0], ptr(t), t.shape[0], ptr(c), c.shape[0], float(max_dist), int(ransac_n),
                                              float(edge_similarity), float(checker_distance), int(max_iteration), float(confidence),
                                              C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.byref(r)))
    return dict(transformation=np.array(r.transformation[:], dtype=np.float64).reshape(4, 4), fitness=r.fitness, inlier_rmse=r.inlier_rmse,
                n_corr=int(r.n_correspondences), iterations=int(r.iterations), validated=int(r.validated))


def fgr_feature_matching(src, tgt, feat_src, feat_tgt, division_factor=1.4, use_absolute_scale=False, decrease_mu=True,
                         maximum_correspondence_distance=0.025, iteration_number=64, tuple_scale=0.95, maximum_tuple_count=1000, tuple_test=True,
                         seed=0, device=0):
    """registration_fgr_based_on_feature_matching -- test/check6.py:236-240. Features [n, dim]. -> (T source->target, matches used)"""
    ctx = get_context(device)
    s, t = ctx.to_device(src, torch.float64), ctx.to_device(tgt, torch.float64)
    _check_n3(s, "source points")
    _check_n3(t, "target points")
    fs, ft = ctx.to_device(feat_src, torch.float64), ctx.to_device(feat_tgt, torch.float64)
    if fs.dim() != 2 or ft.dim() != 2 or fs.shape[1] != ft.shape[1] or fs.shape[0] != s.shape[0] or ft.shape[0] != t.shape[0]:
        raise ValueError("features must be [n_points, dim] arrays of the same dimension")
    opt = N.FgrOption(float(division_factor), int(bool(use_absolute_scale)), int(bool(decrease_mu)), float(maximum_correspondence_distance),
                      int(iteration_number), float(tuple_scale), int(maximum_tuple_count), int(bool(tuple_test)))
    T = (C.c_double * 16)()
    n = C.c_int64(0)
    N.check(N.lib().b3d_fgr_feature_matching(ctx.handle, ptr(s), s.shape[0], ptr(t), t.shape[0], ptr(fs), ptr(ft), int(fs.shape[1]), C.byref(opt),
                                             C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), T, C.byref(n)))
    return np.array(T[:], dtype=np.float64).reshape(4, 4), int(n.value)


def _outlier(fn, points, a, b, device, as_tensor):
    ctx = get_context(device)
    p = ctx.to_device(points, torch.float64)
    _check_n3(p, "points")
    n = p.shape[0]
    keep = ctx.empty((max(n, 1),), torch.uint8)
    idx = ctx.empty((max(n, 1),), torch.int64)
    m = C.c_int64(0)
    N.check(fn(ctx.handle, ptr(p), n, int(a), float(b), ptr(keep), ptr(idx), C.byref(m)))
    return _out(keep[:n].bool(), as_tensor), _out(idx[:m.value], as_tensor)


def remove_statistical_outlier(points, nb_neighbors, std_ratio, device=0, as_tensor=False):
    """PointCloud.remove_statistical_outlier -- pointcloud_processing.py:35-36, test/mini1.py:175. -> (keep mask, kept indices)"""
    return _outlier(N.lib().b3d_statistical_outlier, points, nb_neighbors, std_ratio, device, as_tensor)


def remove_radius_outlier(points, nb_points, radius, device=0, as_tensor=False):
    """PointCloud.remove_radius_outlier -- pointcloud_processing.py:39. -> (keep mask, kept indices)"""
    return _outlier(N.lib().b3d_radius_outlier, points, nb_points, radius, device, as_tensor)


def select_rows(src, idx, device=0, as_tensor=False):
    """select_by_index on an [N, cols] float64 array (pointcloud_processing.py:36)."""
    ctx = get_context(device)
    s = ctx.to_device(src, torch.float64)
    i = ctx.to_device(idx, torch.int64)
    cols = int(np.prod(s.shape[1:])) if s.dim() > 1 else 1
    out = ctx.empty((i.shape[0],) + tuple(s.shape[1:]), torch.float64)
    N.check(N.lib().b3d_gather_rows_f64(ctx.handle, ptr(s), ptr(i), i.shape[0], cols, ptr(out)))
    return _out(out, as_tensor)


# ---- K4 ------------------------------------------------------------------------------------------------------------
def _T16(T):
    return None if T is None else (C.c_double * 16)(*np.asarray(T, dtype=np.float64).reshape(16))


def transform(T, points, normals=None, cov=None, device=0, as_tensor=False):
    """PointCloud.transform(T) -- pointcloud_alignment.py:42. Returns new arrays (points, normals, cov)."""
    ctx = get_context(device)
    p = ctx.to_device(points, torch.float64).clone()
    nr = None if normals is None else ctx.to_device(normals, torch.float64).clone()
    cv = None if cov is None else ctx.to_device(cov, torch.float64).clone()
    N.check(N.lib().b3d_transform_f64(ctx.handle, _T16(T), ptr(p), p.shape[0], ptr(nr), ptr(cv)))
    f = lambda t: None if t is None else _out(t, as_tensor)
    return f(p), f(nr), f(cv)


def correspondences(src, tgt, T=None, max_dist=0.02, device=0, as_tensor=False):
    """One correspondence search at a fixed transform (GetRegistrationResultAndCorrespondences). -> (corr, n, sum_d2)"""
    ctx = get_context(device)
    s, t = ctx.to_device(src, torch.float64), ctx.to_device(tgt, torch.float64)
    corr = ctx.empty((max(s.shape[0], 1),), torch.int32)
    st = (C.c_double * 2)()
    N.check(N.lib().b3d_icp_correspondences(ctx.handle, ptr(s), s.shape[0], ptr(t), t.shape[0], _T16(T), float(max_dist), ptr(corr), st))
    ctx.synchronize()
    return _out(corr[:s.shape[0]], as_tensor), int(st[0]), float(st[1])


def information_matrix(src, tgt, max_dist, T=None, device=0):
    """get_information_matrix_from_point_clouds(source, target, max_dist, T) -- test/mini1.py:302. -> [6,6] float64"""
    ctx = get_context(device)
    s, t = ctx.to_device(src, torch.float64), ctx.to_device(tgt, torch.float64)
    out = (C.c_double * 36)()
    N.check(N.lib().b3d_information_matrix(ctx.handle, ptr(s), s.shape[0], ptr(t), t.shape[0], _T16(T), float(max_dist), out))
    return np.array(out[:], dtype=np.float64).reshape(6, 6)


def _result_dict(r, corr):
    return dict(transformation=np.array(r.transformation[:], dtype=np.float64).reshape(4, 4), fitness=r.fitness, inlier_rmse=r.inlier_rmse,
                iterations=int(r.iterations), converged=bool(r.converged), n_corr=int(r.n_correspondences), corr=corr)


def icp(kind, src, tgt, max_dist, init=None, tgt_normals=None, src_cov=None, tgt_cov=None, rel_fitness=1e-6, rel_rmse=1e-6, max_iter=30,
        device=0, want_corr=True):
    """registration_icp / registration_generalized_icp -- pointcloud_alignment.py:35-39 (P2P), test/mini1.py:293-296 (P2L),
    test/GICP1.py:99-102 (GICP). kind in {0,1,2}."""
    ctx = get_context(device)
    s, t = ctx.to_device(src, torch.float64), ctx.to_device(tgt, torch.float64)
    tn, sc, tc = _dev(ctx, tgt_normals, torch.float64), _dev(ctx, src_cov, torch.float64), _dev(ctx, tgt_cov, torch.float64)
    corr = ctx.empty((max(s.shape[0], 1),), torch.int32) if want_corr else None
    r = N.IcpResult()
    N.check(N.lib().b3d_icp(ctx.handle, int(kind), ptr(s), s.shape[0], ptr(sc), ptr(t), t.shape[0], ptr(tn), ptr(tc), float(max_dist), _T16(init),
                            float(rel_fitness), float(rel_rmse), int(max_iter), C.byref(r), ptr(corr)))
    return _result_dict(r, None if corr is None else corr[:s.shape[0]].cpu().numpy())


def icp_batch(kind, srcs, tgts, max_dist, inits=None, tgt_normals=None, src_covs=None, tgt_covs=None, rel_fitness=1e-6, rel_rmse=1e-6,
              max_iter=30, device=0, want_corr=True):
    """A batch of independent registrations in the same launches (b3d_icp_batch). srcs / tgts (and the optional per-cloud
    normals / covariances): lists of [n_i, 3] arrays. Returns one result dict per pair, each equal to the single-pair icp()."""
    ctx = get_context(device)
    P = len(srcs)
    if P == 0 or len(tgts) != P:
        raise ValueError("icp_batch needs equally long, non-empty lists of sources and targets")
    cat = lambda arrs, cols: None if arrs is None else ctx.to_device(np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1, cols) for a in arrs]), torch.float64)
    so = np.concatenate([[0], np.cumsum([len(a) for a in srcs])]).astype(np.int64)
    to = np.concatenate([[0], np.cumsum([len(a) for a in tgts])]).astype(np.int64)
    s, t = cat(srcs, 3), cat(tgts, 3)
    tn, sc, tc = cat(tgt_normals, 3), cat(src_covs, 9), cat(tgt_covs, 9)
    corr = ctx.empty((max(int(so[-1]), 1),), torch.int32) if want_corr else None
    init = None if inits is None else (C.c_double * (16 * P))(*np.asarray(inits, dtype=np.float64).reshape(-1))
    res = (N.IcpResult * P)()
    N.check(N.lib().b3d_icp_batch(ctx.handle, int(kind), P, ptr(s), so.ctypes.data_as(C.POINTER(C.c_int64)), ptr(sc), ptr(t),
                                  to.ctypes.data_as(C.POINTER(C.c_int64)), ptr(tn), ptr(tc), float(max_dist), init, float(rel_fitness), float(rel_rmse),
                                  int(max_iter), res, ptr(corr)))
    ch = None if corr is None else corr.cpu().numpy()
    return [_result_dict(res[p], None if ch is None else ch[so[p]:so[p + 1]]) for p in range(P)]


# ---- whole path ----------------------------------------------------------------------------------------------------
def make_pair_params(w, h, fx, fy, ppx, ppy, depth_scale=0.001, voxel_size=0.005, normals_max_nn=30, normals_radius=0.01,
                     icp_kind=N.ICP_POINT_TO_PLANE, icp_max_dist=0.02, icp_rel_fitness=1e-6, icp_rel_rmse=1e-6, icp_max_iter=30):
    return N.PairParams(w, h, fx, fy, ppx, ppy, depth_scale, voxel_size, normals_max_nn, normals_radius, icp_kind, icp_max_dist, icp_rel_fitness,
                        icp_rel_rmse, icp_max_iter)


def make_disparity_params(w, h, Q, min_disp16=16, voxel_size=0.005, normals_max_nn=30, normals_radius=0.01, icp_kind=N.ICP_GENERALIZED,
                          icp_max_dist=0.02, icp_rel_fitness=1e-6, icp_rel_rmse=1e-6, icp_max_iter=30):
    Qa = (C.c_double * 16)(*np.asarray(Q, dtype=np.float64).reshape(16))
    return N.DisparityParams(w, h, Qa, min_disp16, voxel_size, normals_max_nn, normals_radius, icp_kind, icp_max_dist, icp_rel_fitness, icp_rel_rmse,
                             icp_max_iter)


def register_disparity_pairs(disp_src, disp_tgt, params, device=0):
    """BASELINE config 3 path for a batch of stereo frame pairs: disparity (int16 x16) + Q -> clouds -> tensor voxel -> normals
    -> ICP / GICP. disp_src / disp_tgt: [P,h,w] int16, host (numpy / pinned tensor) or CUDA tensors."""
    return _register(N.lib().b3d_register_disparity_pairs, disp_src, disp_tgt, params, np.int16, device)


def register_depth_pairs(depth_src, depth_tgt, params, device=0):
    """The full front end + registration for a batch of frame pairs. depth_src / depth_tgt: [P,h,w] uint16 as numpy /
    pinned CPU tensors (copied host->device inside the call) or CUDA tensors (already resident).
    Returns a list of dicts (transformation, fitness, inlier_rmse, iterations, converged, n_corr, n_raw, m_source, m_target)."""
    return _register(N.lib().b3d_register_depth_pairs, depth_src, depth_tgt, params, np.uint16, device)


def _register(fn, src, tgt, params, np_dtype, device):
    ctx = get_context(device)

    def prep(a):
        if isinstance(a, torch.Tensor):
            t = a.contiguous()
            if t.dim() == 2:
                t = t.unsqueeze(0)
            return t, t.is_cuda, t.data_ptr()
        arr = np.ascontiguousarray(a, dtype=np_dtype)
        if arr.ndim == 2:
            arr = arr[None]
        return arr, False, arr.ctypes.data

    s, s_dev, s_ptr = prep(src)
    t, t_dev, t_ptr = prep(tgt)
    if s_dev != t_dev:
        raise ValueError("source and target rasters must both be host or both be device buffers")
    P = s.shape[0]
    if tuple(s.shape) != tuple(t.shape) or tuple(s.shape[1:]) != (params.h, params.w):
        raise ValueError(f"raster stacks must both be [P, {params.h}, {params.w}]")
    res = (N.PairResult * P)()
    N.check(fn(ctx.handle, C.byref(params), C.c_void_p(s_ptr), C.c_void_p(t_ptr), P, int(s_dev), res))
    out = []
    for r in res:
        d = _result_dict(r.icp, None)
        d.update(n_raw=int(r.n_raw), m_source=int(r.m_source), m_target=int(r.m_target))
        out.append(d)
    return out
