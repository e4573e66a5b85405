"""o3d.pipelines.registration as used on the reference's path: registration_icp with the point-to-point
(pointcloud_alignment.py:35-39) and point-to-plane (test/mini1.py:293-296, check2.py:151-154) estimators, and
registration_generalized_icp (test/GICP1.py:99-102). The whole ICP loop runs on the device (b3d_icp)."""
import numpy as np

from . import _native as N
from . import ops
from .geometry import as_cloud


class ICPConvergenceCriteria:
    def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
        self.relative_fitness, self.relative_rmse, self.max_iteration = float(relative_fitness), float(relative_rmse), int(max_iteration)


class TransformationEstimationPointToPoint:
    kind = N.ICP_POINT_TO_POINT

    def __init__(self, with_scaling=False):
        if with_scaling:
            raise RuntimeError("with_scaling=True is not on the reference's path and is not implemented")


class TransformationEstimationPointToPlane:
    kind = N.ICP_POINT_TO_PLANE


class TransformationEstimationForGeneralizedICP:
    kind = N.ICP_GENERALIZED

    def __init__(self, epsilon=1e-3):
        self.epsilon = float(epsilon)


class RegistrationResult:
    def __init__(self, d):
        self.transformation = d["transformation"]
        self.fitness = d["fitness"]
        self.inlier_rmse = d["inlier_rmse"]
        self.iterations = d["iterations"]
        self.converged = d["converged"]
        corr = d.get("corr")
        if corr is None:
            self.correspondence_set = np.zeros((0, 2), np.int32)
        else:
            src = np.nonzero(corr >= 0)[0].astype(np.int32)
            self.correspondence_set = np.stack([src, corr[src]], axis=1)

    def __repr__(self):
        return (f"RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e}, and correspondence_set size of "
                f"{len(self.correspondence_set)}\nAccess transformation to get result.")


def _run(source, target, max_correspondence_distance, init, estimation, criteria):
    source, target = as_cloud(source), as_cloud(target)
    if max_correspondence_distance <= 0:
        raise RuntimeError("Invalid max_correspondence_distance.")
    criteria = criteria or ICPConvergenceCriteria()
    init = np.eye(4) if init is None else np.asarray(init, dtype=np.float64).reshape(4, 4)
    kind = estimation.kind
    kw = {}
    if kind == N.ICP_POINT_TO_PLANE:
        if not target.has_normals():
            raise RuntimeError("TransformationEstimationPointToPlane and TransformationEstimationColoredICP require pre-computed normal vectors "
                               "for target PointCloud.")
        kw["tgt_normals"] = np.asarray(target.normals)
    if kind == N.ICP_GENERALIZED:
        # InitializePointCloudForGeneralizedICP works on COPIES of the inputs: covariances the caller set are used as they are,
        # missing ones are derived here (normals by KNN(20) if absent) and never written back into the caller's clouds
        eps = estimation.epsilon
        covs = []
        for c in (source, target):
            if not c.has_covariances():
                c = c.clone().estimate_covariances_from_normals(eps)
            covs.append(c.covariances.reshape(-1, 9))
        kw["src_cov"], kw["tgt_cov"] = covs
    if not source.has_points() or not target.has_points():
        return RegistrationResult(dict(transformation=init.copy(), fitness=0.0, inlier_rmse=0.0, iterations=0, converged=False, corr=None))
    d = ops.icp(kind, np.asarray(source.points), np.asarray(target.points), max_correspondence_distance, init=init,
                rel_fitness=criteria.relative_fitness, rel_rmse=criteria.relative_rmse, max_iter=criteria.max_iteration, device=source.device, **kw)
    return RegistrationResult(d)


def registration_icp(source, target, max_correspondence_distance, init=None, estimation_method=None, criteria=None):
    return _run(source, target, max_correspondence_distance, init, estimation_method or TransformationEstimationPointToPoint(), criteria)


def registration_generalized_icp(source, target, max_correspondence_distance, init=None, estimation_method=None, criteria=None):
    return _run(source, target, max_correspondence_distance, init, estimation_method or TransformationEstimationForGeneralizedICP(), criteria)


def evaluate_registration(source, target, max_correspondence_distance, transformation=None):
    source, target = as_cloud(source), as_cloud(target)
    T = np.eye(4) if transformation is None else np.asarray(transformation, dtype=np.float64)
    ns = len(source.points)
    if ns == 0 or len(target.points) == 0:
        return RegistrationResult(dict(transformation=T, fitness=0.0, inlier_rmse=0.0, iterations=0, converged=False, corr=None))
    corr, n, s = ops.correspondences(np.asarray(source.points), np.asarray(target.points), T, max_correspondence_distance, device=source.device)
    return RegistrationResult(dict(transformation=T, fitness=(n / ns if n else 0.0), inlier_rmse=(np.sqrt(s / n) if n else 0.0), iterations=0,
                                   converged=False, corr=corr))


def get_information_matrix_from_point_clouds(source, target, max_correspondence_distance, transformation):
    """test/mini1.py:302, test/check2.py:160 -- the 6x6 information matrix the reference's pose graph edges carry."""
    source, target = as_cloud(source), as_cloud(target)
    if not source.has_points() or not target.has_points():
        return np.zeros((6, 6))
    return ops.information_matrix(np.asarray(source.points), np.asarray(target.points), max_correspondence_distance, transformation, device=source.device)


class Feature:
    """o3d.pipelines.registration.Feature: data is [dimension, N]."""

    def __init__(self, data):
        self.data = data

    def dimension(self):
        return self.data.shape[0]

    def num(self):
        return self.data.shape[1]


def compute_fpfh_feature(input, search_param):
    """test/mini1.py:244-250, test/check2.py:95-100: FPFH descriptors (33 x N)."""
    from .geometry import KDTreeSearchParamHybrid, KDTreeSearchParamKNN
    pcd = as_cloud(input)
    if not pcd.has_normals():
        raise RuntimeError("Failed because input point cloud has no normal.")
    if isinstance(search_param, KDTreeSearchParamHybrid):
        k, r = search_param.max_nn, search_param.radius
    elif isinstance(search_param, KDTreeSearchParamKNN):
        k, r = search_param.knn, 0.0
    else:
        raise RuntimeError("compute_fpfh_feature: only KDTreeSearchParamHybrid / KDTreeSearchParamKNN are supported")
    f = ops.compute_fpfh(np.asarray(pcd.points), np.asarray(pcd.normals), k, r, device=pcd.device)
    return Feature(np.ascontiguousarray(f.T))


# ---- global registration from feature matches (test/mini1.py:269-281, test/check2.py:132-144, test/check3.py:181) ----------
class CorrespondenceCheckerBasedOnEdgeLength:
    def __init__(self, similarity_threshold=0.9):
        self.similarity_threshold = float(similarity_threshold)


class CorrespondenceCheckerBasedOnDistance:
    def __init__(self, distance_threshold):
        self.distance_threshold = float(distance_threshold)


class RANSACConvergenceCriteria:
    def __init__(self, max_iteration=100000, confidence=0.999):
        self.max_iteration, self.confidence = int(max_iteration), float(confidence)


def _feature_rows(f):
    """Feature / [dim, N] array (Open3D's layout) -> [N, dim] rows."""
    data = f.data if isinstance(f, Feature) else f
    return np.ascontiguousarray(np.asarray(data, dtype=np.float64).T)


def _checker_params(checkers):
    edge, dist = 0.0, 0.0
    for c in checkers or []:
        if isinstance(c, CorrespondenceCheckerBasedOnEdgeLength):
            edge = c.similarity_threshold
        elif isinstance(c, CorrespondenceCheckerBasedOnDistance):
            dist = c.distance_threshold
        else:
            raise RuntimeError("only CorrespondenceCheckerBasedOnEdgeLength / CorrespondenceCheckerBasedOnDistance are supported on this path")
    return edge, dist


def registration_ransac_based_on_correspondence(source, target, corres, max_correspondence_distance, estimation_method=None, ransac_n=3,
                                                checkers=None, criteria=None, seed=0):
    """Hypotheses from ransac_n random correspondences, cheap checkers, validation of the survivors against the whole target;
    the best (fitness, then rmse) wins. `seed` selects the random picks (the library seeds from the system unless told)."""
    source, target = as_cloud(source), as_cloud(target)
    estimation_method = estimation_method or TransformationEstimationPointToPoint(False)
    if not isinstance(estimation_method, TransformationEstimationPointToPoint):
        raise RuntimeError("RANSAC registration is built for TransformationEstimationPointToPoint (the reference's choice)")
    criteria = criteria or RANSACConvergenceCriteria()
    edge, dist = _checker_params(checkers)
    sp, tp = np.asarray(source.points), np.asarray(target.points)
    r = ops.ransac_correspondence(sp, tp, corres, max_correspondence_distance, ransac_n, edge, dist, criteria.max_iteration, criteria.confidence,
                                  seed, device=source.device)
    if r["n_corr"] == 0:
        return RegistrationResult(dict(transformation=np.eye(4), fitness=0.0, inlier_rmse=0.0, iterations=r["iterations"], converged=False, corr=None))
    # GetRegistrationResultAndCorrespondences at the winning transform: fitness, rmse and the correspondence set
    ev = evaluate_registration(source, target, max_correspondence_distance, r["transformation"])
    ev.iterations = r["iterations"]
    return ev


def registration_ransac_based_on_feature_matching(source, target, source_feature, target_feature, mutual_filter, max_correspondence_distance,
                                                  estimation_method=None, ransac_n=3, checkers=None, criteria=None, seed=0):
    """test/mini1.py:269-281 (ransac_n=4, edge length 0.9 + distance checkers, 4 000 000 iterations at confidence 0.999)."""
    source, target = as_cloud(source), as_cloud(target)
    if ransac_n < 3 or max_correspondence_distance <= 0:
        return RegistrationResult(dict(transformation=np.eye(4), fitness=0.0, inlier_rmse=0.0, iterations=0, converged=False, corr=None))
    fs, ft = _feature_rows(source_feature), _feature_rows(target_feature)
    if len(fs) != len(source.points) or len(ft) != len(target.points):
        raise RuntimeError("feature count does not match the point count")
    ij = ops.match_features(fs, ft, device=source.device)
    corres = np.stack([np.arange(len(ij), dtype=np.int32), ij], axis=1)
    if mutual_filter:
        ji = ops.match_features(ft, fs, device=source.device)
        keep = ji[ij] == np.arange(len(ij))
        if int(keep.sum()) >= ransac_n:  # too few mutual matches: fall back to all of them, like the library
            corres = corres[keep]
    return registration_ransac_based_on_correspondence(source, target, corres, max_correspondence_distance, estimation_method, ransac_n,
                                                       checkers, criteria, seed)


class FastGlobalRegistrationOption:
    """o3d.pipelines.registration.FastGlobalRegistrationOption (test/check6.py:238-239 sets maximum_correspondence_distance)."""

    def __init__(self, division_factor=1.4, use_absolute_scale=False, decrease_mu=True, maximum_correspondence_distance=0.025,
                 iteration_number=64, tuple_scale=0.95, maximum_tuple_count=1000, tuple_test=True, seed=None):
        self.division_factor, self.use_absolute_scale, self.decrease_mu = float(division_factor), bool(use_absolute_scale), bool(decrease_mu)
        self.maximum_correspondence_distance, self.iteration_number = float(maximum_correspondence_distance), int(iteration_number)
        self.tuple_scale, self.maximum_tuple_count, self.tuple_test, self.seed = float(tuple_scale), int(maximum_tuple_count), bool(tuple_test), seed


def registration_fgr_based_on_feature_matching(source, target, source_feature, target_feature, option=None):
    """test/check6.py:236-240, check7.py:245, check8.py:244, check81.py:242: Fast Global Registration from FPFH matches."""
    source, target = as_cloud(source), as_cloud(target)
    option = option or FastGlobalRegistrationOption()
    if not source.has_points() or not target.has_points():
        raise RuntimeError("FastGlobalRegistration: source or target point cloud is empty.")
    fs, ft = _feature_rows(source_feature), _feature_rows(target_feature)
    T, _ = ops.fgr_feature_matching(np.asarray(source.points), np.asarray(target.points), fs, ft, option.division_factor, option.use_absolute_scale,
                                    option.decrease_mu, option.maximum_correspondence_distance, option.iteration_number, option.tuple_scale,
                                    option.maximum_tuple_count, option.tuple_test, seed=(option.seed or 0), device=source.device)
    return evaluate_registration(source, target, option.maximum_correspondence_distance, T)
