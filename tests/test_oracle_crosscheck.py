"""Independent CPU cross-checks of the oracle rows that no reference-held vector pins (a1, a4, a7, a9, a10, a11).
`GPU == oracle` only proves that two restatements by the same hand agree; here the oracle is held against THIRD implementations
written with different libraries and different algorithms: scipy's cKDTree for every neighbour search, numpy.linalg.eigh for the
eigenvectors, a numpy Gauss-Newton / SVD / scipy.linalg.sqrtm for the registration steps, numpy float32 accumulation for the
voxel means. None of this runs on the GPU; the GPU tests compare the CUDA path with the same oracle."""
import numpy as np
import pytest
from scipy.linalg import sqrtm
from scipy.spatial import cKDTree

import oracle
from util import golden_cloud, rot_err, small_rigid, surface_cloud


# ---- a1: librealsense deprojection ----------------------------------------------------------------------------------------------
def test_deproject_z16_vs_numpy_float32():
    rng = np.random.default_rng(31)
    h, w = 60, 80
    depth = rng.integers(0, 4000, (h, w)).astype(np.uint16)
    depth[rng.random((h, w)) < 0.1] = 0
    fx, fy, ppx, ppy, scale = np.float32(61.5), np.float32(60.25), np.float32(39.5), np.float32(30.25), np.float32(0.001)
    out = oracle.deproject_z16(depth, fx, fy, ppx, ppy, scale).reshape(h, w, 3)
    # rs2_deproject_pixel_to_point without distortion, every operation rounded to float32
    j, i = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    z = scale * depth.astype(np.float32)
    x = ((j - ppx) / fx).astype(np.float32)
    y = ((i - ppy) / fy).astype(np.float32)
    ref = np.stack([(z * x).astype(np.float32), (z * y).astype(np.float32), z], -1)
    assert out.dtype == np.float32 and np.array_equal(out, ref)
    assert np.all(out[depth == 0] == 0)


# ---- a4: tensor voxel_down_sample ---------------------------------------------------------------------------------------------
def test_voxel_tensor_vs_numpy_sequential_float32():
    rng = np.random.default_rng(32)
    pts = rng.uniform(-1.0, 1.5, (20000, 3)).astype(np.float32)
    pts[:50] = pts[0]  # a long run inside one voxel
    col = rng.random((20000, 3)).astype(np.float32)
    vs = np.float32(0.05)
    out = oracle.voxel_tensor(pts, vs, attr=col)
    # floor(positions / voxel) in float32, origin 0; IndexAdd in point order with float32 accumulators; divide by the float32 count
    key = np.floor(pts / vs).astype(np.int64)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)  # lexicographic = the canonical (ix, iy, iz) output order
    inv = inv.reshape(-1)
    sums = np.zeros((len(uniq), 3), np.float32)
    csum = np.zeros((len(uniq), 3), np.float32)
    cnt = np.zeros(len(uniq), np.float32)
    for p in range(len(pts)):  # sequential float32 accumulation, the order the CPU IndexAdd uses
        sums[inv[p]] += pts[p]
        csum[inv[p]] += col[p]
        cnt[inv[p]] += np.float32(1)
    assert np.array_equal(out["index"], uniq)
    assert np.array_equal(out["points"], sums / cnt[:, None])
    assert np.array_equal(out["attr"], csum / cnt[:, None])


# ---- a7: radius outlier -------------------------------------------------------------------------------------------------------
def test_radius_outlier_vs_ckdtree():
    pts, _ = golden_cloud("output84_00008")
    rng = np.random.default_rng(33)
    pts = np.concatenate([pts, rng.uniform(-1, 1, (200, 3))])
    r, nb = 0.05, 16
    keep = oracle.radius_outlier(pts, nb, r)
    tree = cKDTree(pts)
    d, _ = tree.query(pts, k=96, distance_upper_bound=r * 1.01)
    assert not np.isfinite(d[:, -1]).any() or (d[:, :nb + 2] < r * 0.99).all(1)[np.isfinite(d[:, -1])].all()  # k is large enough to decide count > nb
    # strict d2 < r2 (self included); a band around the radius is left out of the comparison (sqrt of the tree vs squared distances)
    cnt_lo = (d < r * (1 - 1e-9)).sum(1)
    cnt_hi = (d <= r * (1 + 1e-9)).sum(1)
    sure = cnt_lo == cnt_hi
    assert sure.mean() > 0.999
    assert np.array_equal(keep[sure], cnt_lo[sure] > nb)
    assert 0.05 < keep.mean() < 1.0 and not keep[-200:].any()


# ---- a6 / a9 searches: k nearest, hybrid ----------------------------------------------------------------------------------------
def test_knn_vs_ckdtree():
    pts = surface_cloud(8000, seed=34)
    q = pts[::7] + 1e-3
    idx, d2, cnt = oracle.knn(pts, q, 12)
    dref, iref = cKDTree(pts).query(q, k=12)
    assert (cnt == 12).all()
    assert np.allclose(np.sqrt(d2), dref, rtol=1e-12, atol=0)
    # the index sets agree except where two neighbours tie to rounding
    same = (np.sort(idx, 1) == np.sort(iref, 1)).all(1)
    assert same.mean() > 0.999
    # hybrid: the k nearest cut at the radius
    idx, d2, cnt = oracle.knn(pts, q, 12, radius=0.01)
    assert np.array_equal(cnt, (dref < 0.01 - 1e-12).sum(1)) or np.abs(cnt - (dref < 0.01).sum(1)).max() <= 1


# ---- a9: tensor normals ---------------------------------------------------------------------------------------------------------
def test_normals_tensor_vs_eigh():
    pts = surface_cloud(6000, seed=35).astype(np.float32)
    k, r = 30, 0.02
    out = oracle.normals_tensor(pts, k, r)
    tree = cKDTree(pts.astype(np.float64))
    d, idx = tree.query(pts.astype(np.float64), k=k, distance_upper_bound=r)
    bad = 0
    for i in range(0, len(pts), 5):
        nb = idx[i][np.isfinite(d[i])]
        if len(nb) < 3:
            assert np.array_equal(out[i], [0, 0, 1])
            continue
        P = pts[nb].astype(np.float64)
        C = np.cov(P.T)  # two-pass centred, Bessel (n - 1): SURVEY appendix A.5
        w, v = np.linalg.eigh(C)
        ref = v[:, 0]
        # float32 eigen-solve in the oracle: compare directions where the two smallest eigenvalues are well separated
        if w[1] > 4 * max(w[0], 1e-12):
            c = abs(float(np.dot(out[i].astype(np.float64), ref)))
            bad += c < 1 - 1e-4
    assert bad <= 2


# ---- a10: correspondences, point-to-point, point-to-plane --------------------------------------------------------------------
def _ckd_corr(src, tgt, T, dmax):
    p = src @ T[:3, :3].T + T[:3, 3]
    d, j = cKDTree(tgt).query(p, k=1, distance_upper_bound=dmax)
    ok = np.isfinite(d) & (d * d < dmax * dmax)
    return p, np.where(ok, j, -1), d


def _euler(x):
    ca, sa, cb, sb, cg, sg = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
    T = np.eye(4)
    T[:3, :3] = [[cg * cb, cg * sb * sa - sg * ca, cg * sb * ca + sg * sa], [sg * cb, sg * sb * sa + cg * ca, sg * sb * ca - cg * sa],
                 [-sb, cb * sa, cb * ca]]
    T[:3, 3] = x[3:]
    return T


def _numpy_icp(kind, src, tgt, dmax, nrm=None, scov=None, tcov=None, max_iter=30, rf=1e-6, rr=1e-6):
    """RegistrationICP restated with numpy / scipy: cKDTree correspondences, Umeyama by SVD (P2P), Gauss-Newton
    with the Euler update (P2L), the generalized-ICP rows with scipy.linalg.sqrtm of the inverse (GICP)."""
    T = np.eye(4)
    prev = None
    it = 0
    while True:
        p, j, d = _ckd_corr(src, tgt, T, dmax)
        m = j >= 0
        fit = m.mean()
        rmse = np.sqrt((d[m] ** 2).mean()) if m.any() else 0.0
        if prev is not None and abs(prev[0] - fit) < rf and abs(prev[1] - rmse) < rr:
            break
        if it >= max_iter:
            break
        s, t = p[m], tgt[j[m]]
        if kind == 0:
            ms, mt = s.mean(0), t.mean(0)
            S = (t - mt).T @ (s - ms) / len(s)
            U, _, Vt = np.linalg.svd(S)
            D = np.eye(3)
            if np.linalg.det(U) * np.linalg.det(Vt) < 0:
                D[2, 2] = -1
            R = U @ D @ Vt
            Up = np.eye(4)
            Up[:3, :3] = R
            Up[:3, 3] = mt - R @ ms
        else:
            if kind == 1:
                n = nrm[j[m]]
                r = ((s - t) * n).sum(1)
                J = np.concatenate([np.cross(s, n), n], 1)
            else:
                R = T[:3, :3]
                rows_J, rows_r = [], []
                sc = scov[m]
                for a in range(len(s)):
                    M = tcov[j[m][a]] + R @ sc[a] @ R.T
                    Wm = np.real(sqrtm(np.linalg.inv(M)))
                    sx = np.array([[0, -s[a, 2], s[a, 1]], [s[a, 2], 0, -s[a, 0]], [-s[a, 1], s[a, 0], 0]])
                    rows_J.append(Wm @ np.concatenate([-sx, np.eye(3)], 1))
                    rows_r.append(Wm @ (s[a] - t[a]))
                J, r = np.concatenate(rows_J), np.concatenate(rows_r)
            x = np.linalg.solve(J.T @ J, -J.T @ r)
            Up = _euler(x)
        T = Up @ T
        prev = (fit, rmse)
        it += 1
    return dict(transformation=T, fitness=fit, inlier_rmse=rmse, iterations=it, corr=j)


def test_correspondences_vs_ckdtree():
    tgt, _ = golden_cloud("output_00094")
    T = small_rigid(0.01, -0.02, 0.015, (0.006, -0.004, 0.003))
    src = tgt[::2] + 1e-4
    corr, n, s2 = oracle.correspondences(src, tgt, T, 0.01)
    _, j, d = _ckd_corr(src, tgt, T, 0.01)
    # identical except where the nearest two targets tie to rounding
    assert (corr == j).mean() > 0.9995 and abs(n - (j >= 0).sum()) <= 2
    assert abs(s2 - (d[j >= 0] ** 2).sum()) < 1e-9 * max(1.0, s2) + 1e-6 * 2


@pytest.mark.parametrize("kind", [0, 1])
def test_icp_vs_numpy(kind):
    tgt, nrm = golden_cloud("output_00094")
    T = small_rigid()
    Ti = np.linalg.inv(T)
    src = (tgt @ Ti[:3, :3].T + Ti[:3, 3])[::3]
    max_iter = 60 if kind == 0 else 30
    ref = _numpy_icp(kind, src, tgt, 0.02, nrm=nrm, max_iter=max_iter)
    out = oracle.icp(kind, src, tgt, 0.02, tgt_normals=nrm if kind == 1 else None, max_iter=max_iter)
    assert abs(out["iterations"] - ref["iterations"]) <= 1
    assert rot_err(out["transformation"][:3, :3], ref["transformation"][:3, :3]) < 1e-7
    assert np.linalg.norm(out["transformation"][:3, 3] - ref["transformation"][:3, 3]) < 1e-7
    assert abs(out["fitness"] - ref["fitness"]) < 1e-4 and abs(out["inlier_rmse"] - ref["inlier_rmse"]) < 1e-7
    if kind == 1:  # converges onto the known motion
        assert rot_err(out["transformation"][:3, :3], T[:3, :3]) < 1e-5 and np.linalg.norm(out["transformation"][:3, 3] - T[:3, 3]) < 1e-5


# ---- a11: generalized ICP ---------------------------------------------------------------------------------------------------------
def test_covariances_from_normals_vs_rodrigues():
    rng = np.random.default_rng(36)
    n = rng.normal(size=(300, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    C = oracle.covariances_from_normals(n, 1e-3)
    for i in range(len(n)):
        if n[i, 0] < -0.99:  # the library keeps R = I when the normal is (nearly) opposite to e1
            assert np.array_equal(C[i], np.diag([1e-3, 1.0, 1.0]))
            continue
        # C = R diag(eps, 1, 1) R^T with R e1 = n: eigenvalue eps along the normal, 1 across it
        assert np.allclose(C[i] @ n[i], 1e-3 * n[i], atol=1e-12)
        w = np.linalg.eigvalsh(C[i])
        assert np.allclose(w, [1e-3, 1.0, 1.0], atol=1e-12)


def test_gicp_vs_numpy_sqrtm():
    tgt, nrm = golden_cloud("output_00094")
    tgt, nrm = tgt[::4], nrm[::4]
    T = small_rigid(0.004, -0.006, 0.005, (0.002, -0.0015, 0.001))
    Ti = np.linalg.inv(T)
    src, snrm = tgt @ Ti[:3, :3].T + Ti[:3, 3], nrm @ Ti[:3, :3].T
    src, snrm = src[::3], snrm[::3]
    sc, tc = oracle.covariances_from_normals(snrm), oracle.covariances_from_normals(nrm)
    ref = _numpy_icp(2, src, tgt, 0.02, scov=sc, tcov=tc, max_iter=4, rf=0.0, rr=0.0)
    out = oracle.icp(2, src, tgt, 0.02, src_cov=sc.reshape(-1, 9), tgt_cov=tc.reshape(-1, 9), max_iter=4, rel_fitness=0.0, rel_rmse=0.0)
    assert out["iterations"] == ref["iterations"] == 4
    assert rot_err(out["transformation"][:3, :3], ref["transformation"][:3, :3]) < 1e-8
    assert np.linalg.norm(out["transformation"][:3, 3] - ref["transformation"][:3, 3]) < 1e-8
    assert abs(out["inlier_rmse"] - ref["inlier_rmse"]) < 1e-9
