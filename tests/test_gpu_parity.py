"""GPU parity tests: every entry point of the C ABI (through the ctypes binding) against the CPU oracle on the same
seeded inputs, and against the reference's own golden fixtures. Bar: bit-exact for integer / index / ordered-sum work,
stated tolerances for floating-point solves (transform 1e-5 rad / 1e-5 m, rmse / fitness 1e-4, normals 1e-9 bulk)."""
import os

import numpy as np
import pytest

import oracle
from util import DEPTH_SCALE, GOLDEN, INTR, golden_cloud, lexorder, rot_err, small_rigid, surface_cloud

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import b200recon
    from b200recon import ops as _ops
    return _ops


# ---- K1 ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(480, 640), (480, 848), (5, 7), (1, 1), (3, 4)])
def test_deproject_z16_bit_exact(ops, shape):
    rng = np.random.default_rng(1)
    depth = rng.integers(0, 65536, size=shape, dtype=np.uint16)
    depth[rng.random(shape) < 0.1] = 0
    color = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    args = (616.6349, 616.3090, 312.5787, 242.2195, 0.001)
    ref = oracle.deproject_z16(depth, *args)
    out = ops.deproject_z16(depth, *args)
    assert out.dtype == np.float32 and np.array_equal(out, ref)
    out2, rgb = ops.deproject_z16(depth, *args, color_bgr=color)
    assert np.array_equal(out2, ref)
    assert np.array_equal(rgb, (color.reshape(-1, 3) / 255.0).astype(np.float32))


def test_deproject_z16_all_zero_and_max(ops):
    for v in (0, 65535):
        depth = np.full((8, 12), v, np.uint16)
        assert np.array_equal(ops.deproject_z16(depth, 424.0, 424.0, 424.0, 240.0, 0.001), oracle.deproject_z16(depth, 424.0, 424.0, 424.0, 240.0, 0.001))


@pytest.mark.parametrize("name", ["output84_00008", "output_00008"])
def test_deproject_rgbd_golden_inputs(ops, name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    color = d["color_rgb"]
    rx, rc = oracle.deproject_rgbd(d["depth"], color, INTR["fx"], INTR["fy"], INTR["ppx"], INTR["ppy"], depth_scale=DEPTH_SCALE)
    gx, gc = ops.deproject_rgbd(d["depth"], color, INTR["fx"], INTR["fy"], INTR["ppx"], INTR["ppy"], depth_scale=DEPTH_SCALE)
    assert gx.shape == rx.shape and np.array_equal(gx, rx) and np.array_equal(gc, rc)


def test_deproject_rgbd_edge_cases(ops):
    rng = np.random.default_rng(3)
    depth = rng.integers(0, 5000, size=(37, 53), dtype=np.uint16)
    for flip in (True, False):
        rx, _ = oracle.deproject_rgbd(depth, None, 500.0, 510.0, 26.0, 18.0, 1000.0, 3.0, flip)
        gx, gc = ops.deproject_rgbd(depth, None, 500.0, 510.0, 26.0, 18.0, 1000.0, 3.0, flip)
        assert gc is None and np.array_equal(gx, rx)
    empty = np.zeros((16, 16), np.uint16)
    gx, _ = ops.deproject_rgbd(empty, None, 500.0, 500.0, 8.0, 8.0)
    assert gx.shape == (0, 3)
    # a raster larger than one scan tile, all valid
    big = np.full((300, 300), 1234, np.uint16)
    rx, _ = oracle.deproject_rgbd(big, None, 500.0, 500.0, 150.0, 150.0)
    gx, _ = ops.deproject_rgbd(big, None, 500.0, 500.0, 150.0, 150.0)
    assert np.array_equal(gx, rx)


def test_reproject_disparity_golden(ops):
    d = np.load(os.path.join(GOLDEN, "disparity_cv2.npz"))
    out = ops.reproject_disparity(d["disp16"], d["Q"])
    fin = np.isfinite(d["xyz"])
    assert np.array_equal(np.isfinite(out), fin) and np.array_equal(out[fin], d["xyz"][fin])
    rng = np.random.default_rng(2000)
    disp = rng.integers(16, 2048, size=(245, 326), dtype=np.int16)
    disp[rng.random(disp.shape) < 0.03] = -16
    ref = oracle.reproject_disparity(disp, d["Q"])
    out = ops.reproject_disparity(disp, d["Q"])
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


# ---- K2 ---------------------------------------------------------------------------------------------------------
def _cmp_voxel_legacy(ops, pts, vs, colors=None, normals=None):
    ref = oracle.voxel_legacy(pts, vs, colors=colors, normals=normals)
    out = ops.voxel_down_sample_legacy(pts, vs, colors=colors, normals=normals)
    assert out["points"].shape == ref["points"].shape
    assert np.array_equal(out["index"], ref["index"])
    assert np.array_equal(out["points"], ref["points"])
    if colors is not None:
        assert np.array_equal(out["colors"], ref["colors"])
    if normals is not None:
        assert np.array_equal(out["normals"], ref["normals"])
    assert out["count"].sum() == len(pts)
    return out


@pytest.mark.parametrize("name", ["output84_00008", "output_00050"])
def test_voxel_legacy_golden_replay(ops, name):
    """depth PNG -> RGB-D deprojection -> legacy voxel 0.02 on the GPU reproduces the reference's PLY point set bit-exactly."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    color = d["color_rgb"] if "color_rgb" in d.files else None
    xyz, rgb = ops.deproject_rgbd(d["depth"], color, INTR["fx"], INTR["fy"], INTR["ppx"], INTR["ppy"], depth_scale=DEPTH_SCALE)
    out = _cmp_voxel_legacy(ops, xyz, 0.02, colors=rgb)
    pts = out["points"]
    if name.startswith("output_"):
        keep, _ = ops.remove_statistical_outlier(pts, 20, 2.0)
        pts = pts[keep]
    gp = d["ply_points"]
    assert len(pts) == len(gp)
    assert np.array_equal(pts[lexorder(pts)], gp[lexorder(gp)])


def test_voxel_legacy_random_and_attrs(ops):
    rng = np.random.default_rng(5)
    pts = rng.normal(0, 1, (200_000, 3))
    cols, nrm = rng.random((200_000, 3)), rng.normal(0, 1, (200_000, 3))
    _cmp_voxel_legacy(ops, pts, 0.05, colors=cols, normals=nrm)
    _cmp_voxel_legacy(ops, pts, 0.05, normals=nrm)
    # coarse voxels -> long runs (block-per-run path), and a voxel size larger than the whole cloud
    _cmp_voxel_legacy(ops, pts, 1.0, colors=cols)
    _cmp_voxel_legacy(ops, pts, 100.0, colors=cols, normals=nrm)


def test_voxel_legacy_edge_cases(ops):
    one = np.array([[1.5, -2.0, 3.0]])
    _cmp_voxel_legacy(ops, one, 0.01)
    same = np.repeat(one, 1000, axis=0)
    out = _cmp_voxel_legacy(ops, same, 0.01)
    assert len(out["points"]) == 1 and out["count"][0] == 1000
    empty = ops.voxel_down_sample_legacy(np.zeros((0, 3)), 0.01)
    assert empty["points"].shape == (0, 3)
    with pytest.raises(RuntimeError, match="voxel_size <= 0"):
        ops.voxel_down_sample_legacy(one, 0.0)
    far = np.array([[0.0, 0.0, 0.0], [1e9, 0.0, 0.0]])
    with pytest.raises(RuntimeError, match="voxel_size is too small"):
        ops.voxel_down_sample_legacy(far, 1e-3)
    with pytest.raises(RuntimeError, match="voxel_size is too small"):
        oracle.voxel_legacy(far, 1e-3)


def test_voxel_tensor_depth_frame_with_holes(ops):
    """rs.pointcloud() output keeps zero-depth pixels as (0,0,0): one voxel swallows thousands of identical points."""
    rng = np.random.default_rng(1000)
    depth = rng.integers(500, 4000, size=(480, 848), dtype=np.uint16)
    depth[rng.random(depth.shape) < 0.05] = 0
    xyz = oracle.deproject_z16(depth, 424.0, 424.0, 424.0, 240.0, 0.001)
    cols = rng.random(xyz.shape).astype(np.float32)
    for vs in (0.005, 0.01, 0.05):
        ref = oracle.voxel_tensor(xyz, vs, attr=cols)
        out = ops.voxel_down_sample_tensor(xyz, vs, attr=cols)
        assert np.array_equal(out["index"], ref["index"])
        assert np.array_equal(out["points"], ref["points"])
        assert np.array_equal(out["attr"], ref["attr"])


def test_voxel_tensor_uniform_cube(ops):
    """test/gpu-performance.py shape (uniform [0,1)^3, voxel 0.05) at 1 M points: runs of ~125 points each."""
    rng = np.random.default_rng(5000)
    pts = rng.random((1_000_000, 3), dtype=np.float32)
    ref = oracle.voxel_tensor(pts, 0.05)
    out = ops.voxel_down_sample_tensor(pts, 0.05)
    assert np.array_equal(out["index"], ref["index"]) and np.array_equal(out["points"], ref["points"])
    neg = (pts - 0.5).astype(np.float32)
    ref = oracle.voxel_tensor(neg, 0.013)
    out = ops.voxel_down_sample_tensor(neg, 0.013)
    assert np.array_equal(out["index"], ref["index"]) and np.array_equal(out["points"], ref["points"])
    with pytest.raises(RuntimeError, match="voxel_size must be positive"):
        ops.voxel_down_sample_tensor(pts[:10], 0.0)


# ---- neighbour search -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("k,radius", [(30, 0.02), (20, 0.0), (1, 0.0), (50, 0.05), (64, 0.03)])
def test_knn_indices_bit_exact(ops, dtype, k, radius):
    pts = surface_cloud(40_000, seed=7).astype(dtype)
    q = np.concatenate([pts[::7], (pts[:500] + 0.003).astype(dtype)])
    ri, rd, rc = oracle.knn(pts, q, k, radius)
    gi, gd, gc = ops.knn(pts, q, k, radius)
    assert np.array_equal(gc, rc)
    assert np.array_equal(gi, ri)
    assert np.array_equal(gd, rd)


def test_knn_ties_and_isolated_points(ops):
    """Raw lattices produce exact distance ties (tie-break: smaller index); far-away points force the whole-cloud scan."""
    g = np.stack(np.meshgrid(np.arange(20.0), np.arange(20.0), np.arange(5.0), indexing="ij"), -1).reshape(-1, 3) * 0.01
    pts = np.concatenate([g, [[5.0, 5.0, 5.0], [-3.0, 0.0, 0.0]]])
    for k, r in ((7, 0.0), (27, 0.015), (10, 0.0101)):
        ri, rd, rc = oracle.knn(pts, pts, k, r)
        gi, gd, gc = ops.knn(pts, pts, k, r)
        assert np.array_equal(gc, rc) and np.array_equal(gi, ri) and np.array_equal(gd, rd)
    few = pts[:5]
    ri, rd, rc = oracle.knn(few, few, 30, 0.0)
    gi, gd, gc = ops.knn(few, few, 30, 0.0)
    assert np.array_equal(gc, rc) and np.array_equal(gi, ri)


# ---- K3 ---------------------------------------------------------------------------------------------------------
def _normal_diff(a, b):
    return np.abs(a - b).max(axis=1)


@pytest.mark.parametrize("name,k", [("output84_00060", 20), ("output_00094", 30)])
def test_normals_legacy_golden(ops, name, k):
    pts, gn = golden_cloud(name)
    ref = oracle.normals_legacy(pts, k, 0.04)
    out = ops.estimate_normals_legacy(pts, k, 0.04)
    d = _normal_diff(out, ref)
    # same neighbour order and summation order as the oracle; only libm (acos/cos) last-ulp differences remain, amplified on
    # near-degenerate neighbourhoods
    assert np.quantile(d, 0.999) < 1e-9 and (d > 1e-6).sum() <= 3, (np.quantile(d, 0.999), d.max())
    dg = _normal_diff(out, gn)
    assert np.quantile(dg, 0.999) < 1e-9 and (dg > 1e-9).sum() <= 8


def test_normals_legacy_variants(ops):
    pts = surface_cloud(60_000, seed=11)
    for k, r in ((30, 0.02), (30, 0.0), (10, 0.004), (64, 0.05)):
        ref = oracle.normals_legacy(pts, k, r)
        out = ops.estimate_normals_legacy(pts, k, r)
        d = _normal_diff(out, ref)
        assert np.quantile(d, 0.999) < 1e-9 and (d > 1e-6).sum() <= 3, (k, r, d.max())
    prior = np.tile([0.0, 0.0, -1.0], (len(pts), 1))
    ref = oracle.normals_legacy(pts, 30, 0.02, prior=prior)
    out = ops.estimate_normals_legacy(pts, 30, 0.02, prior=prior)
    assert np.quantile(_normal_diff(out, ref), 0.999) < 1e-9
    assert (out[:, 2] <= 1e-12).mean() > 0.99
    # fewer than 3 neighbours -> (0,0,1); collinear points -> degenerate covariance handled like the oracle
    iso = np.array([[0.0, 0, 0], [10.0, 0, 0], [0, 10.0, 0]])
    assert np.array_equal(ops.estimate_normals_legacy(iso, 30, 0.5), oracle.normals_legacy(iso, 30, 0.5))
    line = np.column_stack([np.linspace(0, 1, 50), np.zeros(50), np.zeros(50)])
    assert np.allclose(ops.estimate_normals_legacy(line, 10, 0.0), oracle.normals_legacy(line, 10, 0.0), atol=1e-12)


def test_normals_dense_patches(ops):
    """Points whose radius holds many more neighbours than k take the warp-per-point queue kernel (k-nearest cut by (d2, index)):
    patches of 40 / 300 / 700 points inside one ball exercise the single batch, the ranking-with-pruning and the multi-batch path."""
    rng = np.random.default_rng(77)
    base = surface_cloud(20_000, seed=12)
    blobs = []
    for n_blob, c in ((40, 100), (300, 5000), (700, 12000), (700, 19000)):
        ctr = base[c]
        blobs.append(ctr + rng.normal(0.0, 0.002, (n_blob, 3)) * np.array([1.0, 1.0, 0.2]))
    pts = np.concatenate([base] + blobs)
    for k, r in ((30, 0.02), (5, 0.01), (32, 0.03)):
        ref = oracle.normals_legacy(pts, k, r)
        out = ops.estimate_normals_legacy(pts, k, r)
        d = _normal_diff(out, ref)
        assert np.quantile(d, 0.999) < 1e-9 and (d > 1e-6).sum() <= 3, (k, r, np.quantile(d, 0.999), d.max())


def test_normals_tensor(ops):
    pts = surface_cloud(60_000, seed=13).astype(np.float32)
    for k, r in ((50, 0.05), (30, 0.01)):
        ref = oracle.normals_tensor(pts, k, r)
        out = ops.estimate_normals_tensor(pts, k, r)
        d = _normal_diff(out, ref)
        # float32 eigen-solve: libm differences show at 1e-6 relative; stated tolerance 1e-4 on 99.9 %, sign-consistent
        assert np.quantile(d, 0.999) < 1e-4, np.quantile(d, 0.999)
        assert (np.sum(out * ref, axis=1) > 0.99).mean() > 0.999


def test_covariances_from_normals(ops):
    rng = np.random.default_rng(17)
    n = rng.normal(0, 1, (5000, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n[0] = [-1.0, 0, 0]
    n[1] = [1.0, 0, 0]
    assert np.allclose(ops.covariances_from_normals(n, 1e-3), oracle.covariances_from_normals(n, 1e-3), rtol=0, atol=1e-15)


@pytest.mark.parametrize("nb,ratio", [(20, 2.0), (30, 1.2)])
def test_statistical_outlier_kept_set(ops, nb, ratio):
    pts, _ = golden_cloud("output84_00008")
    rng = np.random.default_rng(19)
    pts = np.concatenate([pts, rng.uniform(-2, 2, (300, 3))])
    rk, _ = oracle.statistical_outlier(pts, nb, ratio)
    gk, gi = ops.remove_statistical_outlier(pts, nb, ratio)
    assert np.array_equal(gk, rk)
    assert np.array_equal(gi, np.nonzero(rk)[0])
    with pytest.raises(RuntimeError, match="Illegal input parameters"):
        ops.remove_statistical_outlier(pts, 0, 1.0)


def test_radius_outlier_kept_set(ops):
    pts, _ = golden_cloud("output_00050")
    for nbp, r in ((16, 0.05), (4, 0.03), (1, 0.02)):
        rk = oracle.radius_outlier(pts, nbp, r)
        gk, gi = ops.remove_radius_outlier(pts, nbp, r)
        assert np.array_equal(gk, rk) and np.array_equal(gi, np.nonzero(rk)[0])
    with pytest.raises(RuntimeError, match="Illegal input parameters"):
        ops.remove_radius_outlier(pts, 3, 0.0)


# ---- K4 ---------------------------------------------------------------------------------------------------------
def test_transform_bit_exact(ops):
    rng = np.random.default_rng(23)
    p, n = rng.normal(0, 1, (10_000, 3)), rng.normal(0, 1, (10_000, 3))
    c = rng.normal(0, 1, (10_000, 9))
    T = small_rigid()
    rp, rn, rc = oracle.transform(T, p, n, c)
    gp, gn, gc = ops.transform(T, p, n, c)
    assert np.array_equal(gp, rp) and np.array_equal(gn, rn) and np.array_equal(gc, rc)


@pytest.mark.parametrize("dmax", [0.02, 0.008, 0.1])
def test_correspondence_indices_bit_exact(ops, dmax):
    tgt, _ = golden_cloud("output_00094")
    T = small_rigid()
    src = oracle.transform(np.linalg.inv(T), tgt)[0][::2]
    for Tq in (None, T, small_rigid(0.03, 0.02, -0.01, (0.01, 0.0, -0.01))):
        rc, rn, rs = oracle.correspondences(src, tgt, Tq, dmax)
        gc, gn, gs = ops.correspondences(src, tgt, Tq, dmax)
        assert np.array_equal(gc, rc)
        assert gn == rn and abs(gs - rs) <= 1e-12 * max(1.0, abs(rs))


def test_correspondences_with_exact_ties(ops):
    g = np.stack(np.meshgrid(np.arange(30.0), np.arange(30.0), indexing="ij"), -1).reshape(-1, 2) * 0.01
    tgt = np.column_stack([g, np.zeros(len(g))])
    src = tgt + np.array([0.005, 0.005, 0.0])  # equidistant from four targets
    rc, rn, _ = oracle.correspondences(src, tgt, None, 0.02)
    gc, gn, _ = ops.correspondences(src, tgt, None, 0.02)
    assert np.array_equal(gc, rc) and gn == rn


def _check_icp(res, ref, T_true=None):
    assert rot_err(res["transformation"][:3, :3], ref["transformation"][:3, :3]) < 1e-5
    assert np.linalg.norm(res["transformation"][:3, 3] - ref["transformation"][:3, 3]) < 1e-5
    assert abs(res["fitness"] - ref["fitness"]) < 1e-4 and abs(res["inlier_rmse"] - ref["inlier_rmse"]) < 1e-4
    if T_true is not None:
        assert rot_err(res["transformation"][:3, :3], T_true[:3, :3]) < 1e-5
        assert np.linalg.norm(res["transformation"][:3, 3] - T_true[:3, 3]) < 1e-5


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_icp_known_answer_and_oracle(ops, kind):
    """BASELINE config 1 (substitute cloud): fixture cloud vs a known rigid transform of itself."""
    tgt, nrm = golden_cloud("output_00094")
    T = small_rigid()
    src = oracle.transform(np.linalg.inv(T), tgt)[0]
    kw = {}
    if kind == 1:
        kw["tgt_normals"] = nrm
    if kind == 2:
        tc = oracle.covariances_from_normals(nrm)
        sn = oracle.transform(np.linalg.inv(T), tgt, nrm)[1]
        kw.update(src_cov=oracle.covariances_from_normals(sn).reshape(-1, 9), tgt_cov=tc.reshape(-1, 9))
    max_iter = 100 if kind == 0 else 30
    ref = oracle.icp(kind, src, tgt, 0.02, max_iter=max_iter, **kw)
    res = ops.icp(kind, src, tgt, 0.02, max_iter=max_iter, **kw)
    _check_icp(res, ref, T)
    assert res["fitness"] == 1.0 and res["inlier_rmse"] < 1e-4
    assert abs(res["iterations"] - ref["iterations"]) <= 1
    assert np.array_equal(res["corr"], ref["corr"])


def test_gicp_normal_equations_vs_numpy(ops):
    """The generalized-ICP pass accumulates G^T M^-1 G and G^T M^-1 (p - q) through the inverse Cholesky factor of M. An
    independent numpy statement with the plain inverse of M (cKDTree correspondences, no factor at all) must give the same 27 sums,
    the same count and the same sum of squared distances."""
    from scipy.spatial import cKDTree
    from b200recon import distributed as dist
    tgt, nrm = golden_cloud("output_00094")
    T = small_rigid()
    src, sn, _ = oracle.transform(np.linalg.inv(T), tgt, nrm)
    sc = oracle.covariances_from_normals(sn).reshape(-1, 3, 3)
    tc = oracle.covariances_from_normals(nrm).reshape(-1, 3, 3)
    sh = dist.ShardedICP(2, src, len(src), tgt, 0.02, src_cov=sc.reshape(-1, 9), tgt_cov=tc.reshape(-1, 9), max_iter=1)
    sums = sh.accumulate().cpu().numpy().copy()
    sh.update()
    sh.finish()
    d, j = cKDTree(tgt).query(src, k=1)
    keep = d < 0.02
    p, q = src[keep], tgt[j[keep]]
    A = np.linalg.inv(tc[j[keep]] + sc[keep])  # R = I at the identity start
    G = np.zeros((len(p), 3, 6))
    G[:, 0, 1], G[:, 0, 2] = p[:, 2], -p[:, 1]
    G[:, 1, 0], G[:, 1, 2] = -p[:, 2], p[:, 0]
    G[:, 2, 0], G[:, 2, 1] = p[:, 1], -p[:, 0]
    G[:, :, 3:] = np.eye(3)
    JTJ = np.einsum("nki,nkl,nlj->ij", G, A, G)
    JTr = np.einsum("nki,nkl,nl->i", G, A, p - q)
    ref = np.concatenate([JTJ[np.triu_indices(6)], JTr, [keep.sum(), (d[keep] ** 2).sum()]])
    assert sums[27] == keep.sum()
    assert np.allclose(sums[:29], ref, rtol=1e-9, atol=1e-9 * np.abs(ref[:27]).max()), np.abs(sums[:29] - ref).max()


def test_icp_partial_overlap_with_init(ops):
    tgt, nrm = golden_cloud("output_00050")
    other, _ = golden_cloud("output_00008")
    T = small_rigid(0.02, 0.01, -0.02, (0.01, 0.005, -0.008))
    src = oracle.transform(np.linalg.inv(T), tgt)[0][::3]
    src = np.concatenate([src, other[:800] + 5.0])  # far-away clutter: no correspondences
    init = small_rigid(0.015, 0.0, -0.01, (0.005, 0.0, 0.0))
    for max_iter in (0, 1, 30):
        ref = oracle.icp(1, src, tgt, 0.015, T0=init, tgt_normals=nrm, max_iter=max_iter)
        res = ops.icp(1, src, tgt, 0.015, init=init, tgt_normals=nrm, max_iter=max_iter)
        _check_icp(res, ref)
        assert res["iterations"] == ref["iterations"] and res["n_corr"] == ref["n_corr"]


def test_icp_errors_and_empty(ops):
    tgt, nrm = golden_cloud("output84_00008")
    with pytest.raises(RuntimeError, match="Invalid max_correspondence_distance"):
        ops.icp(0, tgt, tgt, 0.0)
    with pytest.raises(RuntimeError, match="require pre-computed normal vectors"):
        ops.icp(1, tgt, tgt, 0.02)
    with pytest.raises(RuntimeError, match="requires covariances"):
        ops.icp(2, tgt, tgt, 0.02)
    r = ops.icp(0, np.zeros((0, 3)), tgt, 0.02)
    assert r["fitness"] == 0 and r["inlier_rmse"] == 0 and np.array_equal(r["transformation"], np.eye(4))
    r = ops.icp(0, tgt[:100] + 50.0, tgt, 0.02)  # nothing within range: identity, fitness 0
    assert r["fitness"] == 0 and r["n_corr"] == 0 and np.array_equal(r["transformation"], np.eye(4))


def test_information_matrix(ops):
    """get_information_matrix_from_point_clouds (test/mini1.py:302) after a multi-scale point-to-plane refinement (check2.py:143-155)."""
    tgt, nrm = golden_cloud("output_00050")
    T = small_rigid(0.02, 0.01, -0.02, (0.01, 0.005, -0.008))
    src = oracle.transform(np.linalg.inv(T), tgt)[0][::2]
    cur_g, cur_o = np.eye(4), np.eye(4)
    for dist, iters in ((0.3, 30), (0.1, 20), (0.03, 10)):  # voxel 0.02 x (15, 5, 1.5)
        cur_g = ops.icp(1, src, tgt, dist, init=cur_g, tgt_normals=nrm, max_iter=iters)["transformation"]
        cur_o = oracle.icp(1, src, tgt, dist, T0=cur_o, tgt_normals=nrm, max_iter=iters)["transformation"]
    assert rot_err(cur_g[:3, :3], cur_o[:3, :3]) < 1e-5 and np.linalg.norm(cur_g[:3, 3] - cur_o[:3, 3]) < 1e-5
    ref = oracle.information_matrix(src, tgt, 0.03, cur_o)
    out = ops.information_matrix(src, tgt, 0.03, cur_o)
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-9) and np.array_equal(out, out.T)
    assert abs(out[3, 3] - out[4, 4]) < 1e-9 and out[3, 3] > 0.9 * len(src)


def test_correspondences_far_from_origin_and_dense_duplicates(ops):
    """The staged search scans float32 offsets from a local centre: results must stay bit-exact for clouds far from the origin
    (large absolute coordinates) and for targets with many coincident / nearly coincident points (float rounding band)."""
    tgt, _ = golden_cloud("output_00094")
    rng = np.random.default_rng(31)
    T = small_rigid()
    for offset in ((1000.0, -2000.0, 500.0), (-1.0e5, 3.0e4, 7.0e4)):
        t2 = tgt + np.array(offset)
        s2 = oracle.transform(np.linalg.inv(T), tgt)[0][::2] + np.array(offset)
        rc, rn, _ = oracle.correspondences(s2, t2, None, 0.02)
        gc, gn, _ = ops.correspondences(s2, t2, None, 0.02)
        assert np.array_equal(gc, rc) and gn == rn
    # duplicates and 1e-9-scale perturbations of the same targets: the winner is decided by the exact float64 rule
    dup = np.concatenate([tgt, tgt[::3], tgt[::5] + rng.normal(0, 1e-9, (len(tgt[::5]), 3))])
    src = tgt[::2] + rng.normal(0, 1e-4, (len(tgt[::2]), 3))
    rc, rn, rs = oracle.correspondences(src, dup, None, 0.02)
    gc, gn, gs = ops.correspondences(src, dup, None, 0.02)
    assert np.array_equal(gc, rc) and gn == rn


def test_knn_far_from_origin(ops):
    pts = surface_cloud(30_000, seed=37) + np.array([5.0e3, -7.0e3, 1.0e3])
    for k, r in ((30, 0.02), (12, 0.0)):
        ri, rd, rc = oracle.knn(pts, pts[::5], k, r)
        gi, gd, gc = ops.knn(pts, pts[::5], k, r)
        assert np.array_equal(gc, rc) and np.array_equal(gi, ri) and np.array_equal(gd, rd)
    ref = oracle.normals_legacy(pts, 30, 0.02)
    out = ops.estimate_normals_legacy(pts, 30, 0.02)
    # raw-moment covariances lose digits far from the origin (in the reference too): only the well-conditioned bulk is compared
    d = np.abs(out - ref).max(axis=1)
    assert np.quantile(d, 0.9) < 1e-3


def test_icp_batch_equals_single_calls(ops):
    """b3d_icp_batch: pairs of different sizes (one with an empty source) in the same launches; each pair bit-identical to b3d_icp."""
    clouds = [golden_cloud(n) for n in ("output_00094", "output84_00008", "output_00050")]
    T = small_rigid()
    srcs = [oracle.transform(np.linalg.inv(T), c[0])[0][::k] for c, k in zip(clouds, (1, 2, 3))]
    srcs.append(np.zeros((0, 3)))
    tgts = [c[0] for c in clouds] + [clouds[0][0][:500]]
    nrms = [c[1] for c in clouds] + [clouds[0][1][:500]]
    res = ops.icp_batch(1, srcs, tgts, 0.02, tgt_normals=nrms, max_iter=30)
    assert len(res) == 4 and res[3]["fitness"] == 0 and np.array_equal(res[3]["transformation"], np.eye(4))
    for i in range(3):
        single = ops.icp(1, srcs[i], tgts[i], 0.02, tgt_normals=nrms[i], max_iter=30)
        assert np.array_equal(res[i]["transformation"], single["transformation"])
        assert res[i]["fitness"] == single["fitness"] and res[i]["inlier_rmse"] == single["inlier_rmse"] and res[i]["iterations"] == single["iterations"]
        assert np.array_equal(res[i]["corr"], single["corr"])
        ref = oracle.icp(1, srcs[i], tgts[i], 0.02, tgt_normals=nrms[i], max_iter=30)
        assert rot_err(res[i]["transformation"][:3, :3], ref["transformation"][:3, :3]) < 1e-5
        assert np.array_equal(res[i]["corr"], ref["corr"])


def test_orient_normals_consistent_tangent_plane(ops):
    """normal_estimation.py:21: orient_normals_consistent_tangent_plane(100). Same unoriented normals into the CUDA path and the
    oracle (Prim + Kruskal + queue walk): the set of flipped normals must be identical."""
    rng = np.random.default_rng(7)
    pts, nrm = golden_cloud("output_00094")
    nrm = nrm * np.where(rng.random(len(nrm)) < 0.5, -1.0, 1.0)[:, None]  # scramble the signs
    for k in (100, 12):
        ref, ref_flip = oracle.orient_normals(pts, nrm, k)
        out, flip = ops.orient_normals_consistent_tangent_plane(pts, nrm, k)
        assert np.array_equal(flip, ref_flip), int((flip != ref_flip).sum())
        assert np.array_equal(out, ref)
        assert np.array_equal(out, np.where(flip[:, None], -nrm, nrm))
    # two separate pieces (the k-NN graph is disconnected: the Euclidean tree has to bridge the gap) + exact ties on a lattice
    g = np.stack(np.meshgrid(np.arange(24), np.arange(24), indexing="ij"), -1).reshape(-1, 2) * 0.01
    plane = np.column_stack([g, 0.02 * np.sin(8 * g[:, 0])])
    pn = np.column_stack([-0.16 * np.cos(8 * g[:, 0]), np.zeros(len(g)), np.ones(len(g))])
    pn /= np.linalg.norm(pn, axis=1, keepdims=True)
    sph = rng.normal(size=(700, 3))
    sph /= np.linalg.norm(sph, axis=1, keepdims=True)
    both = np.concatenate([plane, 0.05 * sph + [0.1, 0.1, 0.6]])
    bn = np.concatenate([pn, sph]) * np.where(rng.random(len(both)) < 0.5, -1.0, 1.0)[:, None]
    ref, ref_flip = oracle.orient_normals(both, bn, 10)
    out, flip = ops.orient_normals_consistent_tangent_plane(both, bn, 10)
    assert np.array_equal(flip, ref_flip), int((flip != ref_flip).sum())
    # the sphere ends up pointing outwards (its top looks at +z) and the sheet follows it across the gap
    assert (np.einsum("ij,ij->i", out[len(plane):], sph) > 0).all()
    with pytest.raises(RuntimeError, match="Not enough points"):
        ops.orient_normals_consistent_tangent_plane(pts[:3], nrm[:3], 100)
    with pytest.raises(RuntimeError, match="No normals"):
        ops.orient_normals_consistent_tangent_plane(pts, None, 100)


def test_match_features_bit_exact(ops):
    """The feature nearest-neighbour search of registration_ransac_based_on_feature_matching (test/mini1.py:269): FPFH of two
    fixture clouds, indices identical to the oracle's brute force; odd sizes and a dimension that is not a multiple of 4."""
    pa, na = golden_cloud("output_00094")
    pb, nb_ = golden_cloud("output84_00060")
    fa, fb = ops.compute_fpfh(pa, na, 100, 0.1), ops.compute_fpfh(pb, nb_, 100, 0.1)
    assert np.array_equal(ops.match_features(fa, fb), oracle.match_features(fa, fb))
    rng = np.random.default_rng(3)
    for na_, nb2, dim in ((1, 1, 1), (130, 33, 5), (257, 1000, 33), (5, 0, 7)):
        a, b = rng.normal(size=(na_, dim)), rng.normal(size=(nb2, dim))
        if nb2 > 4:
            b[3] = b[1]  # exact tie: the smaller index wins
            a[0] = b[1]
        assert np.array_equal(ops.match_features(a, b), oracle.match_features(a, b))


def test_ransac_correspondence_vs_oracle(ops):
    """registration_ransac_based_on_correspondence (test/mini1.py:269-281: ransac_n 4, edge-length 0.9 and distance checkers,
    confidence 0.999): same picks on the device and in the oracle -> same winning hypothesis."""
    tgt, _ = golden_cloud("output_00094")
    T = small_rigid(0.3, -0.2, 0.5, (0.1, -0.05, 0.2))
    rng = np.random.default_rng(11)
    src = oracle.transform(np.linalg.inv(T), tgt)[0] + rng.normal(0, 2e-4, tgt.shape)
    n = len(src)
    corr = np.stack([np.arange(n), np.arange(n)], 1)
    bad = rng.random(n) < 0.9
    corr[bad, 1] = rng.integers(0, n, int(bad.sum()))
    for seed, rn, edge, dist in ((1, 4, 0.9, 0.01), (2, 3, 0.0, 0.01), (3, 3, 0.9, 0.0)):
        ref = oracle.ransac(src, tgt, corr, 0.01, rn, edge, dist, 20000, 0.999, seed=seed)
        out = ops.ransac_correspondence(src, tgt, corr, 0.01, rn, edge, dist, 20000, 0.999, seed=seed)
        assert out["n_corr"] == ref["n_corr"] and out["validated"] == ref["validated"] and out["iterations"] == ref["iterations"]
        assert np.allclose(out["transformation"], ref["transformation"], atol=1e-9)
        assert abs(out["inlier_rmse"] - ref["inlier_rmse"]) < 1e-12
        assert out["fitness"] > 0.95 and rot_err(out["transformation"][:3, :3], T[:3, :3]) < 5e-3
    # degenerate requests give the library's empty result
    empty = ops.ransac_correspondence(src, tgt, corr, 0.01, 2, 0.9, 0.01, 1000, 0.999)
    assert empty["fitness"] == 0.0 and np.array_equal(empty["transformation"], np.eye(4))
    none = ops.ransac_correspondence(src, tgt, corr, 0.01, 3, 0.0, 0.0, 0, 0.999)  # no iterations allowed
    assert none["n_corr"] == 0 and np.array_equal(none["transformation"], np.eye(4))


def test_fgr_feature_matching_vs_oracle(ops):
    """registration_fgr_based_on_feature_matching (test/check6.py:236-240): the same FPFH features into the CUDA path and the
    oracle -> same number of tuple matches, transforms equal to rounding; the large known motion is recovered."""
    tgt, nrm = golden_cloud("output_00094")
    T = small_rigid(0.35, -0.25, 0.4, (0.3, -0.1, 0.2))
    rng = np.random.default_rng(9)
    src, snrm = oracle.transform(np.linalg.inv(T), tgt, normals=nrm)[:2]
    src = src + rng.normal(0, 1e-4, src.shape)
    fs, ft = ops.compute_fpfh(src, snrm, 100, 0.05), ops.compute_fpfh(tgt, nrm, 100, 0.05)
    for kw in (dict(maximum_correspondence_distance=0.015), dict(maximum_correspondence_distance=0.015, tuple_test=False),
               dict(maximum_correspondence_distance=0.015, use_absolute_scale=True, maximum_tuple_count=300, seed=5)):
        ref_T, ref_n = oracle.fgr(src, tgt, fs, ft, **kw)
        out_T, out_n = ops.fgr_feature_matching(src, tgt, fs, ft, **kw)
        assert out_n == ref_n and out_n >= 10
        assert np.allclose(out_T, ref_T, atol=1e-8), np.abs(out_T - ref_T).max()
        assert rot_err(out_T[:3, :3], T[:3, :3]) < 1e-3 and np.linalg.norm(out_T[:3, 3] - T[:3, 3]) < 1e-3
    # features that match nothing consistently: fewer than 10 tuple matches -> identity, like the library
    junk_T, junk_n = ops.fgr_feature_matching(src[:200], tgt[:200], rng.normal(size=(200, 33)), rng.normal(size=(200, 33)),
                                              maximum_correspondence_distance=0.015)
    assert junk_n < 10 and np.array_equal(junk_T, np.eye(4))


def test_orient_normals_fuzz(ops):
    """Small adversarial clouds for the spanning-tree machinery of orient_normals_consistent_tangent_plane: several far-apart
    clusters (the Euclidean tree has to bridge more than one gap, components of equal size), duplicated points (zero-length
    edges, ties broken by the end points' indices), a lattice (exact distance ties everywhere), tiny clouds (k > n)."""
    rng = np.random.default_rng(77)
    for case in range(24):
        kind = case % 4
        n = int(rng.choice([4, 5, 9, 33, 100, 257, 600]))
        if kind == 0:  # clusters far apart, two of them of equal size
            m = max(n // 4, 1)
            pts = np.concatenate([rng.normal(c, 0.02, (m, 3)) for c in ((0, 0, 0), (1.0, 0.2, 0.1), (-0.7, 0.9, 0.3), (0.1, -1.2, 2.0))])
        elif kind == 1:  # duplicates
            pts = rng.normal(0, 0.1, (n, 3))
            pts[rng.random(n) < 0.3] = pts[0]
        elif kind == 2:  # lattice
            s = int(np.ceil(n ** (1 / 3)))
            pts = (np.stack(np.meshgrid(np.arange(s), np.arange(s), np.arange(s), indexing="ij"), -1).reshape(-1, 3)[:max(n, 4)] * 0.01).astype(np.float64)
        else:  # a noisy sphere
            pts = rng.normal(size=(n, 3))
            pts = pts / np.linalg.norm(pts, axis=1, keepdims=True) * (1 + rng.normal(0, 0.01, (n, 1)))
        pts = np.ascontiguousarray(pts + rng.choice([0.0, 5.0]))
        nrm = rng.normal(size=pts.shape)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        if kind == 3:
            nrm = pts / np.linalg.norm(pts, axis=1, keepdims=True) * np.where(rng.random(len(pts)) < 0.5, -1.0, 1.0)[:, None]
        k = int(rng.choice([10, 16, 100]))
        ref, ref_flip = oracle.orient_normals(pts, nrm, k)
        out, flip = ops.orient_normals_consistent_tangent_plane(pts, nrm, k)
        assert np.array_equal(flip, ref_flip), (case, kind, len(pts), k, int((flip != ref_flip).sum()))
        assert np.array_equal(out, ref)


def test_fpfh_features(ops):
    """compute_fpfh_feature(Hybrid(0.1, 100)) as in test/mini1.py:244-250 on a fixture cloud with its own normals."""
    pts, nrm = golden_cloud("output_00094")
    for k, r in ((100, 0.1), (30, 0.05), (20, 0.0)):
        ref = oracle.fpfh(pts, nrm, k, r)
        out = ops.compute_fpfh(pts, nrm, k, r)
        assert out.shape == ref.shape == (len(pts), 33)
        # histogram bins are integer decisions on float64 features (libm acos / atan2 differ in the last ulp): a handful of
        # points may move one increment between adjacent bins; everything else agrees to rounding
        bad = np.abs(out - ref).max(axis=1) > 1e-6
        assert bad.mean() < 2e-3, bad.mean()
        assert np.allclose(out.sum(axis=1), ref.sum(axis=1), atol=1e-6)
    with pytest.raises(RuntimeError, match="max_nn must be in"):
        ops.compute_fpfh(pts[:10], nrm[:10], 200, 0.1)


def test_fuzz_small_random_clouds(ops):
    """Many small random configurations (sizes 1..3000, clustered / duplicated / planar / far-apart points, random voxel sizes,
    k and radii): every integer / index result bit-exact against the oracle. Exercises the tile edges of the scan, the
    one-chunk / one-cell grids and the isolated-point fallbacks."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        n = int(rng.choice([1, 2, 3, 5, 31, 32, 33, 64, 100, 257, 1000, 3000]))
        kind = case % 4
        if kind == 0:
            pts = rng.normal(0, 0.05, (n, 3))
        elif kind == 1:  # planar patch with duplicates
            pts = np.column_stack([rng.random(n) * 0.2, rng.random(n) * 0.2, np.zeros(n)])
            pts[rng.random(n) < 0.2] = pts[0]
        elif kind == 2:  # two clusters far apart
            pts = np.concatenate([rng.normal(0, 0.01, (n // 2 + 1, 3)), rng.normal(5.0, 0.01, (n - n // 2 - 1 if n > 1 else 0, 3))])[:n]
        else:  # lattice (exact ties)
            s = int(np.ceil(n ** (1 / 3)))
            g = np.stack(np.meshgrid(np.arange(s), np.arange(s), np.arange(s), indexing="ij"), -1).reshape(-1, 3)[:n] * 0.01
            pts = g.astype(np.float64)
        pts = np.ascontiguousarray(pts + rng.choice([0.0, 10.0, -3.0]))
        vs = float(rng.choice([0.003, 0.01, 0.05, 1.0]))
        ref = oracle.voxel_legacy(pts, vs)
        out = ops.voxel_down_sample_legacy(pts, vs)
        assert np.array_equal(out["index"], ref["index"]) and np.array_equal(out["points"], ref["points"]), (case, n, vs)
        p32 = pts.astype(np.float32)
        ref = oracle.voxel_tensor(p32, vs)
        out = ops.voxel_down_sample_tensor(p32, vs)
        assert np.array_equal(out["index"], ref["index"]) and np.array_equal(out["points"], ref["points"]), (case, n, vs)
        k = int(rng.choice([1, 5, 30, 64]))
        r = float(rng.choice([0.0, 0.015, 0.05]))
        ri, rd, rc = oracle.knn(pts, pts, k, r)
        gi, gd, gc = ops.knn(pts, pts, k, r)
        assert np.array_equal(gc, rc) and np.array_equal(gi, ri) and np.array_equal(gd, rd), (case, n, k, r)
        src = pts + rng.normal(0, 0.002, pts.shape)
        dmax = float(rng.choice([0.005, 0.02, 0.3]))
        rcorr, rn, _ = oracle.correspondences(src, pts, None, dmax)
        gcorr, gn, _ = ops.correspondences(src, pts, None, dmax)
        assert np.array_equal(gcorr, rcorr) and gn == rn, (case, n, dmax)
        if n >= 3:
            rk = oracle.radius_outlier(pts, 2, 0.02)
            gk, _ = ops.remove_radius_outlier(pts, 2, 0.02)
            assert np.array_equal(gk, rk), (case, n)
