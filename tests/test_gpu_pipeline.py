"""GPU tests of the whole path: the batched frame-pair pipeline (b3d_register_depth_pairs) against the oracle chain, the
reference-facing classes, and the step-wise (sharded) ICP."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from util import GOLDEN, golden_cloud, rot_err, small_rigid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b3():
    import b200recon
    return b200recon


def oracle_pair(src_depth, tgt_depth, cam, voxel=0.005, k=30, radius=0.01, kind=1, dmax=0.02, max_iter=30):
    """CPU restatement of the pipeline: pointcloud_capture.py:35-53 then pointcloud_alignment.py:27-39 (estimator per kind)."""
    a = (cam["fx"], cam["fy"], cam["ppx"], cam["ppy"], cam["depth_scale"])
    vs = oracle.voxel_tensor(oracle.deproject_z16(src_depth, *a), voxel)["points"].astype(np.float64)
    vt = oracle.voxel_tensor(oracle.deproject_z16(tgt_depth, *a), voxel)["points"].astype(np.float64)
    nt = oracle.normals_legacy(vt, k, radius)
    kw = dict(tgt_normals=nt)
    if kind == 2:
        ns = oracle.normals_legacy(vs, k, radius)
        kw = dict(src_cov=oracle.covariances_from_normals(ns).reshape(-1, 9), tgt_cov=oracle.covariances_from_normals(nt).reshape(-1, 9))
    r = oracle.icp(kind, vs, vt, dmax, max_iter=max_iter, **kw)
    r.update(m_source=len(vs), m_target=len(vt))
    return r


SMALL_CAM = dict(w=212, h=120, fx=106.0, fy=106.0, ppx=106.0, ppy=60.0, depth_scale=0.001)


def _pairs(n, cam):
    from b200recon import synth
    return synth.depth_pairs(n, base_seed=3000, cam=cam)


@pytest.mark.parametrize("kind", [1, 0, 2])
def test_pair_pipeline_vs_oracle_small(b3, kind):
    from b200recon import ops
    src, tgt, T_true = _pairs(3, SMALL_CAM)
    voxel, radius, dmax = 0.02, 0.05, 0.05
    params = ops.make_pair_params(**SMALL_CAM, voxel_size=voxel, normals_max_nn=30, normals_radius=radius, icp_kind=kind, icp_max_dist=dmax, icp_max_iter=30)
    batch = ops.register_depth_pairs(src, tgt, params)
    assert len(batch) == 3
    for i in range(3):
        ref = oracle_pair(src[i], tgt[i], SMALL_CAM, voxel, 30, radius, kind, dmax, 30)
        r = batch[i]
        assert r["m_source"] == ref["m_source"] and r["m_target"] == ref["m_target"]
        assert r["n_raw"] == 2 * SMALL_CAM["w"] * SMALL_CAM["h"]
        assert rot_err(r["transformation"][:3, :3], ref["transformation"][:3, :3]) < 1e-5
        assert np.linalg.norm(r["transformation"][:3, 3] - ref["transformation"][:3, 3]) < 1e-5
        assert abs(r["fitness"] - ref["fitness"]) < 1e-4 and abs(r["inlier_rmse"] - ref["inlier_rmse"]) < 1e-4
        # every pair of the batch equals the single-pair call bit for bit (same kernels, same order of operations)
        single = ops.register_depth_pairs(src[i], tgt[i], params)[0]
        assert np.array_equal(single["transformation"], r["transformation"])
        assert single["fitness"] == r["fitness"] and single["inlier_rmse"] == r["inlier_rmse"] and single["iterations"] == r["iterations"]


def test_pair_pipeline_config2_full_size(b3):
    """BASELINE config 2 at full size (848x480, voxel 5 mm, Hybrid(0.01, 30), point-to-plane, d_max 0.02, 30 iterations)."""
    import torch
    from b200recon import ops, synth
    src, tgt, T_true = synth.depth_pairs(2, base_seed=1000)
    params = ops.make_pair_params(**synth.D435)
    host = ops.register_depth_pairs(src, tgt, params)
    dev = ops.register_depth_pairs(torch.from_numpy(src.view(np.int16)).cuda(), torch.from_numpy(tgt.view(np.int16)).cuda(), params)
    for i in range(2):
        ref = oracle_pair(src[i], tgt[i], synth.D435)
        for r in (host[i], dev[i]):
            assert r["m_source"] == ref["m_source"] and r["m_target"] == ref["m_target"]
            assert rot_err(r["transformation"][:3, :3], ref["transformation"][:3, :3]) < 1e-5
            assert np.linalg.norm(r["transformation"][:3, 3] - ref["transformation"][:3, 3]) < 1e-5
            assert abs(r["fitness"] - ref["fitness"]) < 1e-4 and abs(r["inlier_rmse"] - ref["inlier_rmse"]) < 1e-4
        assert np.array_equal(host[i]["transformation"], dev[i]["transformation"])
        # the registration recovers the camera motion of the synthetic pair to depth-quantisation accuracy
        assert rot_err(host[i]["transformation"][:3, :3], T_true[i][:3, :3]) < 2e-3
        assert np.linalg.norm(host[i]["transformation"][:3, 3] - T_true[i][:3, 3]) < 5e-3


def test_sticky_correspondences_change_nothing(b3):
    """The pass kernel skips the search of a lane whose partner provably cannot have changed and bounds the others by the
    previous partner's distance. With that switched off (B3D_ICP_NO_STICKY: every pass searches every lane at d_max) a
    full-size config-2 registration must come out bit for bit the same, correspondence set included."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json, hashlib; sys.path.insert(0, %r); import numpy as np\n"
        "from b200recon import ops, synth\n"
        "cam = synth.D435; out = []\n"
        "for seed in (1000, 3004):\n"
        "    ds, dt, _ = synth.depth_pair(seed, seed + 1, cam)\n"
        "    cl = []\n"
        "    for d in (ds, dt):\n"
        "        x = ops.deproject_z16(d, cam['fx'], cam['fy'], cam['ppx'], cam['ppy'], cam['depth_scale'])\n"
        "        cl.append(np.ascontiguousarray(ops.voxel_down_sample_tensor(x[x[:, 2] > 0], 0.005)['points'], dtype=np.float64))\n"
        "    nrm = ops.estimate_normals_legacy(cl[1], 30, 0.01)\n"
        "    for kind in (1, 0):\n"
        "        r = ops.icp(kind, cl[0], cl[1], 0.02, tgt_normals=nrm, max_iter=30)\n"
        "        out.append([r['transformation'].tobytes().hex(), r['fitness'], r['inlier_rmse'], r['iterations'],\n"
        "                    hashlib.sha1(np.ascontiguousarray(r['corr']).tobytes()).hexdigest()])\n"
        "print(json.dumps(out))\n" % root)
    runs = []
    for off in (False, True):
        env = dict(os.environ)
        env.pop("B3D_ICP_NO_STICKY", None)
        if off:
            env["B3D_ICP_NO_STICKY"] = "1"
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert runs[0] == runs[1]
    assert all(row[3] >= 3 for row in runs[0])


def test_global_registration_recipe(b3):
    """The reference's multiway-registration recipe on one pair (test/mini1.py:236-303): voxel grid, statistical outliers,
    normals, FPFH(5 voxels, 100) -> RANSAC on feature matches (ransac_n 4, edge 0.9 + distance checkers) -> point-to-plane ICP
    at 0.4 voxels from the RANSAC pose -> information matrix. The large motion (25 degrees, 30 cm) is out of plain ICP's reach."""
    from b200recon import registration as reg
    from b200recon.geometry import KDTreeSearchParamHybrid
    pts, _ = golden_cloud("output_00094")
    voxel = 0.01
    T = small_rigid(0.35, -0.25, 0.4, (0.3, -0.1, 0.2))
    rng = np.random.default_rng(5)
    clouds, feats = [], []
    for k, P in enumerate((oracle.transform(np.linalg.inv(T), pts)[0] + rng.normal(0, 1e-4, pts.shape), pts)):
        pcd = b3.PointCloud(P).voxel_down_sample(voxel)
        pcd, _ = pcd.remove_statistical_outlier(nb_neighbors=20, std_ratio=2.0)
        pcd.estimate_normals(KDTreeSearchParamHybrid(radius=voxel * 2, max_nn=30))
        clouds.append(pcd)
        feats.append(reg.compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius=voxel * 5, max_nn=100)))
    src, tgt = clouds
    ransac = reg.registration_ransac_based_on_feature_matching(
        src, tgt, feats[0].data, feats[1].data, mutual_filter=False, max_correspondence_distance=voxel * 1.5,
        estimation_method=reg.TransformationEstimationPointToPoint(False), ransac_n=4,
        checkers=[reg.CorrespondenceCheckerBasedOnEdgeLength(0.9), reg.CorrespondenceCheckerBasedOnDistance(voxel * 1.5)],
        criteria=reg.RANSACConvergenceCriteria(max_iteration=4000000, confidence=0.999))
    assert ransac.transformation.any() and ransac.fitness > 0.5
    assert len(ransac.correspondence_set) == round(ransac.fitness * len(src.points))
    icp = reg.registration_icp(src, tgt, voxel * 0.4, ransac.transformation, reg.TransformationEstimationPointToPlane())
    assert rot_err(icp.transformation[:3, :3], T[:3, :3]) < 2e-3 and np.linalg.norm(icp.transformation[:3, 3] - T[:3, 3]) < 2e-3
    info = reg.get_information_matrix_from_point_clouds(src, tgt, voxel * 1.5, icp.transformation)
    assert info.shape == (6, 6) and info[3, 3] == len(reg.evaluate_registration(src, tgt, voxel * 1.5, icp.transformation).correspondence_set)
    # plain ICP from the identity does not get there
    plain = reg.registration_icp(src, tgt, voxel * 1.5, np.eye(4), reg.TransformationEstimationPointToPlane())
    assert rot_err(plain.transformation[:3, :3], T[:3, :3]) > 0.1
    mutual = reg.registration_ransac_based_on_feature_matching(src, tgt, feats[0], feats[1], True, voxel * 1.5, ransac_n=3,
                                                               criteria=reg.RANSACConvergenceCriteria(100000, 0.999))
    assert mutual.fitness > 0.5
    # the other initialisation the reference uses (test/check6.py:236-240): Fast Global Registration, then the same refinement
    fgr = reg.registration_fgr_based_on_feature_matching(src, tgt, feats[0], feats[1],
                                                         reg.FastGlobalRegistrationOption(maximum_correspondence_distance=voxel * 1.5))
    assert fgr.fitness > 0.5 and rot_err(fgr.transformation[:3, :3], T[:3, :3]) < 2e-2
    icp2 = reg.registration_icp(src, tgt, voxel * 0.4, fgr.transformation, reg.TransformationEstimationPointToPlane())
    assert rot_err(icp2.transformation[:3, :3], T[:3, :3]) < 2e-3 and np.linalg.norm(icp2.transformation[:3, 3] - T[:3, 3]) < 2e-3


def test_capture_and_alignment_classes(b3):
    """The reference's scan loop (main.py:34-49) over replayed fixture frames: capture -> align -> accumulate."""
    from b200recon import synth
    cam = SMALL_CAM
    src, tgt, _ = _pairs(2, cam)
    rng = np.random.default_rng(0)
    frames = [(tgt[0], rng.integers(0, 256, (cam["h"], cam["w"], 3), dtype=np.uint8)), (src[0], rng.integers(0, 256, (cam["h"], cam["w"], 3), dtype=np.uint8)),
              (None, None)]
    intr = b3.realsense_pipeline.Intrinsics(cam["w"], cam["h"], cam["fx"], cam["fy"], cam["ppx"], cam["ppy"])
    mgr = b3.RealSensePipeline(source=b3.ReplayPipeline(frames, intr, depth_scale=cam["depth_scale"]))
    mgr.start_pipeline()
    cap = b3.PointCloudCapture(voxel_size=0.02)
    first = cap.capture_point_cloud(mgr.pipeline)
    second = cap.capture_point_cloud(mgr.pipeline)
    assert cap.capture_point_cloud(mgr.pipeline) is None  # missing frame -> None (pointcloud_capture.py:28-29)
    a = (cam["fx"], cam["fy"], cam["ppx"], cam["ppy"], cam["depth_scale"])
    ref = oracle.voxel_tensor(oracle.deproject_z16(tgt[0], *a), 0.02, attr=(frames[0][1].reshape(-1, 3) / 255.0).astype(np.float32))
    assert np.array_equal(np.asarray(first.points), ref["points"].astype(np.float64))
    assert np.array_equal(np.asarray(first.colors), ref["attr"].astype(np.float64))
    combined = b3.PointCloud()
    combined.points, combined.colors = first.points, first.colors
    align = b3.PointCloudAlignment()
    aligned = align.align_point_clouds(second, combined, threshold=0.05, voxel_size=0.03, max_iter=50)
    # oracle chain of pointcloud_alignment.py:22-42
    s = oracle.voxel_legacy(np.asarray(second.points), 0.03)["points"]
    t = oracle.voxel_legacy(np.asarray(combined.points), 0.03)["points"]
    r = oracle.icp(0, s, t, 0.05, max_iter=50)
    assert rot_err(align.last_result.transformation[:3, :3], r["transformation"][:3, :3]) < 1e-5
    assert np.linalg.norm(align.last_result.transformation[:3, 3] - r["transformation"][:3, 3]) < 1e-5
    exp = oracle.transform(r["transformation"], s)[0]
    assert np.abs(np.asarray(aligned.points) - exp).max() < 1e-5 and aligned.has_normals()
    n0 = len(combined.points)
    combined += aligned
    assert len(combined.points) == n0 + len(aligned.points)


def test_processing_and_normal_estimation_classes(b3, tmp_path):
    """main.py:79-80: process_point_cloud(file) then NormalEstimation.estimate_normals."""
    from b200recon import ops, plyio
    pts, nrm = golden_cloud("output84_00060")
    rng = np.random.default_rng(1)
    pcd = b3.PointCloud(pts)
    pcd.colors = rng.random(pts.shape)
    fn = str(tmp_path / "captured_data_on_the_fly.ply")
    plyio.write_point_cloud(fn, pcd)
    proc = b3.PointCloudProcessingWithCUDA(downsample_voxel_size=0.03)
    out = proc.process_point_cloud(fn)
    back = plyio.read_point_cloud(fn)
    p32 = np.asarray(back.points).astype(np.float32)
    v = oracle.voxel_tensor(p32, 0.03, attr=np.asarray(back.colors).astype(np.float32))
    vp = v["points"].astype(np.float64)
    k1, _ = oracle.statistical_outlier(vp, 30, 1.2)
    vp1 = vp[k1]
    # the reference's radius (0.01) is tuned to a 2.5 mm cloud; on this 3 cm cloud it removes everything, like Open3D would
    k2 = oracle.radius_outlier(vp1, 16, 0.01)
    assert len(out.points) == int(k2.sum())
    stat_only, ind = b3.PointCloud(vp).remove_statistical_outlier(30, 1.2)
    assert np.array_equal(np.asarray(stat_only.points), vp1) and ind == np.nonzero(k1)[0].tolist()
    ne = b3.NormalEstimation()
    with_n = ne.estimate_normals(b3.PointCloud(pts))
    ref = oracle.normals_tensor(pts.astype(np.float32), 50, 0.05).astype(np.float64)
    got = np.asarray(with_n.normals)
    d = np.minimum(np.abs(got - ref).max(axis=1), np.abs(got + ref).max(axis=1))  # up to the sign the orientation step chose
    assert np.quantile(d, 0.999) < 1e-4
    # orient_normals_consistent_tangent_plane(100) (normal_estimation.py:21): same flips as the oracle on the same normals
    raw = ops.estimate_normals_tensor(pts.astype(np.float32), 50, 0.05).astype(np.float64)
    want, _ = oracle.orient_normals(pts.astype(np.float32).astype(np.float64), raw, 100)
    assert np.array_equal(got, want)


def test_stepwise_icp_two_shards_equal_single(b3):
    """BASELINE config 5 mechanics on one GPU: the source split in two shards, the 29 sums added (what the all-reduce does),
    every shard applying the same update -> same transform as the unsharded run (to summation-order rounding)."""
    import torch
    from b200recon import distributed as dist, ops
    tgt, nrm = golden_cloud("output_00094")
    T = small_rigid()
    src = oracle.transform(np.linalg.inv(T), tgt)[0]
    single = ops.icp(1, src, tgt, 0.02, tgt_normals=nrm, max_iter=30)
    half = len(src) // 2
    shards = [dist.ShardedICP(1, src[:half], len(src), tgt, 0.02, tgt_normals=nrm, max_iter=30),
              dist.ShardedICP(1, src[half:], len(src), tgt, 0.02, tgt_normals=nrm, max_iter=30)]
    for _ in range(40):
        sums = [s.accumulate() for s in shards]
        total = sums[0] + sums[1]
        for s, t in zip(shards, sums):
            t.copy_(total)
        done = [s.update() for s in shards]
        assert done[0] == done[1]
        if done[0]:
            break
    res = [s.finish() for s in shards]
    assert np.array_equal(res[0]["transformation"], res[1]["transformation"])
    assert rot_err(res[0]["transformation"][:3, :3], single["transformation"][:3, :3]) < 1e-9
    assert np.linalg.norm(res[0]["transformation"][:3, 3] - single["transformation"][:3, 3]) < 1e-9
    assert res[0]["iterations"] == single["iterations"] and abs(res[0]["fitness"] - single["fitness"]) < 1e-12
    corr = np.concatenate([res[0]["corr"], res[1]["corr"]])
    assert np.array_equal(corr, single["corr"])


def test_sharded_icp_million_points_equals_single(b3):
    """BASELINE config 5 at 1.2 M points on one GPU: the height-field cloud cut into three unequal shards, the 29 sums added in
    rank order (what the all-reduce does), every shard applying the same update -> the unsharded registration's correspondence
    set bit for bit and its transform to summation-order rounding; the known motion is recovered."""
    from b200recon import distributed as dist, ops, synth
    src, nrm = synth.height_field_cloud(1100, seed=4001)
    T = synth.rigid(0.0003, -0.0002, 0.0004, (0.0008, -0.0006, 0.001))
    tgt, tn = src @ T[:3, :3].T + T[:3, 3], nrm @ T[:3, :3].T
    n = len(src)
    assert n >= 1_000_000
    kw = dict(tgt_normals=tn, rel_fitness=0.0, rel_rmse=0.0, max_iter=6)
    single = ops.icp(1, src, tgt, 0.005, **kw)
    cuts = [0, n // 5, n // 5 + n // 2, n]
    shards = [dist.ShardedICP(1, src[a:b], n, tgt, 0.005, **kw) for a, b in zip(cuts[:-1], cuts[1:])]
    for _ in range(10):
        sums = [s.accumulate() for s in shards]
        total = (sums[0] + sums[1]) + sums[2]
        for t in sums:
            t.copy_(total)
        done = [s.update() for s in shards]
        assert len(set(done)) == 1
        if done[0]:
            break
    res = [s.finish() for s in shards]
    for r in res[1:]:
        assert np.array_equal(r["transformation"], res[0]["transformation"]) and r["fitness"] == res[0]["fitness"]
    assert res[0]["iterations"] == single["iterations"] == 6
    assert np.array_equal(np.concatenate([r["corr"] for r in res]), single["corr"])
    assert rot_err(res[0]["transformation"][:3, :3], single["transformation"][:3, :3]) < 1e-10
    assert np.linalg.norm(res[0]["transformation"][:3, 3] - single["transformation"][:3, 3]) < 1e-10
    assert abs(res[0]["fitness"] - single["fitness"]) < 1e-12 and abs(res[0]["inlier_rmse"] - single["inlier_rmse"]) < 1e-12
    assert rot_err(single["transformation"][:3, :3], T[:3, :3]) < 1e-5 and np.linalg.norm(single["transformation"][:3, 3] - T[:3, 3]) < 1e-5


def test_fused_pass_single_rank_equals_icp(b3):
    """b3d_icp_pass_peers with a world of one: the pass kernel runs the update itself (one launch per pass, no exchange)
    and must land on exactly the single-call result."""
    import ctypes as C
    import torch
    from b200recon import _native as N, distributed as dist, ops
    tgt, nrm = golden_cloud("output_00094")
    src = oracle.transform(np.linalg.inv(small_rigid()), tgt)[0]
    single = ops.icp(1, src, tgt, 0.02, tgt_normals=nrm, max_iter=30)
    sh = dist.ShardedICP(1, src, len(src), tgt, 0.02, tgt_normals=nrm, max_iter=30)
    buf = torch.zeros(2 * 32 + 2, dtype=torch.float64, device="cuda")
    ptrs = (C.c_void_p * 1)(C.c_void_p(buf.data_ptr()))
    N.check(N.lib().b3d_icp_set_peers(sh.ctx.handle, sh.handle, 0, 1, ptrs))
    res = sh.run_fused(check_every=1)
    assert np.array_equal(res["transformation"], single["transformation"])
    assert res["iterations"] == single["iterations"] and res["fitness"] == single["fitness"]
    assert np.array_equal(res["corr"], single["corr"])


def test_fused_peer_exchange_two_gpus(b3):
    """Config 5 over two GPUs: the all-reduce inside the pass kernel (peer memory) gives the NCCL path's result bit for bit."""
    import json, os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = []
    for k, extra in enumerate(([], ["--fused"])):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(29641 + k), os.path.join(root, "tools", "bench_sharded_icp.py"), "--side", "700"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        out.append(json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]))
    for key in ("fitness", "inlier_rmse", "rot_err_rad", "trans_err_m", "iterations", "passes"):
        assert out[0][key] == out[1][key], (key, out[0], out[1])
    assert out[1]["exchange"].startswith("peer-memory")


def test_reproject_disparity_valid_bit_exact(b3):
    from b200recon import ops, synth
    s, _, Q, _ = synth.disparity_pair(2000, 2001, w=408, h=306, scale=0.425)
    ref = oracle.reproject_disparity(s, Q).reshape(-1, 3)[(s >= 16).reshape(-1)]
    out = ops.reproject_disparity_valid(s, Q, 16)
    assert out.shape == ref.shape and np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert ops.reproject_disparity_valid(np.full((8, 8), -16, np.int16), Q, 16).shape == (0, 3)


def oracle_disparity_pair(ds, dt, Q, voxel, k, radius, kind, dmax, max_iter):
    """CPU restatement of the stereo flavour: reproject valid pixels -> tensor voxel -> legacy normals -> ICP / GICP."""
    xs = oracle.reproject_disparity(ds, Q).reshape(-1, 3)[(ds >= 16).reshape(-1)]
    xt = oracle.reproject_disparity(dt, Q).reshape(-1, 3)[(dt >= 16).reshape(-1)]
    vs = oracle.voxel_tensor(xs, voxel)["points"].astype(np.float64)
    vt = oracle.voxel_tensor(xt, voxel)["points"].astype(np.float64)
    nt = oracle.normals_legacy(vt, k, radius)
    kw = dict(tgt_normals=nt)
    if kind == 2:
        ns = oracle.normals_legacy(vs, k, radius)
        kw = dict(src_cov=oracle.covariances_from_normals(ns).reshape(-1, 9), tgt_cov=oracle.covariances_from_normals(nt).reshape(-1, 9))
    r = oracle.icp(kind, vs, vt, dmax, max_iter=max_iter, **kw)
    r.update(m_source=len(vs), m_target=len(vt), n_raw=len(xs) + len(xt))
    return r


@pytest.mark.parametrize("kind", [2, 1])
def test_disparity_pair_pipeline_vs_oracle(b3, kind):
    """BASELINE config 3 at reduced size (same rig scaled to 408x306): stereo disparity -> cloud -> voxel -> normals -> GICP."""
    from b200recon import ops, synth
    pairs = [synth.disparity_pair(2000 + 2 * i, 2001 + 2 * i, w=408, h=306, scale=0.425) for i in range(2)]
    Q = pairs[0][2]
    ds = np.stack([p[0] for p in pairs])
    dt = np.stack([p[1] for p in pairs])
    voxel, radius, dmax = 0.01, 0.03, 0.04
    params = ops.make_disparity_params(408, 306, Q, 16, voxel_size=voxel, normals_max_nn=30, normals_radius=radius, icp_kind=kind, icp_max_dist=dmax, icp_max_iter=30)
    res = ops.register_disparity_pairs(ds, dt, params)
    for i in range(2):
        ref = oracle_disparity_pair(ds[i], dt[i], Q, voxel, 30, radius, kind, dmax, 30)
        r = res[i]
        assert (r["m_source"], r["m_target"], r["n_raw"]) == (ref["m_source"], ref["m_target"], ref["n_raw"])
        assert rot_err(r["transformation"][:3, :3], ref["transformation"][:3, :3]) < 1e-5
        assert np.linalg.norm(r["transformation"][:3, 3] - ref["transformation"][:3, 3]) < 1e-5
        assert abs(r["fitness"] - ref["fitness"]) < 1e-4 and abs(r["inlier_rmse"] - ref["inlier_rmse"]) < 1e-4


def test_config3_full_size_vs_oracle(b3):
    """BASELINE config 3 at FULL size -- one 3264x2448 disparity pair (8 MP, Q of jetson_stereo_8MP x3.4) -> clouds -> tensor voxel
    5 mm -> hybrid normals on both clouds -> covariances -> generalized ICP -- through b3d_register_disparity_pairs with HOST
    rasters, against the oracle chain on the same rasters (~30 s of host time: the ray-cast and the CPU chain)."""
    from b200recon import ops, synth
    ds, dt, Q, T = synth.disparity_pair(2000, 2001)
    assert ds.shape == (2448, 3264)
    params = ops.make_disparity_params(3264, 2448, Q, 16, icp_kind=2)
    got = ops.register_disparity_pairs(ds, dt, params)[0]
    ref = oracle_disparity_pair(ds, dt, Q, 0.005, 30, 0.01, 2, 0.02, 30)
    assert (got["m_source"], got["m_target"], got["n_raw"]) == (ref["m_source"], ref["m_target"], ref["n_raw"])  # voxel sets: same sizes
    # stated tolerances: 1e-5 rad / 1e-5 m on the transform, 1e-4 on fitness and rmse
    assert rot_err(got["transformation"][:3, :3], ref["transformation"][:3, :3]) < 1e-5
    assert np.linalg.norm(got["transformation"][:3, 3] - ref["transformation"][:3, 3]) < 1e-5
    assert abs(got["fitness"] - ref["fitness"]) < 1e-4 and abs(got["inlier_rmse"] - ref["inlier_rmse"]) < 1e-4
    assert abs(got["iterations"] - ref["iterations"]) <= 1
    # and the registration is right: the synthetic motion is recovered to the sensor's quantisation
    assert rot_err(got["transformation"][:3, :3], T[:3, :3]) < 2e-3 and np.linalg.norm(got["transformation"][:3, 3] - T[:3, 3]) < 5e-3


def test_replay_scan_example(b3, tmp_path):
    """The reference's whole flow (main.py:14-86) on replayed synthetic frames through the reference-facing classes."""
    import importlib.util
    from b200recon import synth
    spec = importlib.util.spec_from_file_location("replay_scan", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "replay_scan.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cam = SMALL_CAM
    rng = np.random.default_rng(0)
    frames = []
    for i in range(3):
        pose = synth.rigid(0.004 * i, -0.003 * i, 0.002 * i, (0.004 * i, 0.0, -0.002 * i))
        frames.append((synth.render_depth(cam["w"], cam["h"], cam["fx"], cam["fy"], cam["ppx"], cam["ppy"], pose=pose, rng=rng),
                       rng.integers(0, 256, (cam["h"], cam["w"], 3), dtype=np.uint8)))
    combined, processed, with_normals = mod.run(frames, cam, str(tmp_path), voxel_size=0.02)
    assert len(combined.points) > 30000 and combined.has_colors()
    assert os.path.isfile(tmp_path / "captured_data_on_the_fly.ply")
    # the reference's fixed filter (16 neighbours within 1 cm, pointcloud_processing.py:39) is tuned to 2.5 mm clouds: on this
    # coarse replay it may remove everything, exactly like Open3D would
    assert len(processed.points) <= len(combined.points)
    ne = b3.NormalEstimation().estimate_normals(combined)
    assert ne.has_normals() and np.allclose(np.linalg.norm(np.asarray(ne.normals), axis=1), 1.0, atol=1e-5)
    if len(processed.points):
        assert with_normals.has_normals()


def test_reference_main_runs_unchanged(b3):
    """the reference's own main.py, byte for byte, runs main.main() to its last line on b200recon (shims first on
    sys.path, replayed camera, Enter fed once the frames are consumed): capture + P2P alignment in the scan thread, the saved PLY,
    the post-processing chain and the normal estimation."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("run_reference_main", os.path.join(root, "tools", "run_reference_main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.find_reference_main() is None:
        pytest.skip("the reference's main.py is neither in /root/reference nor in baseline/_ref")
    rc, text, ply = mod.run(n_frames=3)
    assert rc == 0, text[-3000:]
    assert text.count("Captured point cloud with") == 3 and "Saved point cloud to captured_data_on_the_fly.ply" in text
    assert "Traceback" not in text and ply is not None
    from b200recon import plyio
    assert len(plyio.read_point_cloud(ply).points) > 1000


def test_pair_pipeline_degenerate_frames(b3):
    """A batch mixing a normal pair with an empty (all-zero depth) source frame and an empty target frame: rs.pointcloud keeps
    zero-depth pixels as (0,0,0), so an empty frame is a one-voxel cloud; results must match the oracle chain pair by pair."""
    from b200recon import ops
    src, tgt, _ = _pairs(3, SMALL_CAM)
    src[1][:] = 0
    tgt[2][:] = 0
    params = ops.make_pair_params(**SMALL_CAM, voxel_size=0.02, normals_max_nn=30, normals_radius=0.05, icp_kind=1, icp_max_dist=0.05, icp_max_iter=30)
    res = ops.register_depth_pairs(src, tgt, params)
    for i in range(3):
        ref = oracle_pair(src[i], tgt[i], SMALL_CAM, 0.02, 30, 0.05, 1, 0.05, 30)
        r = res[i]
        assert r["m_source"] == ref["m_source"] and r["m_target"] == ref["m_target"]
        assert abs(r["fitness"] - ref["fitness"]) < 1e-4 and abs(r["inlier_rmse"] - ref["inlier_rmse"]) < 1e-4
        assert np.abs(r["transformation"] - ref["transformation"]).max() < 1e-5
    assert res[1]["m_source"] == 1 and res[2]["m_target"] == 1
