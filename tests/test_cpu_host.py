"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/b200recon.h declares, the ctypes
signature table covers the header, host logic of the reference-facing classes, PLY I/O, and the multi-process (gloo,
world_size 2) sharding / all-reduce driver."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200recon.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import b200recon
    return b200recon


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(b3d_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    from b200recon import _native
    names = header_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200recon.h but not exported by libb200recon.so"
    assert sorted(_native.SIGNATURES) == names, "ctypes signature table and header disagree"
    assert _native.lib().b3d_version() >= 100


def test_no_cpu_fallback(built):
    """Without a CUDA device the product must fail loudly, never route through the oracle."""
    import torch
    from b200recon import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.voxel_down_sample_legacy(np.zeros((4, 3)), 0.1)
    src = ""
    pkg = os.path.join(ROOT, "3d_reconstruction_project_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src += open(os.path.join(pkg, fn)).read()
    assert "import oracle" not in src and "from oracle" not in src


def test_native_argument_errors_without_gpu(built):
    """Argument validation happens before any CUDA call."""
    from b200recon import _native as N
    L = N.lib()
    m = ctypes.c_int64(0)
    assert L.b3d_voxel_downsample_legacy(None, None, None, None, 0, 0.01, None, None, None, None, None, ctypes.byref(m)) == N.E_INVALID
    assert b"ctx is NULL" in L.b3d_last_error()
    h = ctypes.c_void_p()
    import torch
    if not torch.cuda.is_available():
        assert L.b3d_ctx_create(0, None, ctypes.byref(h)) == N.E_CUDA
        assert b"no CPU fallback" in L.b3d_last_error()


def test_pointcloud_duck_type(built):
    b3 = built
    a = b3.PointCloud(np.arange(12.0).reshape(4, 3))
    assert len(a.points) == 4 and np.asarray(a.points).dtype == np.float64 and not a.has_normals() and bool(a)
    a.colors = b3.Vector3dVector(np.ones((4, 3)))
    b = b3.PointCloud(np.zeros((2, 3)))
    b.colors = np.zeros((2, 3))
    b.normals = np.tile([0.0, 0, 1], (2, 1))
    empty = b3.PointCloud()
    empty.points, empty.colors = a.points, a.colors  # main.py:44-45
    assert len(empty.points) == 4 and empty.has_colors()
    a += b  # main.py:49: normals dropped (a has none), colours kept
    assert len(a.points) == 6 and a.has_colors() and not a.has_normals()
    c = b3.PointCloud()
    c += b
    assert c.has_normals() and c.has_colors() and len(c.points) == 2
    with pytest.raises(RuntimeError):
        b3.Vector3dVector(np.zeros((3, 2)))
    with pytest.raises(RuntimeError, match="voxel_size <= 0"):
        a.voxel_down_sample(0)
    with pytest.raises(RuntimeError, match="Illegal input parameters"):
        a.remove_statistical_outlier(0, 1.0)
    with pytest.raises(RuntimeError, match="Illegal input parameters"):
        a.remove_radius_outlier(1, 0.0)
    assert len(b3.PointCloud().voxel_down_sample(0.01).points) == 0


def test_ply_roundtrip_matches_open3d_layout(built, tmp_path, golden_dir):
    from b200recon import plyio
    d = np.load(os.path.join(golden_dir, "output84_00008.npz"))
    pcd = built.PointCloud(d["ply_points"])
    pcd.normals = d["ply_normals"]
    pcd.colors = d["ply_colors"] / 255.0
    fn = str(tmp_path / "t.ply")
    plyio.write_point_cloud(fn, pcd)
    head = open(fn, "rb").read(400).decode("latin1")
    assert "format binary_little_endian 1.0" in head and "property double nx" in head and "property uchar blue" in head
    back = plyio.read_point_cloud(fn)
    assert np.array_equal(np.asarray(back.points), d["ply_points"]) and np.array_equal(np.asarray(back.normals), d["ply_normals"])
    assert np.array_equal(np.floor(np.asarray(back.colors) * 255 + 0.5).astype(np.uint8), d["ply_colors"])
    fa = str(tmp_path / "a.ply")
    small = built.PointCloud(d["ply_points"][:50])
    plyio.write_point_cloud(fa, small, write_ascii=True)
    assert np.array_equal(np.asarray(plyio.read_point_cloud(fa).points), d["ply_points"][:50])


def test_replay_pipeline_and_realsense_wrapper(built):
    b3 = built
    intr = b3.realsense_pipeline.Intrinsics(4, 3, 2.0, 2.0, 2.0, 1.5)
    frames = [(np.ones((3, 4), np.uint16), np.zeros((3, 4, 3), np.uint8)), (None, None)]
    mgr = b3.RealSensePipeline(source=b3.ReplayPipeline(frames, intr))
    mgr.start_pipeline()
    d, c = mgr.get_frames()
    assert d.dtype == np.uint16 and d.shape == (3, 4) and c.shape == (3, 4, 3)
    with pytest.raises(RuntimeError, match="Failed to capture frames"):
        mgr.get_frames()
    with pytest.raises(RuntimeError, match="Frame didn't arrive"):
        mgr.get_frames()
    mgr.stop_pipeline()
    with pytest.raises(RuntimeError, match="pyrealsense2 is not installed"):
        b3.RealSensePipeline().start_pipeline()


def test_registration_option_classes(built):
    """Host mirror of o3d.pipelines.registration's option objects on the global-registration row (test/mini1.py:269-281,
    test/check6.py:236-240): names, defaults, layout conversion, argument screening -- no device needed."""
    import b200recon
    from b200recon import registration as reg
    o = reg.FastGlobalRegistrationOption()
    assert (o.division_factor, o.use_absolute_scale, o.decrease_mu, o.maximum_correspondence_distance, o.iteration_number, o.tuple_scale,
            o.maximum_tuple_count, o.tuple_test) == (1.4, False, True, 0.025, 64, 0.95, 1000, True)
    c = reg.RANSACConvergenceCriteria()
    assert (c.max_iteration, c.confidence) == (100000, 0.999)
    assert reg._checker_params([reg.CorrespondenceCheckerBasedOnEdgeLength(0.9), reg.CorrespondenceCheckerBasedOnDistance(0.015)]) == (0.9, 0.015)
    assert reg._checker_params(None) == (0.0, 0.0)
    with pytest.raises(RuntimeError, match="only CorrespondenceChecker"):
        reg._checker_params([object()])
    f = reg.Feature(np.arange(66, dtype=np.float64).reshape(33, 2))  # Open3D layout: [dimension, N]
    rows = reg._feature_rows(f)
    assert f.dimension() == 33 and f.num() == 2 and rows.shape == (2, 33) and rows.flags["C_CONTIGUOUS"] and rows[1, 0] == 1.0
    # degenerate requests return the library's empty result before anything touches the device
    pc = b200recon.PointCloud(np.zeros((5, 3)))
    r = reg.registration_ransac_based_on_feature_matching(pc, pc, f, f, False, 0.01, ransac_n=2)
    assert r.fitness == 0.0 and np.array_equal(r.transformation, np.eye(4)) and len(r.correspondence_set) == 0
    # the open3d stand-in exposes the same names under pipelines.registration
    import importlib, sys
    shim_dir = os.path.join(os.path.dirname(b200recon.__file__), "shims")
    sys.path.insert(0, shim_dir)
    try:
        o3d = importlib.import_module("open3d")
        for name in ("registration_icp", "registration_generalized_icp", "registration_ransac_based_on_feature_matching",
                     "registration_fgr_based_on_feature_matching", "FastGlobalRegistrationOption", "compute_fpfh_feature",
                     "get_information_matrix_from_point_clouds", "CorrespondenceCheckerBasedOnEdgeLength", "RANSACConvergenceCriteria"):
            assert hasattr(o3d.pipelines.registration, name), name
    finally:
        sys.path.remove(shim_dir)
        sys.modules.pop("open3d", None)


def test_partitioning(built):
    from b200recon import distributed as dist
    for n, w in ((512, 8), (10, 4), (3, 8), (0, 2)):
        owned = sorted(i for r in range(w) for i in dist.pair_indices(n, r, w))
        assert owned == list(range(n))
        cover = [dist.shard_range(n, r, w) for r in range(w)]
        assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        sizes = [b - a for a, b in cover]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["B3D_ROOT"])
import oracle  # test infrastructure: stands in for the GPU shard on this CPU-only box
from b200recon import distributed as D
sys.path.insert(0, os.path.join(os.environ["B3D_ROOT"], "tests"))
from util import golden_cloud, small_rigid

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["B3D_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()

# config 4: disjoint pair ownership, results gathered on every rank
local = {i: np.full(18, float(i + 1)) for i in D.pair_indices(5, rank, world)}
rows = D.gather_pair_results(local, 5, rank, world)
assert np.array_equal(rows[:, 0], np.arange(1.0, 6.0)), rows[:, 0]

# config 5: the sharded point-to-plane ICP loop with a CPU stand-in shard (same 29-sum contract as b3d_icp_accumulate/update)
tgt, nrm = golden_cloud("output84_00008")
tgt, nrm = tgt[::4].copy(), nrm[::4].copy()
T_true = small_rigid()
src = oracle.transform(np.linalg.inv(T_true), tgt)[0]
lo, hi = D.shard_range(len(src), rank, world)

class Shard:
    def __init__(self, pts):
        self.pts, self.T, self.prev, self.iter, self.result = pts, np.eye(4), None, 0, None
    def accumulate(self):
        p = oracle.transform(self.T, self.pts)[0]
        corr, n, s = oracle.correspondences(p, tgt, None, 0.02)
        m = corr >= 0
        q, nq, pp = tgt[corr[m]], nrm[corr[m]], p[m]
        r = np.sum((pp - q) * nq, axis=1)
        J = np.concatenate([np.cross(pp, nq), nq], axis=1)
        JtJ, Jtr = J.T @ J, J.T @ r
        self.sums = torch.from_numpy(np.concatenate([JtJ[np.triu_indices(6)], Jtr, [float(n), s]]))
        return self.sums
    def update(self):
        a = self.sums.numpy()
        n, fit = a[27], a[27] / len(src)
        rmse = np.sqrt(a[28] / n) if n else 0.0
        if self.prev is not None and abs(self.prev[0] - fit) < 1e-6 and abs(self.prev[1] - rmse) < 1e-6:
            self.result = (fit, rmse); return True
        if self.iter >= 30:
            self.result = (fit, rmse); return True
        A = np.zeros((6, 6)); A[np.triu_indices(6)] = a[:21]; A = A + A.T - np.diag(np.diag(A))
        x = np.linalg.solve(A, -a[21:27])
        U = small_rigid(x[0], x[1], x[2], x[3:6])
        self.T = U @ self.T; self.prev = (fit, rmse); self.iter += 1
        return False

sh = Shard(src[lo:hi])
passes = D.icp_loop(sh.accumulate, sh.update)
ref = oracle.icp(1, src, tgt, 0.02, tgt_normals=nrm, max_iter=30)
assert np.abs(sh.T - ref["transformation"]).max() < 1e-9, np.abs(sh.T - ref["transformation"]).max()
assert sh.iter == ref["iterations"] and abs(sh.result[0] - ref["fitness"]) < 1e-12
# every rank ends with the same transform (same all-reduced sums)
t = torch.from_numpy(sh.T.copy()); g = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(g, t)
assert torch.equal(g[0], g[1])
dist.destroy_process_group()
print("rank", rank, "ok", passes)
'''


def test_two_rank_gloo_sharding(built, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), B3D_ROOT=ROOT, B3D_PORT=port, OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


_ORDER_WORKER = r'''
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["B3D_ROOT"])
from b200recon import distributed as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["B3D_PORT"], rank=int(os.environ["RANK"]), world_size=3)
rank = dist.get_rank()
# three addends whose float64 sum depends on the association: ((1e16 + 1) + -1e16) = 0, ((1e16 + -1e16) + 1) = 1
vals = [[1.0e16, 3.0], [1.0, 1.0e-16], [-1.0e16, 1.0]]
t = torch.tensor(vals[rank] + [float(rank)] * 27, dtype=torch.float64)
D.all_reduce_sums(t)
want0 = (1.0e16 + 1.0) + -1.0e16     # rank order: 0.0 (the 1 is absorbed); r1 + r2 first would give 0.0 as well, r0 + r2 first gives 1.0
want1 = (3.0 + 1.0e-16) + 1.0        # 4.0 in any order at this precision
assert t[0].item() == want0 and t[1].item() == want1 and t[2].item() == 3.0, t[:3]
g = [torch.zeros_like(t) for _ in range(3)]
dist.all_gather(g, t)
assert torch.equal(g[0], g[1]) and torch.equal(g[0], g[2])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_three_rank_sum_in_rank_order(built, tmp_path):
    """all_reduce_sums adds the ranks' vectors as ((r0 + r1) + r2): every rank gets the same bits, and they are the bits of that order
    ((1e16 + 1) - 1e16 = 0, whereas (1e16 - 1e16) + 1 = 1)."""
    script = tmp_path / "order_worker.py"
    script.write_text(_ORDER_WORKER)
    port = str(29900 + os.getpid() % 90)
    procs = []
    for r in range(3):
        env = dict(os.environ, RANK=str(r), B3D_ROOT=ROOT, B3D_PORT=port, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_shims_resolve_reference_main(built):
    """With shims/ first on sys.path the reference's unchanged main.py imports b200recon's classes."""
    ref = "/root/reference"
    if not os.path.isfile(os.path.join(ref, "main.py")):
        pytest.skip("reference tree only exists in the build container")
    shims = os.path.join(ROOT, "3d_reconstruction_project_b200", "shims")
    code = ("import sys; sys.path[:0] = [%r, %r, %r]; import main, b200recon; "
            "assert main.RealSensePipeline is b200recon.RealSensePipeline; assert main.PointCloudCapture is b200recon.PointCloudCapture; "
            "assert main.PointCloudAlignment is b200recon.PointCloudAlignment; assert main.o3d.geometry.PointCloud is b200recon.PointCloud; "
            "import realsense_pipeline, pyrealsense2.pyrealsense2 as rs, pycuda.driver, pycuda.autoinit, cupy; assert cupy.eye(4).get().shape == (4, 4); print('ok')"
            % (shims, ROOT, ref))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT))
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
