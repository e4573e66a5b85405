"""Debug probe: one config-2 frame pair run pass by pass through the step-wise ICP API, with the staged-search counters
(B3D_ICP_STATS=1) and the device time of every pass. Shows where the passes of one registration spend their time."""
import ctypes as C
import os
import sys
os.environ["B3D_ICP_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from b200recon import ops, synth, distributed as D, _native as N

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cam = synth.D435
ds, dt, _ = synth.depth_pair(seed, seed + 1, cam)
clouds = []
for d in (ds, dt):
    xyz = ops.deproject_z16(d, cam["fx"], cam["fy"], cam["ppx"], cam["ppy"], cam["depth_scale"])
    xyz = xyz[xyz[:, 2] > 0]
    v = ops.voxel_down_sample_tensor(xyz, 0.005)["points"]
    clouds.append(np.ascontiguousarray(v, dtype=np.float64))
src, tgt = clouds
nrm = ops.estimate_normals_legacy(tgt, 30, 0.01)
L = N.lib()
L.b3d_debug_icp_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
for rep in range(2):  # the second repetition is the warm one
    sh = D.ShardedICP(1, src, len(src), tgt, 0.02, tgt_normals=nrm, max_iter=30)
    rows = []
    for k in range(40):
        L.b3d_debug_icp_stats(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sh.accumulate()
        e1.record()
        done = sh.update(True)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 8)()
        L.b3d_debug_icp_stats(out, 0)
        rows.append((k, e0.elapsed_time(e1) * 1e3, [int(v) for v in out]))
        if done:
            break
    res = sh.finish()
print(f"source {len(src)}, target {len(tgt)}, iterations {res['iterations']}, fitness {res['fitness']:.4f}, rmse {res['inlier_rmse'] * 1e3:.3f} mm")
print("pass  us     chunks  kept%  lanes/searched-chunk  overflow  cand/staging  multi-batch%  walks/1k-lanes  ties/1k-lanes")
moves = bool(os.environ.get("B3D_PROBE_MOVES"))  # library built with -DB3D_ICP2_STATS_MOVES: counters 1, 4, 5 mean something else
for k, us, o in rows:
    c1, c2, ovf, cand, vol, edge, kept, lanes = o
    if moves:
        print(f"{k:3d} {us:7.1f} us  lanes with a partner {c2:8d}  mean movement since the last search {vol / max(c2, 1):8.1f} um  "
              f"mean room (bound - distance) {edge / max(c2, 1):8.1f} um  searched lanes {lanes}")
        continue
    print(f"{k:3d} {us:7.1f} {kept + c1:7d} {100.0 * kept / max(kept + c1, 1):6.1f} {lanes / max(c1, 1):10.1f} {ovf:8d} "
          f"{cand / max(c1 - ovf, 1):10.1f} {100.0 * c2 / max(c1, 1):10.1f} {1e3 * vol / max(lanes, 1):12.2f} {1e3 * edge / max(lanes, 1):12.2f}")
