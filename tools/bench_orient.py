"""Times NormalEstimation's two steps (tensor normals, orient_normals_consistent_tangent_plane(100)) on one config-2 frame
(848x480 depth -> 5 mm voxels) and prints the per-kernel breakdown of the orientation step."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from b200recon import ops, synth
from b200recon.context import get_context

cam = synth.D435
depth, _, _ = synth.depth_pair(3000, 3001, cam)
xyz = ops.deproject_z16(depth, cam["fx"], cam["fy"], cam["ppx"], cam["ppy"], cam["depth_scale"])
xyz = xyz[xyz[:, 2] > 0]
pts32 = ops.voxel_down_sample_tensor(xyz, 0.005)["points"]
pts = torch.from_numpy(pts32.astype(np.float64)).cuda()
nrm32 = ops.estimate_normals_tensor(pts32, 50, 0.05, as_tensor=True)
nrm = nrm32.double()
ctx = get_context(0)
for k in (100, 30):
    ops.orient_normals_consistent_tangent_plane(pts, nrm, k, as_tensor=True)  # warm-up
    torch.cuda.synchronize()
    ctx.profile(True)
    t0 = time.perf_counter()
    out, flip = ops.orient_normals_consistent_tangent_plane(pts, nrm, k, as_tensor=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rep = ctx.profile_report()
    ctx.profile(False)
    top = [{"name": n, "launches": c, "ms": round(ms, 3)} for n, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]]
    print(json.dumps({"op": "orient_normals_consistent_tangent_plane", "k": k, "n_points": int(pts.shape[0]), "ms": dt * 1e3,
                      "flipped": int(flip.sum()), "kernels": top}))
