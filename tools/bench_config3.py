#!/usr/bin/env python
"""BASELINE config 3: one 8 MP stereo frame pair (3264x2448 int16 disparity, jetson_stereo_8MP Q scaled x3.4) ->
valid-pixel clouds -> tensor voxel 5 mm -> hybrid normals (1 cm, 30) on both clouds -> covariances -> generalized ICP
(d_max 2 cm) on a single B200. Prints one JSON line: device ms per pair, Mpoints/s, per-kernel table, parity vs the CPU
oracle (optional, --check: slow, ~1-2 min of host time)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--w", type=int, default=3264)
    ap.add_argument("--h", type=int, default=2448)
    ap.add_argument("--kind", type=int, default=2, help="0 point-to-point, 1 point-to-plane, 2 generalized (config 3)")
    a = ap.parse_args()
    import torch
    from b200recon import ops, synth
    from b200recon.context import get_context
    scale = 3.4 * a.w / 3264.0
    ds, dt, Q, T = synth.disparity_pair(2000, 2001, w=a.w, h=a.h, scale=scale)
    params = ops.make_disparity_params(a.w, a.h, Q, 16, icp_kind=a.kind)
    ctx = get_context(0)
    sd, td = torch.from_numpy(ds).cuda(), torch.from_numpy(dt).cuda()
    for _ in range(3):
        res = ops.register_disparity_pairs(sd, td, params)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        res = ops.register_disparity_pairs(sd, td, params)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    ctx.profile(True)
    ops.register_disparity_pairs(sd, td, params)
    rep = ctx.profile_report()
    ctx.profile(False)
    r = res[0]
    rot, tr = synth.transform_error(r["transformation"], T)
    out = {"config": "config3: 8MP stereo pair -> cloud -> voxel 5mm -> normals -> GICP", "w": a.w, "h": a.h, "ms_per_pair": ms, "pairs_per_sec": 1e3 / ms,
           "mpoints_per_sec": r["n_raw"] / (ms * 1e-3) / 1e6, "n_raw": r["n_raw"], "m_source": r["m_source"], "m_target": r["m_target"],
           "iterations": r["iterations"], "fitness": r["fitness"], "inlier_rmse": r["inlier_rmse"], "rot_err_vs_truth_rad": rot, "trans_err_vs_truth_m": tr,
           "kernels": [{"name": k, "launches": v[0], "ms": v[1]} for k, v in list(rep.items())[:10]]}
    if a.check:
        import oracle
        t0 = time.perf_counter()
        xs = oracle.reproject_disparity(ds, Q).reshape(-1, 3)[(ds >= 16).reshape(-1)]
        xt = oracle.reproject_disparity(dt, Q).reshape(-1, 3)[(dt >= 16).reshape(-1)]
        vs = oracle.voxel_tensor(xs, 0.005)["points"].astype(np.float64)
        vt = oracle.voxel_tensor(xt, 0.005)["points"].astype(np.float64)
        ns, nt = oracle.normals_legacy(vs, 30, 0.01), oracle.normals_legacy(vt, 30, 0.01)
        ref = oracle.icp(2, vs, vt, 0.02, src_cov=oracle.covariances_from_normals(ns).reshape(-1, 9), tgt_cov=oracle.covariances_from_normals(nt).reshape(-1, 9))
        cpu_s = time.perf_counter() - t0
        drot, dtr = synth.transform_error(r["transformation"], ref["transformation"])
        out["cpu_oracle"] = {"seconds": cpu_s, "threads": oracle.num_threads(), "rot_diff_rad": drot, "trans_diff_m": dtr, "fitness_diff": abs(r["fitness"] - ref["fitness"]),
                             "rmse_diff": abs(r["inlier_rmse"] - ref["inlier_rmse"]), "voxels_equal": (r["m_source"], r["m_target"]) == (len(vs), len(vt))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
