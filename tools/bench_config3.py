#!/usr/bin/env python
"""BASELINE config 3: one 8 MP stereo frame pair (3264x2448 int16 disparity, jetson_stereo_8MP Q scaled x3.4) ->
valid-pixel clouds -> tensor voxel 5 mm -> hybrid normals (1 cm, 30) on both clouds -> covariances -> generalized ICP
(d_max 2 cm) on a single B200. Prints one JSON line: device ms per pair, Mpoints/s, per-kernel table. The full-size comparison with
the CPU oracle lives in tests/test_gpu_pipeline.py (only tests/, smoke() and bench.py's CPU legs may touch oracle/); `bench.py` reports
the same leg as extra.config3."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
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
    print(json.dumps(out))


if __name__ == "__main__":
    main()
