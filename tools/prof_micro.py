#!/usr/bin/env python
"""Kernel table (CUDA-event time per launch site) of small single calls: the 10 M-point voxel micro-benchmark
(test/gpu-performance.py), one config-1 registration, one config-3 pair.   python tools/prof_micro.py [voxel|c1|c3]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from b200recon import ops, synth
from b200recon.context import get_context


def table(ctx, run, steps=5):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    ctx.profile(True)
    for _ in range(steps):
        run()
    rep = ctx.profile_report()
    ctx.profile(False)
    tot = sum(v[1] for v in rep.values()) / steps
    print(f"wall {wall:.3f} ms per call, kernels {tot:.3f} ms")
    for name, (cnt, ms, declared) in sorted(rep.items(), key=lambda x: -x[1][1]):
        print(f"  {name:34s} {cnt / steps:5.1f} launches {ms / steps:8.4f} ms")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "voxel"
    ctx = get_context(0)
    if what == "voxel":
        g = torch.Generator(device=ctx.device)
        g.manual_seed(5000)
        pts = torch.rand((10_000_000, 3), device=ctx.device, dtype=torch.float32, generator=g)
        table(ctx, lambda: ops.voxel_down_sample_tensor(pts, 0.05, as_tensor=True))
    elif what == "c3":
        ds, dt, Q, _ = synth.disparity_pair(2000, 2001)
        params = ops.make_disparity_params(ds.shape[1], ds.shape[0], Q, 16, voxel_size=0.005, normals_max_nn=30, normals_radius=0.01, icp_kind=2,
                                           icp_max_dist=0.02, icp_max_iter=30)
        sd, td = torch.from_numpy(ds[None]).cuda(), torch.from_numpy(dt[None]).cuda()
        table(ctx, lambda: ops.register_disparity_pairs(sd, td, params))
        if os.environ.get("B3D_ICP_STATS"):  # staged-normals counters of one call (needs a -DB3D_NRM2_STATS build, B3D_LIB=...)
            import ctypes as C
            from b200recon import _native as N
            L = N.lib()
            L.b3d_debug_normals_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
            L.b3d_debug_normals_stats(None, 1)
            r = ops.register_disparity_pairs(sd, td, params)[0]
            torch.cuda.synchronize()
            out = (C.c_ulonglong * 8)()
            L.b3d_debug_normals_stats(out, 1)
            ch, ovf, cut, cand, lanes = [int(v) for v in out[:5]]
            print(f"normals of both clouds ({r['m_source']} + {r['m_target']} points): stagings {ch}, multi-batch or fallback {ovf}, lanes queued for the "
                  f"k-nearest cut {cut} ({100.0 * cut / max(lanes, 1):.1f} % of {lanes} lanes), candidates per staging {cand / max(ch - ovf, 1):.1f}")
    else:
        import bench
        src, tgt = bench.make_inputs(1, 1, 3000)
        params = ops.make_pair_params(**synth.D435, **bench.PIPE)
        sd, td = torch.from_numpy(src.view("int16")).cuda(), torch.from_numpy(tgt.view("int16")).cuda()
        table(ctx, lambda: ops.register_depth_pairs(sd, td, params))


if __name__ == "__main__":
    main()
