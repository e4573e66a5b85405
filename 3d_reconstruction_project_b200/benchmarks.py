"""Timed legs of the BASELINE configurations that are not the headline batch (bench.py prints them under "extra"; the tools/
scripts wrap the same functions): config 3 (one 8 MP stereo pair -> cloud -> voxel -> normals -> generalized ICP on one GPU) and
config 5 (one cloud sharded by source points over the ranks, 29-double all-reduce per pass -- NCCL, and fused into the pass
kernel over peer memory). Device timing with CUDA events on the context's stream; algorithmic bytes as SURVEY.md 8(d) states them.
"""
import os
import time

import numpy as np
import torch

from . import distributed as D
from . import ops, synth
from .context import get_context

C3_WORKLOAD = "config3: 8MP stereo disparity pair 3264x2448 (Q of jetson_stereo_8MP x3.4) -> clouds -> tensor voxel 5mm -> hybrid normals(0.01,30) on both -> covariances -> generalized ICP(0.02, 30 it)"
C5_WORKLOAD = "config5: one height-field cloud (1 mm pitch) sharded by source points, point-to-plane ICP(d_max 5 mm, 10 iterations), 29-double all-reduce per pass"


def _timed(ctx, fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ctx.stream)
    out = None
    for _ in range(steps):
        out = fn()
    e1.record(ctx.stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def config3_bytes(r):
    """Compulsory HBM traffic of one config-3 pair: reprojection 2N + 12 Nv, voxel 12 Nv + 12 M, grid 20 M, normals
    24 M and covariances 36 M on both clouds, generalized ICP (12 + 24) Ns + 36 nC per executed pass."""
    n_px, nv, m = r["n_px"], r["n_raw"], r["m_source"] + r["m_target"]
    front = 2 * n_px + 12 * nv + (12 * nv + 12 * m) + 20 * m + 24 * m + 36 * m
    icp = (r["iterations"] + 1) * (36 * r["m_source"] + 36 * r["n_corr"])
    return front + icp, icp / (r["iterations"] + 1)


def config3_leg(device=0, steps=5, warmup=3, w=3264, h=2448, peak_gbs=None, inputs=None):
    """BASELINE config 3 on one GPU. Returns the dict bench.py prints as extra.config3."""
    ctx = get_context(device)
    scale = 3.4 * w / 3264.0
    ds, dt, Q, T_true = inputs if inputs is not None else synth.disparity_pair(2000, 2001, w=w, h=h, scale=scale)
    params = ops.make_disparity_params(w, h, Q, 16, icp_kind=2)
    sh, th = torch.from_numpy(ds).pin_memory(), torch.from_numpy(dt).pin_memory()
    sd, td = sh.to(ctx.device), th.to(ctx.device)
    dev = lambda: ops.register_disparity_pairs(sd, td, params, device=device)
    e2e = lambda: ops.register_disparity_pairs(sh, th, params, device=device)
    for _ in range(max(warmup, 3)):
        res = dev()
    e2e()
    l0 = ctx.launches
    ms_dev, res = _timed(ctx, dev, steps)
    launches = (ctx.launches - l0) // steps
    ms_e2e, _ = _timed(ctx, e2e, steps)
    ctx.profile(True)
    dev()
    rep = ctx.profile_report()
    ctx.profile(False)
    r = dict(res[0])
    r["n_px"] = 2 * w * h
    total_bytes, icp_bytes_per_pass = config3_bytes(r)
    rot, tr = synth.transform_error(r["transformation"], T_true)
    kern_ms = sum(v[1] for v in rep.values()) or 1.0
    kernels = []
    for name, (cnt, ms, _) in list(rep.items())[:10]:
        b = icp_bytes_per_pass if name.startswith("icp_pass") else None
        kernels.append({"name": name, "launches": cnt, "ms": ms, "share": ms / kern_ms, "gbps": (b / (ms / cnt * 1e-3) / 1e9) if b else None})
    gbps = total_bytes / (ms_dev * 1e-3) / 1e9
    out = {"workload": C3_WORKLOAD, "ms_per_pair": ms_dev, "pairs_per_sec": 1e3 / ms_dev, "mpoints_per_sec": r["n_raw"] / (ms_dev * 1e-3) / 1e6,
           "e2e": {"ms_per_pair": ms_e2e, "pairs_per_sec": 1e3 / ms_e2e, "h2d_bytes_per_step": int(2 * w * h * 2), "d2h_bytes_per_step": 176},
           "gpu_launches": int(launches), "n_valid_px": r["n_raw"], "m_source": r["m_source"], "m_target": r["m_target"], "iterations": r["iterations"],
           "fitness": r["fitness"], "inlier_rmse": r["inlier_rmse"], "rot_err_vs_truth_rad": rot, "trans_err_vs_truth_m": tr,
           "roofline": {"bound": "hbm", "scope": "whole pair (every stage's compulsory bytes, SURVEY 8d)", "achieved": gbps, "peak": peak_gbs, "unit": "GB/s",
                        "frac": (gbps / peak_gbs) if peak_gbs else None, "algorithmic_bytes_per_pair": total_bytes},
           "kernels": kernels}
    return out


def _height_field_device(side, dev, seed=4000):
    """side x side points of the synthetic wall at 1 mm pitch + N(0, 0.2 mm) jitter with analytic normals, generated on the device
    (same seed on every rank -> identical replicas)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    S = side
    u = (torch.arange(S, device=dev, dtype=torch.float64) - S / 2) * 0.001
    x = u.repeat(S) + 0.0002 * torch.randn(S * S, device=dev, dtype=torch.float64, generator=g)
    y = u.repeat_interleave(S) + 0.0002 * torch.randn(S * S, device=dev, dtype=torch.float64, generator=g)
    z = 2.0 + 0.15 * torch.sin(3 * x) * torch.cos(2 * y) + 0.05 * torch.sin(11 * x + 1)
    nx = -(0.45 * torch.cos(3 * x) * torch.cos(2 * y) + 0.55 * torch.cos(11 * x + 1))
    ny = 0.30 * torch.sin(3 * x) * torch.sin(2 * y)
    inv = torch.rsqrt(nx * nx + ny * ny + 1.0)
    return torch.stack([x, y, z], dim=1), torch.stack([nx * inv, ny * inv, inv], dim=1)


def balanced_ranges(weights, world):
    """Contiguous slices of len(weights) items with (nearly) equal total weight: [lo_0 = 0, ..., lo_world = n]."""
    c = np.concatenate([[0.0], np.cumsum(np.asarray(weights, dtype=np.float64))])
    cuts = [int(np.searchsorted(c, c[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, len(weights)
    return [min(max(v, 0), len(weights)) for v in cuts]


def config5_run(side, iters, dmax, world, rank, local, fused, T=None):
    """One sharded registration; returns (result dict, device ms of the whole loop max'ed over ranks by the caller, passes)."""
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    if T is None:
        T = synth.rigid(0.0003, -0.0002, 0.0004, (0.0008, -0.0006, 0.001))  # a motion well inside d_max = 5 mm
    src_all, nrm_all = _height_field_device(side, dev)
    Tt = torch.from_numpy(T).to(dev)
    tgt = src_all @ Tt[:3, :3].T + Tt[:3, 3]
    tn = nrm_all @ Tt[:3, :3].T
    del nrm_all
    n = side * side
    lo, hi = D.shard_range(n, rank, world)
    src = src_all[lo:hi].clone()
    del src_all
    torch.cuda.empty_cache()
    sh = D.ShardedICP(1, src, n, tgt, dmax, tgt_normals=tn, rel_fitness=0.0, rel_rmse=0.0, max_iter=iters, device=local)
    ctx = get_context(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    use_fused = fused and world > 1
    if use_fused:
        sh.enable_peers()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1 = ev(), ev()
    passes = 0
    marks = []  # per pass: (before the shard's pass kernel, after it, after the all-reduce)
    e0.record(ctx.stream)
    while True:
        look = passes % 2 == 1  # the done flag is read back (host sync) every other pass only
        if use_fused:
            done = sh.pass_fused(look)
        else:
            a, b, c = ev(), ev(), ev()
            a.record(ctx.stream)
            sums = sh.accumulate()
            b.record(ctx.stream)
            D.all_reduce_sums(sums)
            c.record(ctx.stream)
            marks.append((a, b, c))
            done = sh.update(look)
        passes += 1
        if done:
            break
    e1.record(ctx.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # this rank's own kernel time and its wait in the collective (the wait includes the skew against the slowest shard)
    acc_ms = sum(a.elapsed_time(b) for a, b, _ in marks) / max(1, len(marks))
    red_ms = sum(b.elapsed_time(c) for _, b, c in marks) / max(1, len(marks))
    res = sh.finish()
    del sh, src, tgt, tn
    torch.cuda.empty_cache()
    return res, ms, passes, T, (hi - lo, acc_ms, red_ms)


def config5_leg(world, rank, local, points_per_rank=10_000_000, iters=10, dmax=0.005, peak_gbs=None):
    """BASELINE config 5 at `points_per_rank` source points per rank (weak scaling of the shard, the target is replicated): the
    NCCL all-reduce variant and the fused peer-memory variant back to back; their transforms must be bit-equal. At world = 1
    the same loop without a collective (the N = 1 point of the curve). Returns the dict for extra.config5 (rank 0; None elsewhere)."""
    import torch.distributed as dist
    side = int(np.sqrt(points_per_rank * world))
    n = side * side
    out = None
    runs = {}
    for name, fused in (("nccl", False), ("fused", True)):
        if fused and world == 1:
            continue
        config5_run(side, 2, dmax, world, rank, local, fused)  # warm-up: allocator, symmetric-memory rendezvous, NCCL channels
        res, ms, passes, T, n_local = config5_run(side, iters, dmax, world, rank, local, fused)
        t = torch.zeros((world, 3), dtype=torch.float64, device=torch.device("cuda", local))
        t[rank, 0], t[rank, 1], t[rank, 2] = ms, n_local[1], n_local[2]
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)  # every rank's own times (disjoint rows: a gather)
        per_rank = t.tolist()
        runs[name] = (res, max(r[0] for r in per_rank), passes, T, n_local, per_rank)
    if rank == 0:
        res, ms, passes, T, _, _ = runs["nccl"]
        rot, tr = synth.transform_error(res["transformation"], T)
        bytes_per_pass = 12 * n + 24 * res["n_corr"]  # whole job: every rank's shard (point-to-plane)
        out = {"workload": C5_WORKLOAD, "n_points": n, "points_per_rank": n // world, "n_gpus": world, "passes": passes, "iterations": res["iterations"],
               "fitness": res["fitness"], "inlier_rmse": res["inlier_rmse"], "rot_err_rad": rot, "trans_err_m": tr, "variants": {}}
        for name, (r, ms_v, p_v, _, _, per_rank) in runs.items():
            per_pass = ms_v / p_v
            gb = bytes_per_pass / (per_pass * 1e-3) / 1e9
            out["variants"][name] = {"exchange": "all-reduce inside the pass kernel over peer memory (NVLink)" if name == "fused" else ("nccl all_gather of the 29 sums + sum in rank order" if world > 1 else "none (one rank)"),
                                     "ms_per_pass": per_pass, "mpoints_per_sec": n / (per_pass * 1e-3) / 1e6, "algorithmic_gbps": gb,
                                     "frac_of_n_x_peak": (gb / (peak_gbs * world)) if peak_gbs else None,
                                     # per rank: the whole loop; for the NCCL variant also the shard's own pass kernel and its time in the
                                     # all-reduce per pass (232 bytes: that time is the wait for the slowest shard, i.e. the skew)
                                     "rank_loop_ms_per_pass": [r[0] / p_v for r in per_rank]}
            if name == "nccl":
                out["variants"][name]["rank_pass_kernel_ms"] = [r[1] for r in per_rank]
                out["variants"][name]["rank_allreduce_ms"] = [r[2] for r in per_rank]
        if "fused" in runs:
            a, b = runs["nccl"][0], runs["fused"][0]
            out["fused_equals_nccl_bitwise"] = bool(np.array_equal(a["transformation"], b["transformation"]) and a["fitness"] == b["fitness"]
                                                    and a["inlier_rmse"] == b["inlier_rmse"])
            assert out["fused_equals_nccl_bitwise"], "fused peer-memory exchange and NCCL all-reduce disagree"
    return out


def config1_leg(device=0, steps=10, peak_gbs=None, golden_dir=None):
    """BASELINE config 1 (substitute cloud): point-to-plane ICP of the reference's own fixture cloud
    test/output/pcd_00094.ply (18 449 points, committed under tests/golden/) against a known rigid transform of itself."""
    ctx = get_context(device)
    golden_dir = golden_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    d = np.load(os.path.join(golden_dir, "output_00094.npz"))
    tgt, nrm = d["ply_points"].astype(np.float64), d["ply_normals"].astype(np.float64)
    T = synth.rigid(0.01, -0.015, 0.02, (0.004, -0.003, 0.005))
    Ti = np.linalg.inv(T)
    src = tgt @ Ti[:3, :3].T + Ti[:3, 3]
    sd, td, nd = ctx.to_device(src, torch.float64), ctx.to_device(tgt, torch.float64), ctx.to_device(nrm, torch.float64)
    run = lambda: ops.icp(1, sd, td, 0.02, tgt_normals=nd, max_iter=30, device=device)
    for _ in range(3):
        res = run()
    ms, res = _timed(ctx, run, steps)
    rot, tr = synth.transform_error(res["transformation"], T)
    b = (res["iterations"] + 1) * (12 * len(src) + 24 * res["n_corr"])
    gbps = b / (ms * 1e-3) / 1e9
    return {"workload": "config1: point-to-plane ICP of the fixture cloud pcd_00094 (18449 points) vs a known rigid transform of itself (grid build + all passes per call)",
            "ms_per_registration": ms, "registrations_per_sec": 1e3 / ms, "iterations": res["iterations"], "fitness": res["fitness"],
            "inlier_rmse": res["inlier_rmse"], "rot_err_rad": rot, "trans_err_m": tr,
            "roofline": {"bound": "hbm", "achieved": gbps, "peak": peak_gbs, "unit": "GB/s", "frac": (gbps / peak_gbs) if peak_gbs else None,
                         "note": "18 k points: launch- and latency-bound, not a bandwidth case"}}


def micro_voxel_leg(device=0, steps=20, n=10_000_000, voxel=0.05, peak_gbs=None):
    """The reference's own micro-benchmark shape (test/gpu-performance.py:13-26): 10 M uniform [0,1)^3 float32 points, tensor
    voxel_down_sample(0.05)."""
    ctx = get_context(device)
    g = torch.Generator(device=ctx.device)
    g.manual_seed(5000)
    pts = torch.rand((n, 3), device=ctx.device, dtype=torch.float32, generator=g)
    run = lambda: ops.voxel_down_sample_tensor(pts, voxel, device=device, as_tensor=True)
    for _ in range(3):
        out = run()
    # per-call device times (one event between calls): the leg reports the MEDIAN call; the mean of a 0.7 ms call is at the mercy of a
    # single host hiccup (the mean and the slowest call are reported next to it)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    torch.cuda.synchronize()
    ev[0].record(ctx.stream)
    for k in range(steps):
        out = run()
        ev[k + 1].record(ctx.stream)
    torch.cuda.synchronize()
    per_call = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(steps))
    ms, ms_mean, ms_max = per_call[steps // 2], sum(per_call) / steps, per_call[-1]
    m = int(out["points"].shape[0])
    b = 12 * n + 12 * m
    gbps = b / (ms * 1e-3) / 1e9
    return {"workload": "micro: 10M uniform [0,1)^3 float32 points, tensor voxel_down_sample(0.05) (test/gpu-performance.py:13-26)", "ms": ms,
            "ms_mean": ms_mean, "ms_slowest_call": ms_max, "calls": steps, "mpoints_per_sec": n / (ms * 1e-3) / 1e6, "voxels": m,
            "roofline": {"bound": "hbm", "achieved": gbps, "peak": peak_gbs, "unit": "GB/s", "frac": (gbps / peak_gbs) if peak_gbs else None,
                         "algorithmic_bytes": b}}
