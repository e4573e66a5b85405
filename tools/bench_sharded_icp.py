#!/usr/bin/env python
"""BASELINE config 5: ONE large cloud, point-to-plane ICP sharded by source points over the GPUs of one box, the 29
normal-equation sums all-reduced (NCCL over NVLink) once per pass.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sharded_icp.py --side 3162 [--iters 10]

--side S: the cloud is an S x S height-field grid (S = 10000 -> 1e8 points). Target = T * source with the
C1 transform; every rank holds the whole target (replicated grid) and a contiguous slice of the source. Prints one JSON line
(rank 0): ms per pass (device time, max over ranks), Mpoints/s, the share of the all-reduce, transform error vs truth.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=3162)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dmax", type=float, default=0.005)
    ap.add_argument("--fused", action="store_true", help="all-reduce inside the pass kernel over peer memory (NVLink) instead of NCCL")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from b200recon import distributed as D, ops, synth
    # the cloud is generated on the device (same seed on every rank -> identical replicas of the target); only this rank's
    # slice of the source is kept
    T = synth.rigid(0.0003, -0.0002, 0.0004, (0.0008, -0.0006, 0.001))  # a motion well inside d_max = 5 mm
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev)
    g.manual_seed(4000)
    S = a.side
    u = (torch.arange(S, device=dev, dtype=torch.float64) - S / 2) * 0.001
    x = u.repeat(S) + 0.0002 * torch.randn(S * S, device=dev, dtype=torch.float64, generator=g)
    y = u.repeat_interleave(S) + 0.0002 * torch.randn(S * S, device=dev, dtype=torch.float64, generator=g)
    z = 2.0 + 0.15 * torch.sin(3 * x) * torch.cos(2 * y) + 0.05 * torch.sin(11 * x + 1)
    nx = -(0.45 * torch.cos(3 * x) * torch.cos(2 * y) + 0.55 * torch.cos(11 * x + 1))
    ny = 0.30 * torch.sin(3 * x) * torch.sin(2 * y)
    inv = torch.rsqrt(nx * nx + ny * ny + 1.0)
    src_all = torch.stack([x, y, z], dim=1)
    nrm_all = torch.stack([nx * inv, ny * inv, inv], dim=1)
    del x, y, z, nx, ny, inv
    Tt = torch.from_numpy(T).to(dev)
    tgt = src_all @ Tt[:3, :3].T + Tt[:3, 3]
    tn = nrm_all @ Tt[:3, :3].T
    del nrm_all
    n = S * S
    lo, hi = D.shard_range(n, rank, world)
    src = src_all[lo:hi].clone()
    del src_all
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sh = D.ShardedICP(1, src, n, tgt, a.dmax, tgt_normals=tn, rel_fitness=0.0, rel_rmse=0.0, max_iter=a.iters, device=local)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks, passes = [], 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if a.fused and world > 1:
        sh.enable_peers()
    while True:
        e0, e1, e2 = ev(), ev(), ev()
        look = passes % 2 == 1  # the done flag is read back (host sync) every other pass only
        e0.record()
        if a.fused and world > 1:
            e1.record()
            e2.record()
            done = sh.pass_fused(look)
        else:
            sums = sh.accumulate()
            e1.record()
            D.all_reduce_sums(sums)
            e2.record()
            done = sh.update(look)
        marks.append((e0, e1, e2))
        passes += 1
        if done:
            break
    e_end = ev()
    e_end.record()
    torch.cuda.synchronize()
    acc_ms = sum(a.elapsed_time(b) for a, b, _ in marks)
    red_ms = sum(b.elapsed_time(c) for _, b, c in marks)
    total_ms = marks[0][0].elapsed_time(e_end)
    res = sh.finish()
    t = torch.tensor([acc_ms, red_ms, total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    acc_ms, red_ms, total_ms = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        rot, tr = synth.transform_error(res["transformation"], T)
        per_pass = total_ms / passes  # device time of the whole loop (updates and flag reads included), max over ranks
        fused = a.fused and world > 1
        print(json.dumps({"config": "config5: one cloud sharded by source points, point-to-plane, all-reduce of 29 doubles per pass",
                          "n_points": n, "n_gpus": world, "exchange": ("peer-memory, fused in the pass kernel" if fused else "nccl all_gather + sum in rank order"),
                          "passes": passes, "ms_per_pass": per_pass, "accumulate_ms_per_pass": None if fused else acc_ms / passes,
                          "allreduce_ms_per_pass": None if fused else red_ms / passes,
                          "allreduce_share": None if fused else red_ms / (acc_ms + red_ms),
                          "mpoints_per_sec": n / (per_pass * 1e-3) / 1e6, "setup_s": t_setup, "fitness": res["fitness"],
                          "inlier_rmse": res["inlier_rmse"], "rot_err_rad": rot, "trans_err_m": tr, "iterations": res["iterations"]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
