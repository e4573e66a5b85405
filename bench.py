#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json: ICP pairs/s and Mpoints/s for deproject + voxel + normals +
ICP; % of the HBM roofline).

Workload (config.workload): BASELINE config 2 -- RealSense D435 848x480 synthetic depth pairs: rs.pointcloud deprojection,
tensor voxel_down_sample 5 mm, legacy hybrid normals (r = 1 cm, k = 30) on the target, point-to-plane ICP (d_max 2 cm,
<= 30 iterations). One STEP = one batch of --pairs frame pairs per GPU through b3d_register_depth_pairs (config 4 is the
same batch spread over 1/2/4/8 GPUs: weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl reference]

value  : pairs/s, whole job, depth rasters already resident in HBM when the timed region starts
e2e    : pairs/s through the public API (ops.register_depth_pairs) with pinned HOST rasters: H2D of both rasters and the
         D2H of the results are inside the timed region
roofline: the kernel with the largest share of the step, timed with CUDA events on the launching stream
cpu_baseline / --impl reference: the CPU restatement of the reference's Open3D/librealsense path (oracle/, OpenMP, all host
         cores) on a bounded sample of the same workload. Only this file's CPU legs execute oracle/.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture of
# this command at 64 pairs (profiles/, see profiles/README.md for the file the figure comes from)
NCU_TRAFFIC_PER_LAUNCH = {"icp_pass_kernel": 17.628067e9 / 10.0,                  # round 1 (profiles/r01n_ncu_icp_pass_p64_step_digest.txt)
                          "icp_pass2_kernel": (19.127729e9 + 2.125340e9) / 10.0}  # profiles/r02g_ncu_icp_pass2_p64_digest.txt: ten passes of one step (read + write)

WORKLOAD = "config2: D435 848x480 depth pairs -> deproject + tensor voxel 5mm + hybrid normals(0.01,30) + point-to-plane ICP(0.02, 30 it)"
PIPE = dict(voxel_size=0.005, normals_max_nn=30, normals_radius=0.01, icp_kind=1, icp_max_dist=0.02, icp_max_iter=30)


def bench_config(world):
    """The workload description both arms print under "config" (identical dictionaries; per-run details go under "run")."""
    return {"workload": WORKLOAD, "unit_of_work": "one frame pair (two 848x480 depth frames)",
            "parallelism": f"dp{world} (independent pairs per GPU, no data-path collective)",
            "l2": "per-step working set (raw points of the batch) >> 126 MB L2: inputs larger than L2, no flush"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=64, help="frame pairs per GPU per step (64 x 8 GPUs = BASELINE config 4)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--unique", type=int, default=8, help="distinct synthetic pairs rendered per rank (tiled up to --pairs with fresh hole masks)")
    ap.add_argument("--cpu-sample", type=int, default=6, help="pairs in the cpu_baseline sample")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (config 1, config 3, micro, config 5)")
    ap.add_argument("--c5-points", type=int, default=10_000_000, help="config-5 leg: source points per rank")
    return ap.parse_args()


def make_inputs(n_pairs, unique, seed0):
    """[P,h,w] uint16 source / target stacks. Rendering is the slow part on the host, so `unique` scenes are ray-cast and the
    rest of the batch re-uses them with fresh 5 % hole masks (different voxel sets, different ICP problems)."""
    from b200recon import synth
    u = max(1, min(unique, n_pairs))
    src_u, tgt_u, _ = synth.depth_pairs(u, base_seed=seed0)
    src = np.empty((n_pairs,) + src_u.shape[1:], np.uint16)
    tgt = np.empty_like(src)
    rng = np.random.default_rng(seed0 + 77)
    for i in range(n_pairs):
        src[i], tgt[i] = src_u[i % u], tgt_u[i % u]
        if i >= u:
            src[i][rng.random(src[i].shape) < 0.05] = 0
            tgt[i][rng.random(tgt[i].shape) < 0.05] = 0
    return src, tgt


# ---- CPU arm (oracle) --------------------------------------------------------------------------------------------------
def cpu_pair(oracle, s, t, cam):
    a = (cam["fx"], cam["fy"], cam["ppx"], cam["ppy"], cam["depth_scale"])
    vs = oracle.voxel_tensor(oracle.deproject_z16(s, *a), PIPE["voxel_size"])["points"].astype(np.float64)
    vt = oracle.voxel_tensor(oracle.deproject_z16(t, *a), PIPE["voxel_size"])["points"].astype(np.float64)
    nt = oracle.normals_legacy(vt, PIPE["normals_max_nn"], PIPE["normals_radius"])
    return oracle.icp(oracle.P2L, vs, vt, PIPE["icp_max_dist"], tgt_normals=nt, max_iter=PIPE["icp_max_iter"])


def cpu_time_pairs(src, tgt, cam):
    import oracle
    oracle.lib()
    # all the host threads the process may use (torchrun exports OMP_NUM_THREADS=1 to its children)
    try:
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        oracle.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    for i in range(len(src)):
        cpu_pair(oracle, src[i], tgt[i], cam)
    return time.perf_counter() - t0, oracle.num_threads()


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from b200recon import synth
    cam = synth.D435
    per_step = 2
    src, tgt = make_inputs(per_step, per_step, 3000)
    for _ in range(min(args.warmup, 1)):
        cpu_time_pairs(src[:1], tgt[:1], cam)
    t_total, cores = 0.0, 1
    for _ in range(args.steps):
        dt, cores = cpu_time_pairs(src, tgt, cam)
        t_total += dt
    pairs = per_step * args.steps
    v = pairs / t_total
    n_px = cam["w"] * cam["h"]
    sample = f"{per_step} pairs per step x {args.steps} steps of the same workload (oracle C++/OpenMP restatement of the Open3D/librealsense CPU path)"
    print(json.dumps({
        "impl": "reference", "metric": "icp_pairs_per_sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "mpoints_per_sec": v * 2 * n_px / 1e6,
        "config": bench_config(args.gpus),
        "run": {"pairs_per_step": per_step, "note": "bounded sample of the workload; the per-pair CPU cost does not depend on the batch size"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- clocks ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region through NVML in-process (an `nvidia-smi` subprocess
    per sample takes the driver lock for tens of ms and shows up as jitter in short timed regions)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        if nv is None:
            return
        while not self._stop_ev.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            self._stop_ev.wait(0.05)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        if self.nv is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted({n for _, rs in self.rows for n, b in bits.items() if rs & b})
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.max_sm), "reasons": reasons, "samples": len(sm)}


# ---- algorithmic bytes per launch (DESIGN.md "roofline") ---------------------------------------------------
def kernel_bytes(name, N_raw, M_total, Mt, Ms, n_corr, icp_bytes_per_launch=None):
    """Compulsory (algorithmic) HBM traffic of one launch of `name` for a batch with N_raw raw points, M_total voxels (Ms sources,
    Mt targets) -- SURVEY.md 8d. Sort passes and other temporaries are implementation traffic: they carry the bytes their launch
    site declares (profile column 4) under "traffic_gbps" instead."""
    icp = icp_bytes_per_launch if icp_bytes_per_launch is not None else 12 * Ms + 24 * n_corr
    table = {
        "deproject_z16_vec4_kernel": 14 * N_raw / 2,  # two launches per batch (sources, targets)
        "bounds_partial_kernel": 12 * N_raw,
        "cell_key_kernel": (12 * N_raw + 8 * N_raw + 12 * Mt + 8 * Mt) / 2,  # points in, keys out; two launches (voxel lattice, target grid)
        "chunk_key_kernel": 24 * Ms + 8 * Ms,
        "compact_kernel": 8 * N_raw + 4 * M_total,
        "voxel_reduce_short_kernel": 12 * N_raw + 12 * M_total,
        "voxel_reduce_long_kernel": 12 * N_raw * 0.05,
        "widen_kernel": 12 * M_total + 24 * M_total,
        "gather_sorted_kernel": 24 * Mt + 32 * Mt + 16 * Mt,
        "gather_by_sorted_kernel": 24 * Mt + 24 * Mt,
        "chunk_gather_kernel": 24 * Ms + 32 * Ms,
        "normals_kernel": 24 * Mt,  # 12 M in + 12 M out (float32 units); neighbour gathers are cache traffic
        "normals_staged_kernel": 24 * Mt,
        "normals_cov2_kernel": 24 * Mt,   # the stage's compulsory bytes; its 48 M covariance hand-over to the eigen kernel is extra
        "normals_eig2_kernel": 48 * Mt + 24 * Mt,
        # 12 Ns + 24 nC per EXECUTED pass of a pair (finished pairs return at once); averaged over all launches
        "icp_pass_kernel": icp,
        "icp_pass2_kernel": icp,
    }
    return table.get(name)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on the C-level stdout when NCCL_DEBUG is set in the environment; the contract is
        # ONE JSON line on stdout, so fd 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from b200recon import ops, synth
    from b200recon.context import get_context
    cam = synth.D435
    P = args.pairs
    n_px = cam["w"] * cam["h"]
    src, tgt = make_inputs(P, args.unique, 3000 + 1000 * rank)
    params = ops.make_pair_params(**cam, **PIPE)
    ctx = get_context(local_rank)

    # host (pinned) and device copies of the batch
    src_h = torch.from_numpy(src.view(np.int16)).pin_memory()
    tgt_h = torch.from_numpy(tgt.view(np.int16)).pin_memory()
    src_d, tgt_d = src_h.cuda(non_blocking=True), tgt_h.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(ctx.stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    step_dev = lambda: ops.register_depth_pairs(src_d, tgt_d, params, device=local_rank)
    step_e2e = lambda: ops.register_depth_pairs(src_h, tgt_h, params, device=local_rank)

    for _ in range(max(args.warmup, 3)):
        res = step_dev()
    step_e2e()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ctx.launches
    ms_dev, res = timed(step_dev, args.steps)
    launches = ctx.launches - l0
    ms_e2e, res_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None

    # per-kernel CUDA-event timing over the same K steps (profiling brackets every launch with events on the launching stream)
    ctx.profile(True)
    ms_prof, _ = timed(step_dev, args.steps)
    report = ctx.profile_report()
    ctx.profile(False)

    total_pairs = P * world * args.steps
    value = total_pairs / (ms_dev / 1e3)
    e2e_value = total_pairs / (ms_e2e / 1e3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    out = None
    if rank == 0:
        Ms = sum(r["m_source"] for r in res)
        Mt = sum(r["m_target"] for r in res)
        n_corr = sum(r["n_corr"] for r in res)
        N_raw = 2 * n_px * P
        kern_ms = sum(v[1] for v in report.values())
        icp_name = "icp_pass2_kernel" if "icp_pass2_kernel" in report else "icp_pass_kernel"
        icp_launches = max(1, report.get(icp_name, (PIPE["icp_max_iter"] + 1, 0.0, 0))[0] // args.steps)
        icp_bpl = sum((r["iterations"] + 1) * (12 * r["m_source"] + 24 * r["n_corr"]) for r in res) / icp_launches
        kernels = []
        for name, (cnt, ms, declared) in report.items():
            b = kernel_bytes(name, N_raw, Ms + Mt, Mt, Ms, n_corr, icp_bpl)
            kernels.append({"name": name, "launches_per_step": cnt / args.steps, "ms_per_step": ms / args.steps, "share": ms / kern_ms if kern_ms else 0.0,
                            "avg_us": 1e3 * ms / cnt, "gbps": (b / (ms / cnt * 1e-3) / 1e9) if b else None,
                            "traffic_gbps": (declared / (ms * 1e-3) / 1e9) if declared else None})
        top = kernels[0] if kernels else None
        roofline = None
        if top:
            b = kernel_bytes(top["name"], N_raw, Ms + Mt, Mt, Ms, n_corr, icp_bpl)
            ach = top["gbps"] if top["gbps"] is not None else 0.0
            # DRAM traffic per launch of the dominant kernel from the committed `ncu --set full` capture of this command
            # (profiles/r01n_ncu_icp_pass_p64_step_digest.txt: dram__bytes_read.sum + dram__bytes_write.sum over the ten
            # passes of one 64-pair step = 15.94 GB + 1.69 GB, divided by the ten launches -- per launch, like `achieved`)
            traffic = NCU_TRAFFIC_PER_LAUNCH.get(top["name"]) if P == 64 else None
            roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": b, "avg_launch_us": top["avg_us"], "share_of_step": top["share"]}
        # whole-pipeline roofline: compulsory bytes of every stage (SURVEY 8d "pipeline per pair") over the device-timed step
        iters = [r["iterations"] for r in res]
        pipe_bytes = (14 * N_raw + (12 * N_raw + 12 * (Ms + Mt)) + 20 * Mt + 24 * Mt + sum((it + 1) for it in iters) / max(1, len(iters)) * (12 * Ms + 24 * n_corr))
        pipe_gbps = pipe_bytes / (ms_dev / args.steps * 1e-3) / 1e9

        cpu_base = None
        if world == 1:  # the CPU arm is timed on rank 0 at N = 1 only
            cpu_n = max(1, min(args.cpu_sample, P))
            cpu_s, cores = cpu_time_pairs(src[:cpu_n], tgt[:cpu_n], cam)
            cpu_base = {"value": cpu_n / cpu_s, "unit": "pairs/s", "cores": cores, "kind": "port",
                        "sample": f"{cpu_n} pairs of the same batch, oracle C++/OpenMP restatement of the Open3D CPU path, {cpu_s:.1f} s"}
        out = {
            "metric": "icp_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "mpoints_per_sec": value * 2 * n_px / 1e6,
            "config": bench_config(world),
            "run": {"pairs_per_gpu_per_step": P, "global_pairs_per_step": P * world, "raw_points_mb_per_step": N_raw * 12 / 1e6,
                    "avg_icp_iterations": float(np.mean(iters)), "voxels_per_frame": (Ms + Mt) / (2 * P)},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * P * n_px * 2), "d2h_bytes_per_step": int(P * 176),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "pipeline_roofline": {"achieved": pipe_gbps, "peak": peak, "unit": "GB/s", "frac": pipe_gbps / peak, "algorithmic_bytes_per_step": pipe_bytes},
            "cpu_baseline": cpu_base,
            "kernels": kernels[:16],
            "profiled_ms_per_step": ms_prof / args.steps,
        }
    # ---- extra legs, outside the headline's timed region: the other BASELINE configurations ----------------------------------
    extra = {}
    if not args.no_extra:
        from b200recon import benchmarks as B
        src_d = tgt_d = src_h = tgt_h = None  # release the headline batch before the large extra legs
        torch.cuda.empty_cache()
        peak_x = peak if rank == 0 else None
        if world == 1:
            for name, fn in (("config1", lambda: B.config1_leg(local_rank, peak_gbs=peak_x)), ("config3", lambda: B.config3_leg(local_rank, peak_gbs=peak_x)),
                             ("micro_voxel_10m", lambda: B.micro_voxel_leg(local_rank, peak_gbs=peak_x))):
                try:
                    extra[name] = fn()
                except Exception as e:  # a failing extra leg must not take the headline line with it
                    extra[name] = {"error": f"{type(e).__name__}: {e}"}
        try:
            c5 = B.config5_leg(world, rank, local_rank, points_per_rank=args.c5_points, peak_gbs=peak_x)
        except Exception as e:
            c5 = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0:
            extra["config5"] = c5
    if rank == 0:
        out["extra"] = extra
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
