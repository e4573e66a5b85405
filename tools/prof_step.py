#!/usr/bin/env python
"""One batch of config-2 frame pairs through b3d_register_depth_pairs, nothing else: the short command the ncu captures wrap
(profiles/*_ncu_*). --pairs P frame pairs (rendered from --unique scenes), --steps passes over the same batch."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--unique", type=int, default=4)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--kind", type=int, default=1)
    ap.add_argument("--stats", action="store_true", help="print the staged-search counters (needs B3D_ICP_STATS=1 in the environment)")
    a = ap.parse_args()
    import torch
    import bench
    from b200recon import ops, synth
    src, tgt = bench.make_inputs(a.pairs, a.unique, 3000)
    pipe = dict(bench.PIPE)
    pipe["icp_kind"] = a.kind
    params = ops.make_pair_params(**synth.D435, **pipe)
    sd, td = torch.from_numpy(src.view("int16")).cuda(), torch.from_numpy(tgt.view("int16")).cuda()
    for _ in range(a.steps):
        res = ops.register_depth_pairs(sd, td, params)
    torch.cuda.synchronize()
    print("ok", len(res), res[0]["iterations"], res[0]["fitness"])
    if a.stats:
        import ctypes as C
        from b200recon import _native as N
        for fn, names in (("b3d_debug_icp_stats", ["staged chunks", "multi-batch chunks", "fallback chunks", "candidates", "-", "-", "chunks without a search", "searching lanes"]),
                          ("b3d_debug_normals_stats", ["chunks", "chunks not single-batch", "lanes needing the k cut", "candidates", "valid lanes", "-", "-", "-"])):
            out = (C.c_ulonglong * 8)()
            getattr(N.lib(), fn)(out, 1)
            print(fn, {k: int(v) for k, v in zip(names, out) if k != "-"})


if __name__ == "__main__":
    main()
