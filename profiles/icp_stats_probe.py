"""Debug probe: staged-search statistics of the ICP pass kernel (B3D_ICP_STATS=1) on a small batch of config-2 pairs."""
import ctypes as C
import os
import sys
os.environ["B3D_ICP_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from b200recon import ops, synth, _native as N

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
src, tgt, _ = synth.depth_pairs(P, base_seed=3000)
params = ops.make_pair_params(**synth.D435)
L = N.lib()
L.b3d_debug_icp_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
ops.register_depth_pairs(src, tgt, params)
L.b3d_debug_icp_stats(None, 1)
L.b3d_debug_normals_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
L.b3d_debug_normals_stats(None, 1)
res = ops.register_depth_pairs(src, tgt, params)
torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
L.b3d_debug_icp_stats(out, 1)
c1, c2, ovf, cand, vol, edge = [int(v) for v in out[:6]]
kept_chunks, searched_lanes = int(out[6]), int(out[7])
print(f"chunks visited {kept_chunks + c1}: {kept_chunks} kept every partner without a search ({100.0 * kept_chunks / max(kept_chunks + c1, 1):.1f} %), "
      f"{c1} searched for {searched_lanes / max(c1, 1):.1f} lanes on average")
ns = sum(r["m_source"] for r in res)
its = [r["iterations"] for r in res]
print(f"pairs {P}, source points {ns}, iterations {its}")
print(f"staging calls {c1}, overflows {ovf} ({100.0 * ovf / max(c1, 1):.2f} %)")
print(f"staged candidates per staging call {cand / max(c1 + c2 - ovf, 1):.1f}, mean box volume {vol / max(c1 + c2, 1):.1f} cm^3, mean longest edge {edge / max(c1 + c2, 1) / 100:.2f} cm")
out = (C.c_ulonglong * 8)()
L.b3d_debug_normals_stats(out, 1)
ch, ovf, cut, cand, lanes = [int(v) for v in out[:5]]
print(f"normals: chunks {ch}, overflowed {ovf} ({100.0 * ovf / max(ch, 1):.2f} %), lanes needing the k-nearest cut {cut}, "
      f"candidates per chunk {cand / max(ch - ovf, 1):.1f}, valid lanes per chunk {lanes / max(ch, 1):.1f}")
