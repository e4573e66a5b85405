"""Per-step wall / device time of the batched pair pipeline (looks for host-side stalls between steps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200recon import ops, synth
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src, tgt = bench.make_inputs(P, 8, 3000)
params = ops.make_pair_params(**synth.D435, **bench.PIPE)
sd, td = torch.from_numpy(src.view(np.int16)).cuda(), torch.from_numpy(tgt.view(np.int16)).cuda()
for _ in range(3):
    ops.register_depth_pairs(sd, td, params)
torch.cuda.synchronize()
rows = []
for k in range(24):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    ops.register_depth_pairs(sd, td, params)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    rows.append((1e3 * (t1 - t0), e0.elapsed_time(e1)))
print("wall ms :", " ".join(f"{r[0]:.0f}" for r in rows))
print("event ms:", " ".join(f"{r[1]:.0f}" for r in rows))
