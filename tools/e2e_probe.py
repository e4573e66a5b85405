#!/usr/bin/env python
"""Times the end-to-end call (pinned host rasters -> results) against the device-resident call for one batch of config-2 pairs:
the front end's group count (B3D_E2E_HALVES) is read once per process, so run it once per setting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from b200recon import ops, synth
from b200recon.context import get_context
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src, tgt = bench.make_inputs(P, 8, 3000)
params = ops.make_pair_params(**synth.D435, **bench.PIPE)
ctx = get_context(0)
sh, th = torch.from_numpy(src.view(np.int16)).pin_memory(), torch.from_numpy(tgt.view(np.int16)).pin_memory()
sd, td = sh.cuda(), th.cuda()
def timed(fn, n=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
dev = timed(lambda: ops.register_depth_pairs(sd, td, params))
e2e = timed(lambda: ops.register_depth_pairs(sh, th, params))
cp = timed(lambda: (sh.cuda(non_blocking=True), th.cuda(non_blocking=True)))
print(f"halves={os.environ.get('B3D_E2E_HALVES','2')} device {dev:.2f} ms  e2e {e2e:.2f} ms  gap {e2e-dev:.2f}  plain H2D of both stacks {cp:.2f} ms")
