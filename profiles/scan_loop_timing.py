import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import b200recon as b3
from b200recon import synth
cam = dict(w=640, h=480, fx=616.6, fy=616.3, ppx=312.6, ppy=242.2, depth_scale=0.001)
rng = np.random.default_rng(0)
frames = []
for i in range(6):
    pose = synth.rigid(0.004 * i, -0.003 * i, 0.002 * i, (0.004 * i, 0.0, -0.002 * i))
    frames.append((synth.render_depth(cam["w"], cam["h"], cam["fx"], cam["fy"], cam["ppx"], cam["ppy"], pose=pose, rng=rng), rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)))
intr = b3.realsense_pipeline.Intrinsics(640, 480, cam["fx"], cam["fy"], cam["ppx"], cam["ppy"])
mgr = b3.RealSensePipeline(source=b3.ReplayPipeline(frames, intr, depth_scale=0.001, loop=True)); mgr.start_pipeline()
cap = b3.PointCloudCapture(); al = b3.PointCloudAlignment()
import io, contextlib
combined = b3.PointCloud()
for i in range(6):
    t0 = time.perf_counter(); f = cap.capture_point_cloud(mgr.pipeline); t1 = time.perf_counter()
    if len(combined.points) == 0:
        combined.points, combined.colors = f.points, f.colors; print(f"frame {i}: capture {1e3*(t1-t0):.1f} ms, {len(f.points)} pts"); continue
    with contextlib.redirect_stdout(io.StringIO()):
        a = al.align_point_clouds(f, combined)
    t2 = time.perf_counter(); combined += a; t3 = time.perf_counter()
    print(f"frame {i}: capture {1e3*(t1-t0):.1f} ms, align {1e3*(t2-t1):.1f} ms (iters {al.last_result.iterations}, fitness {al.last_result.fitness:.3f}), += {1e3*(t3-t2):.1f} ms, map {len(combined.points)} pts")
