#!/usr/bin/env python
"""The reference's whole flow (main.py:14-86) on replayed frames, non-interactive: capture -> align -> accumulate for every
frame, save the cloud, post-process it, estimate normals. Uses only the reference-facing classes of b200recon.

    python examples/replay_scan.py [--frames 4] [--out /tmp/scan]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(frames, cam, out_dir, voxel_size=0.01, align_method="point_to_point"):
    import b200recon as b3
    from b200recon import plyio
    intr = b3.realsense_pipeline.Intrinsics(cam["w"], cam["h"], cam["fx"], cam["fy"], cam["ppx"], cam["ppy"])
    pipeline_manager = b3.RealSensePipeline(source=b3.ReplayPipeline(frames + [(None, None)], intr, depth_scale=cam["depth_scale"]))
    point_cloud_capture = b3.PointCloudCapture(voxel_size=voxel_size)
    point_cloud_alignment = b3.PointCloudAlignment()
    point_cloud_processing = b3.PointCloudProcessingWithCUDA(downsample_voxel_size=voxel_size)
    normal_estimation = b3.NormalEstimation()
    pipeline_manager.start_pipeline()
    combined_pcd = b3.PointCloud()
    for _ in range(len(frames) + 1):  # main.py:34-54
        pcd_frame = point_cloud_capture.capture_point_cloud(pipeline_manager.pipeline)
        if pcd_frame and len(pcd_frame.points) > 0:
            print(f"Captured point cloud with {len(pcd_frame.points)} points.")
            if len(combined_pcd.points) == 0:
                combined_pcd.points = pcd_frame.points
                combined_pcd.colors = pcd_frame.colors
            else:
                aligned = point_cloud_alignment.align_point_clouds(pcd_frame, combined_pcd, threshold=0.05, voxel_size=2 * voxel_size, method=align_method)
                combined_pcd += aligned
        else:
            print("No valid point cloud captured, skipping frame.")
    pipeline_manager.stop_pipeline()
    os.makedirs(out_dir, exist_ok=True)
    fn = os.path.join(out_dir, "captured_data_on_the_fly.ply")
    plyio.write_point_cloud(fn, combined_pcd)  # main.py:72
    pcd = point_cloud_processing.process_point_cloud(fn)  # main.py:79
    pcd_with_normals = normal_estimation.estimate_normals(pcd) if len(pcd.points) else pcd  # main.py:80
    return combined_pcd, pcd, pcd_with_normals


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--out", default="/tmp/b200recon_scan")
    a = ap.parse_args()
    from b200recon import synth
    cam = dict(w=424, h=240, fx=212.0, fy=212.0, ppx=212.0, ppy=120.0, depth_scale=0.001)
    rng = np.random.default_rng(0)
    frames = []
    for i in range(a.frames):
        pose = synth.rigid(0.004 * i, -0.003 * i, 0.002 * i, (0.004 * i, 0.0, -0.002 * i))
        depth = synth.render_depth(cam["w"], cam["h"], cam["fx"], cam["fy"], cam["ppx"], cam["ppy"], pose=pose, rng=rng)
        frames.append((depth, rng.integers(0, 256, (cam["h"], cam["w"], 3), dtype=np.uint8)))
    combined, processed, with_normals = run(frames, cam, a.out)
    print(f"combined {len(combined.points)} points -> processed {len(processed.points)} points, normals: {with_normals.has_normals()}")


if __name__ == "__main__":
    main()
