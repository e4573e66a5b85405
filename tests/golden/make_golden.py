#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the reference's own artefacts.

Run in the BUILD container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs (data files written by a real Open3D run of the reference):
  test/output84/{depth,color,pcd}_NNNNN.{png,png,ply}  <- test/check84.py:139-186
  test/output/{depth,color,pcd}_NNNNN.{png,png,ply}    <- test/mini1.py:132-181
  test/dataset/realsense/camera_intrinsic.json          <- test/generate_intrinsics.py:28-41
  Calib_depth/jetson_stereo_8MP_stereo.npz (Q matrix)   <- Calib_depth/depth4.py:98

Outputs: one compressed .npz per selected frame (depth u16, colour u8 RGB, golden PLY arrays),
intrinsics.json, stereo_Q.npz, and disparity_cv2.npz (cv2.reprojectImageTo3D known-answer vectors).
No reference SOURCE is copied, only data artefacts.
"""
import json
import os
import sys

import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

# (directory, frame, keep colour image?)
FRAMES = [
    ("output84", 8, True),
    ("output84", 60, False),
    ("output", 8, True),
    ("output", 50, False),
    ("output", 94, False),
]


def read_ply(path):
    with open(path, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        dt = np.dtype([("p", "<f8", 3), ("n", "<f8", 3), ("c", "u1", 3)])
        a = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
    return a["p"].copy(), a["n"].copy(), a["c"].copy()


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    for d, fr, keep_color in FRAMES:
        base = os.path.join(REF, "test", d)
        depth = cv2.imread(os.path.join(base, f"depth_{fr:05d}.png"), cv2.IMREAD_UNCHANGED)
        assert depth.dtype == np.uint16 and depth.shape == (480, 640)
        bgr = cv2.imread(os.path.join(base, f"color_{fr:05d}.png"), cv2.IMREAD_COLOR)
        rgb = np.ascontiguousarray(bgr[:, :, ::-1])
        p, n, c = read_ply(os.path.join(base, f"pcd_{fr:05d}.ply"))
        arrs = dict(depth=depth, ply_points=p, ply_normals=n, ply_colors=c)
        if keep_color:
            arrs["color_rgb"] = rgb
        np.savez_compressed(os.path.join(OUT, f"{d}_{fr:05d}.npz"), **arrs)
        print(d, fr, "points", len(p))
    with open(os.path.join(REF, "test/dataset/realsense/camera_intrinsic.json")) as f:
        intr = json.load(f)
    with open(os.path.join(OUT, "intrinsics.json"), "w") as f:
        json.dump(intr, f, indent=1)
    st = np.load(os.path.join(REF, "Calib_depth/jetson_stereo_8MP_stereo.npz"))
    np.savez(os.path.join(OUT, "stereo_Q.npz"), Q=st["Q"], P1=st["P1"], P2=st["P2"])
    # cv2.reprojectImageTo3D known-answer vectors (a3 oracle anchor): float32 disparity in, float32 xyz out
    rng = np.random.default_rng(2000)
    disp16 = rng.integers(16, 128 * 16, size=(60, 80), dtype=np.int16)
    disp16[rng.random(disp16.shape) < 0.03] = -16
    dispf = disp16.astype(np.float32) / np.float32(16.0)
    xyz = cv2.reprojectImageTo3D(dispf, st["Q"].astype(np.float64), handleMissingValues=False)
    np.savez_compressed(os.path.join(OUT, "disparity_cv2.npz"), disp16=disp16, Q=st["Q"], xyz=xyz)
    print("Q=\n", st["Q"])


if __name__ == "__main__":
    main()
