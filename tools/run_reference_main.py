#!/usr/bin/env python
"""Runs the reference's OWN, UNMODIFIED main.py (main.main(): capture -> align -> accumulate in a scan thread, save, post-process,
estimate normals -- /root/reference/main.py:14-86) against b200recon: shims/ first on sys.path, a replayed camera behind
pyrealsense2 ($B3D_REPLAY), Enter fed on stdin once the recorded frames are consumed. SURVEY.md 8f rank 1.

    python tools/run_reference_main.py [--frames 4] [--log profiles/r02_reference_main.log]

main.py is looked up in /root/reference (the build container) and in baseline/_ref/ (the copy __graft_entry__.build() leaves
there so that it travels to the GPU box; git-ignored, never committed). Exit code 0 = main.main() ran to its last line."""
import argparse
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def find_reference_main():
    for d in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(d, "main.py")):
            return d
    return None


def make_replay(path, n_frames):
    """n_frames of the synthetic scene at the reference's stream size (640x480 z16 + bgr8), slowly moving camera."""
    from b200recon import synth
    w, h, fx, fy, ppx, ppy = 640, 480, 616.6348876953125, 616.3090209960938, 312.57867431640625, 242.21949768066406  # test/dataset/realsense/camera_intrinsic.json
    rng = np.random.default_rng(0)
    depth = np.empty((n_frames, h, w), np.uint16)
    color = rng.integers(0, 256, (n_frames, h, w, 3), dtype=np.uint8)
    for i in range(n_frames):
        pose = synth.rigid(0.002 * i, -0.0015 * i, 0.001 * i, (0.002 * i, 0.0, -0.001 * i))
        depth[i] = synth.render_depth(w, h, fx, fy, ppx, ppy, pose=pose, rng=rng)
    np.savez(path, depth=depth, color=color, intrinsics=np.array([fx, fy, ppx, ppy]), depth_scale=0.001)


def run(n_frames=4, log_path=None, timeout=600):
    ref = find_reference_main()
    if ref is None:
        raise FileNotFoundError("reference main.py not found (neither /root/reference nor baseline/_ref)")
    work = tempfile.mkdtemp(prefix="b3d_refmain_")
    replay = os.path.join(work, "frames.npz")
    make_replay(replay, n_frames)
    env = dict(os.environ)
    env["B3D_REPLAY"] = replay
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "3d_reconstruction_project_b200", "shims"), ROOT, env.get("PYTHONPATH", "")])
    env["PYTHONUNBUFFERED"] = "1"
    p = subprocess.Popen([sys.executable, os.path.join(ref, "main.py")], cwd=work, env=env, stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    lines, sent = [], [False]

    def pump():
        for line in p.stdout:
            lines.append(line)
            # all recorded frames consumed: the scan loop now reports empty framesets -> press Enter
            if not sent[0] and "No valid point cloud captured" in line:
                sent[0] = True
                try:
                    p.stdin.write("\n")
                    p.stdin.flush()
                except OSError:
                    pass

    t = threading.Thread(target=pump, daemon=True)
    t.start()
    t0 = time.time()
    while p.poll() is None and time.time() - t0 < timeout:
        time.sleep(0.2)
    if p.poll() is None:
        p.kill()
    t.join(timeout=5)
    # the empty-frame message repeats every 50 ms until Enter arrives: keep the first few
    out, skipped = [], 0
    for line in lines:
        if "No valid point cloud captured" in line:
            skipped += 1
            if skipped > 3:
                continue
        out.append(line)
    text = f"# {ref}/main.py (unmodified) under b200recon shims, {n_frames} replayed frames; rc={p.returncode}; {skipped} empty-frame polls\n" + "".join(out)
    if log_path:
        with open(log_path, "w") as f:
            f.write(text)
    ply = os.path.join(work, "captured_data_on_the_fly.ply")
    return p.returncode, text, ply if os.path.isfile(ply) else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--log", default=None)
    a = ap.parse_args()
    rc, text, ply = run(a.frames, a.log)
    print(text)
    print("saved cloud:", ply)
    sys.exit(0 if rc == 0 and ply else 1)
