"""Shared helpers of the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INTR = json.load(open(os.path.join(GOLDEN, "intrinsics.json")))
# depth_scale = 1.0 / get_depth_scale(), get_depth_scale() a C float 0.001f (reference test/check84.py:158)
DEPTH_SCALE = np.float32(1.0 / float(np.float32(0.001)))


def lexorder(p):
    return np.lexsort((p[:, 2], p[:, 1], p[:, 0]))


def surface_cloud(n, seed=0, noise=0.0005, extent=1.0):
    """A bumpy surface patch with mild noise: the kind of cloud the reference registers (non-degenerate neighbourhoods)."""
    rng = np.random.default_rng(seed)
    xy = (rng.random((n, 2)) - 0.5) * extent
    z = 0.1 * np.sin(4 * xy[:, 0]) * np.cos(3 * xy[:, 1]) + 0.03 * np.sin(9 * xy[:, 0] + 1)
    p = np.column_stack([xy, z]) + rng.normal(0, noise, (n, 3))
    return np.ascontiguousarray(p)


def golden_cloud(name="output_00094"):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return d["ply_points"].copy(), d["ply_normals"].copy()


def small_rigid(rx=0.01, ry=-0.015, rz=0.02, t=(0.004, -0.003, 0.005)):
    ca, sa, cb, sb, cg, sg = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    T = np.eye(4)
    T[:3, :3] = [[cg * cb, cg * sb * sa - sg * ca, cg * sb * ca + sg * sa], [sg * cb, sg * sb * sa + cg * ca, sg * sb * ca - cg * sa],
                 [-sb, cb * sa, cb * ca]]
    T[:3, 3] = t
    return T


def rot_err(Ra, Rb):
    """Angle of Ra^T Rb in radians; chord form (arccos of the trace cannot resolve angles below ~1e-8)."""
    return float(2.0 * np.arcsin(min(1.0, np.linalg.norm(Ra - Rb) / (2.0 * np.sqrt(2.0)))))
