"""Synthetic inputs of the named benchmark shapes: an analytic scene (height-field wall + sphere + floor)
ray-cast into uint16 depth rasters for a pinhole camera at two nearby poses. numpy only; deterministic per seed."""
import numpy as np

D435 = dict(w=848, h=480, fx=424.0, fy=424.0, ppx=424.0, ppy=240.0, depth_scale=0.001)


def rot_xyz(rx, ry, rz):
    """R = Rz(rz) Ry(ry) Rx(rx) (the Euler convention of Open3D's 6-vector update)."""
    ca, sa, cb, sb, cg, sg = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    return np.array([[cg * cb, cg * sb * sa - sg * ca, cg * sb * ca + sg * sa],
                     [sg * cb, sg * sb * sa + cg * ca, sg * sb * ca - cg * sa],
                     [-sb, cb * sa, cb * ca]])


def rigid(rx, ry, rz, t):
    T = np.eye(4)
    T[:3, :3] = rot_xyz(rx, ry, rz)
    T[:3, 3] = t
    return T


def _wall(x, y, z0):
    return z0 + 0.15 * np.sin(3.0 * x) * np.cos(2.0 * y) + 0.05 * np.sin(11.0 * x + 1.0)


def render_depth(w, h, fx, fy, ppx, ppy, pose=None, z0=2.0, depth_scale=0.001, zmin=0.5, zmax=4.0, holes=0.05, rng=None, as_metres=False):
    """Depth raster (uint16, units of depth_scale metres) of the scene seen from camera pose ``pose`` (camera -> world, 4x4).
    as_metres=True returns the float64 depth in metres instead (0 = no return)."""
    pose = np.eye(4) if pose is None else pose
    j, i = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d_cam = np.stack([(j - ppx) / fx, (i - ppy) / fy, np.ones_like(j)], axis=-1)
    R, o = pose[:3, :3], pose[:3, 3]
    d = d_cam @ R.T  # world-frame ray directions (not normalised; parameter t == camera-frame depth)
    # wall: Newton iteration on the ray parameter t (== camera-frame depth)
    t = np.full(j.shape, z0)
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    for _ in range(7):
        px, py = o[0] + t * dx, o[1] + t * dy
        s3, c3, s2, c2 = np.sin(3.0 * px), np.cos(3.0 * px), np.sin(2.0 * py), np.cos(2.0 * py)
        g = z0 + 0.15 * s3 * c2 + 0.05 * np.sin(11.0 * px + 1.0)
        gx = 0.45 * c3 * c2 + 0.55 * np.cos(11.0 * px + 1.0)
        gy = -0.30 * s3 * s2
        t = t - (o[2] + t * dz - g) / (dz - gx * dx - gy * dy)
    best = np.where(t > 0, t, np.inf)
    # sphere
    c, r = np.array([0.3, 0.1, z0 - 0.6]), 0.25
    oc = o - c
    a = np.sum(d * d, axis=-1)
    b = 2.0 * (d @ oc)
    cc = oc @ oc - r * r
    disc = b * b - 4 * a * cc
    ts = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
    best = np.minimum(best, np.where(ts > 0, ts, np.inf))
    # floor plane y = 0.8 (y points down in the camera frame)
    with np.errstate(divide="ignore", invalid="ignore"):
        tf = (0.8 - o[1]) / d[..., 1]
    best = np.minimum(best, np.where(np.isfinite(tf) & (tf > 0), tf, np.inf))
    z = np.where((best >= zmin) & (best <= zmax), best, 0.0)
    if as_metres:
        return z
    depth = np.rint(z / depth_scale).astype(np.uint16)
    if holes > 0:
        rng = rng or np.random.default_rng(0)
        depth[rng.random(depth.shape) < holes] = 0
    return depth


def depth_pair(seed_a, seed_b, cam=D435, max_rot_deg=2.0, max_trans=0.02):
    """A (source, target) depth pair: target from the identity pose, source from a pose moved by a random small rigid motion.
    Returns (depth_src, depth_tgt, T_true) with T_true mapping source-camera coordinates onto target-camera coordinates."""
    ra, rb = np.random.default_rng(seed_a), np.random.default_rng(seed_b)
    ang = np.deg2rad(max_rot_deg) * (2 * rb.random(3) - 1)
    tr = max_trans * (2 * rb.random(3) - 1)
    pose_src = rigid(ang[0], ang[1], ang[2], tr)  # source camera -> world (= target camera frame)
    kw = dict(w=cam["w"], h=cam["h"], fx=cam["fx"], fy=cam["fy"], ppx=cam["ppx"], ppy=cam["ppy"], depth_scale=cam["depth_scale"])
    tgt = render_depth(pose=np.eye(4), rng=ra, **kw)
    src = render_depth(pose=pose_src, rng=rb, **kw)
    return src, tgt, pose_src


def depth_pairs(n_pairs, base_seed=3000, cam=D435):
    """BASELINE config 4 inputs: pair i uses seeds (base+2i, base+2i+1). -> (src [P,h,w] u16, tgt [P,h,w] u16, T_true [P,4,4])"""
    src = np.empty((n_pairs, cam["h"], cam["w"]), np.uint16)
    tgt = np.empty_like(src)
    T = np.empty((n_pairs, 4, 4))
    for i in range(n_pairs):
        src[i], tgt[i], T[i] = depth_pair(base_seed + 2 * i, base_seed + 2 * i + 1, cam)
    return src, tgt, T


def height_field_cloud(n_side, pitch=0.001, jitter=0.0002, seed=4000, z0=2.0):
    """BASELINE config 5 shape: n_side^2 points of the wall sampled on a regular grid + jitter, with analytic normals."""
    rng = np.random.default_rng(seed)
    u = (np.arange(n_side) - n_side / 2) * pitch
    x, y = np.meshgrid(u, u)
    x = x + rng.normal(0, jitter, x.shape)
    y = y + rng.normal(0, jitter, y.shape)
    z = _wall(x, y, z0)
    dzdx = 0.45 * np.cos(3 * x) * np.cos(2 * y) + 0.55 * np.cos(11 * x + 1)
    dzdy = -0.30 * np.sin(3 * x) * np.sin(2 * y)
    nrm = np.stack([-dzdx, -dzdy, np.ones_like(z)], axis=-1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    return np.stack([x, y, z], axis=-1).reshape(-1, 3), nrm.reshape(-1, 3)


def rotation_angle(R):
    """Rotation angle in radians; chord form (arccos of the trace cannot resolve angles below ~1e-8)."""
    return float(2.0 * np.arcsin(min(1.0, np.linalg.norm(R - np.eye(3)) / (2.0 * np.sqrt(2.0)))))


def transform_error(Ta, Tb):
    """(rotation angle of Ra^T Rb in rad, |ta - tb| in m)"""
    return rotation_angle(Ta[:3, :3].T @ Tb[:3, :3]), float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))


# ---- stereo (BASELINE config 3) -------------------------------------------------------------------------------------------
def stereo_q(scale=3.4, f=525.60716928, cx=107.42533255, cy=250.54165268, inv_baseline_per_mm=3.17597752e-02):
    """Q of Calib_depth/jetson_stereo_8MP_stereo.npz (960x540 calibration) scaled to `scale` x the resolution, output in
    METRES (the file's Q yields millimetres: last row x 1000). SURVEY.md 8d, config 3."""
    Q = np.zeros((4, 4))
    Q[0, 0] = Q[1, 1] = 1.0
    Q[0, 3], Q[1, 3], Q[2, 3] = -cx * scale, -cy * scale, f * scale
    Q[3, 2] = inv_baseline_per_mm * 1000.0
    return Q


def disparity_pair(seed_a, seed_b, w=3264, h=2448, scale=3.4, invalid=0.03, max_rot_deg=2.0, max_trans=0.02):
    """(disp_src, disp_tgt [h,w] int16 fixed point x16, Q, T_true) for the synthetic scene seen by the scaled stereo rig.
    Invalid pixels (no return, out of the 1..128 px range, or a Bernoulli dropout of `invalid`) carry -16."""
    Q = stereo_q(scale)
    f, cx, cy = Q[2, 3], -Q[0, 3], -Q[1, 3]
    fb = f / Q[3, 2]  # disparity = f * B / Z  (px * m)
    ra, rb = np.random.default_rng(seed_a), np.random.default_rng(seed_b)
    ang = np.deg2rad(max_rot_deg) * (2 * rb.random(3) - 1)
    tr = max_trans * (2 * rb.random(3) - 1)
    pose_src = rigid(ang[0], ang[1], ang[2], tr)
    out = []
    for pose, rng in ((pose_src, rb), (np.eye(4), ra)):
        z = render_depth(w, h, f, f, cx, cy, pose=pose, holes=0.0, as_metres=True)
        with np.errstate(divide="ignore"):
            d = np.where(z > 0, fb / z, 0.0)
        d16 = np.rint(d * 16.0)
        bad = (d16 < 16) | (d16 > 128 * 16) | (rng.random(z.shape) < invalid)
        out.append(np.where(bad, -16, d16).astype(np.int16))
    return out[0], out[1], Q, pose_src
