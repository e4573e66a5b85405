"""PLY point-cloud I/O (o3d.io.read_point_cloud / write_point_cloud as used at main.py:72 and pointcloud_processing.py:23).

Writer layout = Open3D's binary little-endian layout of the reference's fixtures:
``double x,y,z[,nx,ny,nz]; uchar red,green,blue``, colours quantised as floor(clamp(c,0,1)*255 + 0.5).
Reader: ascii / binary_little_endian vertex elements with float or double coordinates, optional normals and colours.
"""
import numpy as np

from .geometry import PointCloud

_PLY_DTYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4", "double": "f8",
               "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def write_point_cloud(filename, pcd, write_ascii=False, compressed=False, print_progress=False):
    pts = np.asarray(pcd.points, dtype=np.float64).reshape(-1, 3)
    n = len(pts)
    has_n = len(pcd.normals) == n and n > 0
    has_c = len(pcd.colors) == n and n > 0
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if has_n:
        fields += [("nx", "<f8"), ("ny", "<f8"), ("nz", "<f8")]
    if has_c:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    rec = np.zeros(n, dtype=np.dtype(fields))
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if has_n:
        nr = np.asarray(pcd.normals, dtype=np.float64)
        rec["nx"], rec["ny"], rec["nz"] = nr[:, 0], nr[:, 1], nr[:, 2]
    if has_c:
        q = np.floor(np.clip(np.asarray(pcd.colors, dtype=np.float64), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
        rec["red"], rec["green"], rec["blue"] = q[:, 0], q[:, 1], q[:, 2]
    names = {"<f8": "double", "u1": "uchar"}
    header = ["ply", "format ascii 1.0" if write_ascii else "format binary_little_endian 1.0", "comment Created by b200recon",
              f"element vertex {n}"] + [f"property {names[t]} {nm}" for nm, t in fields] + ["end_header"]
    with open(filename, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        if write_ascii:
            for r in rec:
                f.write((" ".join(repr(float(v)) if isinstance(v, np.floating) else str(int(v)) for v in r) + "\n").encode("ascii"))
        else:
            f.write(rec.tobytes())
    return True


def read_point_cloud(filename, format="auto", remove_nan_points=False, remove_infinite_points=False, print_progress=False, device=0):
    with open(filename, "rb") as f:
        if f.readline().strip() != b"ply":
            raise RuntimeError(f"Read PLY failed: {filename} is not a PLY file")
        fmt, n, props, in_vertex, elements = None, 0, [], False, []
        while True:
            line = f.readline()
            if not line:
                raise RuntimeError("Read PLY failed: unexpected end of header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                elements.append(tok[1])
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise RuntimeError("Read PLY failed: list property in vertex element")
                props.append((tok[2], _PLY_DTYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if elements and elements[0] != "vertex":
            raise RuntimeError("Read PLY failed: vertex must be the first element")
        if fmt == "ascii":
            rows = [f.readline().split() for _ in range(n)]
            cols = {nm: np.array([r[i] for r in rows], dtype=np.float64) for i, (nm, _) in enumerate(props)}
        elif fmt in ("binary_little_endian", "binary_big_endian"):
            e = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(nm, e + t if t[1] != "1" else t) for nm, t in props])
            a = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
            cols = {nm: a[nm] for nm, _ in props}
        else:
            raise RuntimeError(f"Read PLY failed: unsupported format {fmt}")
    pcd = PointCloud(device=device)
    if n == 0 or not all(k in cols for k in ("x", "y", "z")):
        return pcd
    pts = np.stack([cols["x"], cols["y"], cols["z"]], axis=1).astype(np.float64)
    keep = np.ones(n, bool)
    if remove_nan_points:
        keep &= ~np.isnan(pts).any(axis=1)
    if remove_infinite_points:
        keep &= ~np.isinf(pts).any(axis=1)
    pcd.points = pts[keep]
    if all(k in cols for k in ("nx", "ny", "nz")):
        pcd.normals = np.stack([cols["nx"], cols["ny"], cols["nz"]], axis=1).astype(np.float64)[keep]
    if all(k in cols for k in ("red", "green", "blue")):
        pcd.colors = (np.stack([cols["red"], cols["green"], cols["blue"]], axis=1).astype(np.float64) / 255.0)[keep]
    return pcd
