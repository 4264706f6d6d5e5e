"""PLY point-cloud I/O (vertex x y z [nx ny nz]); ascii and binary_little_endian, float or double properties."""
import numpy as np

from .geometry import PointCloud

_DT = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
       "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2", "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def read_point_cloud(filename, format="auto", **_):
    with open(filename, "rb") as f:
        if f.readline().strip() != b"ply":
            raise RuntimeError(f"read_point_cloud: {filename} is not a PLY file (the shim reads PLY only)")
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise RuntimeError("read_point_cloud: unterminated PLY header")
            t = line.decode("ascii", "replace").split()
            if not t:
                continue
            if t[0] == "format":
                fmt = t[1]
            elif t[0] == "element":
                in_vertex = t[1] == "vertex"
                if in_vertex:
                    n = int(t[2])
            elif t[0] == "property" and in_vertex:
                if t[1] == "list":
                    raise RuntimeError("read_point_cloud: list properties on vertices are not supported")
                props.append((t[2], _DT[t[1]]))
            elif t[0] == "end_header":
                break
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=n, ndmin=2) if n else np.zeros((0, len(props)))
            cols = {name: data[:, i] for i, (name, _) in enumerate(props)}
        elif fmt == "binary_little_endian":
            rec = np.frombuffer(f.read(n * np.dtype(props).itemsize), dtype=np.dtype(props), count=n)
            cols = {name: rec[name] for name, _ in props}
        else:
            raise RuntimeError(f"read_point_cloud: PLY format '{fmt}' not supported")
    pcd = PointCloud(np.stack([cols["x"], cols["y"], cols["z"]], 1) if n else None)
    if n and all(k in cols for k in ("nx", "ny", "nz")):
        pcd.normals = np.stack([cols["nx"], cols["ny"], cols["nz"]], 1)
    return pcd


def write_point_cloud(filename, pointcloud, write_ascii=False, compressed=False, print_progress=False):
    pts = np.asarray(pointcloud.points, dtype=np.float64)
    has_n = pointcloud.has_normals()
    cols = [pts] + ([np.asarray(pointcloud.normals, dtype=np.float64)] if has_n else [])
    names = ["x", "y", "z"] + (["nx", "ny", "nz"] if has_n else [])
    data = np.concatenate(cols, 1) if len(pts) else np.zeros((0, len(names)))
    hdr = "ply\nformat {} 1.0\ncomment Created by lidiff_b200 (open3d shim)\nelement vertex {}\n".format(
        "ascii" if write_ascii else "binary_little_endian", len(pts)) + "".join(f"property double {n}\n" for n in names) + "end_header\n"
    with open(filename, "wb") as f:
        f.write(hdr.encode("ascii"))
        if write_ascii:
            np.savetxt(f, data, fmt="%.10g")
        else:
            f.write(np.ascontiguousarray(data, dtype="<f8").tobytes())
    return True
