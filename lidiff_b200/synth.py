"""Synthetic KITTI-shaped LiDAR scans and a PLY reader for the one fixture
the reference ships (`lidiff/Datasets/test/000123.ply`: binary little-endian, double x,y,z).

Used by bench.py / tests to build inputs of the shape BASELINE.json's configs name; there is no
dataset on the GPU box.
"""
from __future__ import annotations

import numpy as np


def synthetic_scan(seed: int = 0, beams: int = 64, azimuths: int = 2048) -> np.ndarray:
    """(~131k, 3) float64 raw scan: 64 beams (+2..-24.8 deg) x 2048 azimuths, sensor 1.73 m above a
    ground plane, walls at a piecewise-constant radius U(8,50) m in 64 sectors, 2 cm noise."""
    g = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(2.0, -24.8, beams))[:, None]
    az = np.linspace(-np.pi, np.pi, azimuths, endpoint=False)[None, :]
    R = g.uniform(8, 50, 64)[(np.arange(azimuths) * 64) // azimuths][None, :]
    with np.errstate(divide="ignore"):
        rg = np.where(el < 0, 1.73 / np.tan(-el), np.inf)
    rh = np.minimum(rg, R)
    z = np.where(rg <= R, -1.73, rh * np.tan(el))
    p = np.stack([rh * np.cos(az), rh * np.sin(az), np.broadcast_to(z, rh.shape)], -1).reshape(-1, 3)
    p = p + g.normal(0, 0.02, p.shape)
    return p


def range_filter(scan: np.ndarray, max_range: float = 50.0, min_range: float = 3.5) -> np.ndarray:
    """tools/diff_completion_pipeline.py:93-94"""
    d = np.sqrt(np.sum(scan ** 2, -1))
    return scan[(d < max_range) & (d > min_range)][:, :3]


def read_ply_xyz(path: str) -> np.ndarray:
    """Minimal binary-little-endian PLY vertex reader (x,y,z float or double)."""
    with open(path, "rb") as f:
        n, props, fmt = 0, [], None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element vertex"):
                n = int(line.split()[-1])
            elif line.startswith("property"):
                props.append(tuple(line.split()[1:3]))
            elif line == "end_header":
                break
        if fmt != "binary_little_endian":
            raise ValueError(f"unsupported PLY format {fmt}")
        tmap = {"double": "<f8", "float": "<f4", "uchar": "u1", "int": "<i4"}
        dt = np.dtype([(name, tmap[t]) for t, name in props])
        raw = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
    return np.stack([raw["x"], raw["y"], raw["z"]], 1).astype(np.float64)
