"""Minimal `open3d` for the reference's inference script (diff_completion_pipeline.py:97-99,175,204-212):
`geometry.PointCloud` (points / normals, `farthest_point_down_sample`, `estimate_normals`), `utility.Vector3dVector`,
`io.read_point_cloud` / `io.write_point_cloud` for PLY.  Farthest point sampling runs on the GPU through
lb2_farthest_point_sample (same first-index start and first-argmax tie rule as open3d 0.17)."""
from . import geometry, io, utility  # noqa: F401

__version__ = "0.17.0+lidiff_b200.shim"
