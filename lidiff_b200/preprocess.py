"""`preprocess_scan` pieces on the GPU (/root/reference/lidiff/tools/diff_completion_pipeline.py:92-105).

`farthest_point_sample` replaces open3d's `PointCloud.farthest_point_down_sample` (start at point 0,
repeatedly take the first argmax of the running min squared distance, fp64) and, like open3d's
`SelectByIndex`, returns the selection in ORIGINAL index order.
"""
from __future__ import annotations

import torch

from . import _lib


def farthest_point_sample(points: torch.Tensor, n_samples: int, ordered: bool = True) -> torch.Tensor:
    if not points.is_cuda:
        raise RuntimeError("farthest_point_sample: CUDA tensor required (no CPU fallback)")
    pts = points.to(torch.float64).contiguous()
    n = pts.shape[0]
    if n_samples > n:
        raise RuntimeError(f"farthest_point_sample: asked for {n_samples} of {n} points")
    h = _lib.get_handle(pts.device)
    idx = torch.empty(n_samples, dtype=torch.int32, device=pts.device)
    dist = torch.empty(n, dtype=torch.float64, device=pts.device)
    h.farthest_point_sample(pts, n, n_samples, idx, dist)
    idx = idx.long()
    return torch.sort(idx).values if ordered else idx
