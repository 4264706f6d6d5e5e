import numpy as np

from .utility import Vector3dVector


class KDTreeSearchParamKNN:
    def __init__(self, knn=30):
        self.knn = int(knn)


class KDTreeSearchParamHybrid:
    def __init__(self, radius, max_nn):
        self.radius, self.max_nn = float(radius), int(max_nn)


class PointCloud:
    def __init__(self, points=None):
        self._points = Vector3dVector(points if points is not None else ())
        self._normals = Vector3dVector(())
        self._colors = Vector3dVector(())

    points = property(lambda s: s._points, lambda s, v: setattr(s, "_points", Vector3dVector(v)))
    normals = property(lambda s: s._normals, lambda s, v: setattr(s, "_normals", Vector3dVector(v)))
    colors = property(lambda s: s._colors, lambda s, v: setattr(s, "_colors", Vector3dVector(v)))

    def has_points(self):
        return len(self._points) > 0

    def has_normals(self):
        return len(self._normals) == len(self._points) > 0

    def __repr__(self):
        return f"PointCloud with {len(self._points)} points."

    def farthest_point_down_sample(self, num_samples):
        """open3d 0.17 semantics: start at index 0, repeatedly add the point farthest from the selected set (first index on
        ties); like open3d's SelectByIndex the result lists the selected points in ORIGINAL index order.  GPU only
        (lb2_farthest_point_sample); no CPU fallback."""
        import torch
        from lidiff_b200.preprocess import farthest_point_sample
        if not torch.cuda.is_available():
            raise RuntimeError("open3d shim: farthest_point_down_sample needs the lidiff_b200 CUDA library and a GPU")
        n = int(num_samples)
        if n <= 0 or n > len(self._points):
            raise RuntimeError("Illegal number of samples")
        sel = farthest_point_sample(torch.as_tensor(np.asarray(self._points), device="cuda"), n)
        out = PointCloud(np.asarray(self._points)[sel.cpu().numpy()])
        if self.has_normals():
            out.normals = np.asarray(self._normals)[sel.cpu().numpy()]
        return out

    def estimate_normals(self, search_param=None, fast_normal_computation=True):
        """PCA normal of the k nearest neighbours (k = 30 as open3d's default KNN search), sign left unoriented.  Post-processing
        only, not on the timed path.  The k-NN search is bucketed: points are sorted into cells of the k-th-neighbour scale and
        every query chunk (the points of a block of consecutive cells in sorted order) searches only the candidates of its own
        bounding box grown by the current search radius; a chunk whose k-th distance exceeds the margin is redone with a larger one.
        Exact (same neighbours as a brute-force search), O(n * local density) instead of O(n^2) distance evaluations."""
        import torch
        k = getattr(search_param, "knn", None) or getattr(search_param, "max_nn", None) or 30
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        p = torch.as_tensor(np.asarray(self._points), dtype=torch.float32, device=dev)
        n = p.shape[0]
        k = min(k, n)
        out = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        if n >= 3:
            lo, hi = p.min(0).values, p.max(0).values
            vol = float(torch.clamp(hi - lo, min=1e-3).prod())
            cell = max((vol * k / n) ** (1.0 / 3.0), 1e-3)                      # a cell holds ~k points at the mean density
            ijk = torch.floor((p - lo) / cell).long()
            dims = ijk.max(0).values + 1
            key = (ijk[:, 2] * dims[1] + ijk[:, 1]) * dims[0] + ijk[:, 0]
            order = torch.argsort(key)
            ps = p[order]
            chunk = 4096
            for a in range(0, n, chunk):
                q = ps[a:a + chunk]
                qlo, qhi = q.min(0).values, q.max(0).values
                margin = 2.0 * cell
                while True:
                    m = ((ps >= qlo - margin) & (ps <= qhi + margin)).all(1)
                    cand = ps[m]
                    if cand.shape[0] >= k:
                        d, idx = torch.cdist(q, cand).topk(k, dim=1, largest=False)
                        if float(d[:, -1].max()) <= margin or cand.shape[0] == n:    # every k-th neighbour lies inside the searched box
                            break
                    if cand.shape[0] == n:
                        d, idx = torch.cdist(q, cand).topk(k, dim=1, largest=False)
                        break
                    margin *= 2.0
                nb = cand[idx]                                                        # (c, k, 3)
                c = nb - nb.mean(1, keepdim=True)
                cov = c.transpose(1, 2) @ c
                out[order[a:a + chunk]] = torch.linalg.eigh(cov.double())[1][:, :, 0].float()   # eigenvector of the smallest eigenvalue
        self._normals = Vector3dVector(out.cpu().numpy())
        return True
