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
        ties), return the selected points in selection order.  GPU only (lb2_farthest_point_sample); no CPU fallback."""
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
        """PCA normal of the k nearest neighbours (k = 30 as open3d's default KNN search), sign left unoriented.  Brute-force
        k-NN in chunks with torch (on the GPU when there is one); post-processing only, not on the timed path."""
        import torch
        k = getattr(search_param, "knn", None) or getattr(search_param, "max_nn", None) or 30
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        p = torch.as_tensor(np.asarray(self._points), dtype=torch.float32, device=dev)
        n = p.shape[0]
        k = min(k, n)
        out = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        if n >= 3:
            chunk = max(1, min(n, (1 << 26) // max(n, 1)))
            for a in range(0, n, chunk):
                q = p[a:a + chunk]
                idx = torch.cdist(q, p).topk(k, dim=1, largest=False).indices         # (c, k)
                nb = p[idx]                                                           # (c, k, 3)
                c = nb - nb.mean(1, keepdim=True)
                cov = c.transpose(1, 2) @ c
                out[a:a + chunk] = torch.linalg.eigh(cov.double())[1][:, :, 0].float()     # eigenvector of the smallest eigenvalue
        self._normals = Vector3dVector(out.cpu().numpy())
        return True
