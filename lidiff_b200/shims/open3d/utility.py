import numpy as np


class Vector3dVector(np.ndarray):
    """(n, 3) float64 array; `np.array(v)` / `np.asarray(v)` give the points back as open3d's does"""

    def __new__(cls, data=()):
        a = np.asarray(data, dtype=np.float64)
        if a.size == 0:
            a = a.reshape(0, 3)
        if a.ndim != 2 or a.shape[1] != 3:
            raise RuntimeError(f"Vector3dVector expects shape (n, 3), got {a.shape}")
        return np.ascontiguousarray(a).view(cls)
