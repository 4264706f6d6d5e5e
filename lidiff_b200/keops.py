"""`pykeops.torch.LazyTensor` surface for the one expression the reference evaluates
(/root/reference/lidiff/models/minkunet.py:412-416):

    f = LazyTensor(full_c[:, None, :]); p = LazyTensor(part_c[None, :, :])
    ((f - p) ** 2).sum(-1).argKmin(1, dim=1)          # -> (N_full, 1) int64

It is evaluated by the exact integer nearest-neighbour CUDA kernel `lb2_nn_match` (ties -> lowest
key index).  Any other symbolic expression raises.
"""
from __future__ import annotations

import torch

from . import _lib, me


class LazyTensor:
    def __init__(self, x: torch.Tensor, _kind=None, _a=None, _b=None):
        if _kind is None:
            if x.dim() != 3 or (x.shape[0] != 1 and x.shape[1] != 1):
                raise RuntimeError("LazyTensor: expected x[:, None, :] or x[None, :, :]")
            # (1,1,D) is both an i- and a j-variable (a single row broadcasts either way), as in KeOps
            self.kind = "var"
            self.can_i, self.can_j = x.shape[1] == 1, x.shape[0] == 1
            self.data = x.reshape(-1, x.shape[-1])
        else:
            self.kind, self.a, self.b = _kind, _a, _b

    def __sub__(self, o):
        if self.kind == "var" and getattr(o, "kind", None) == "var" and self.can_i and o.can_j:
            return LazyTensor(None, "diff", self, o)
        raise RuntimeError("LazyTensor shim: only (x_i - y_j) is supported")

    def __pow__(self, e):
        if self.kind == "diff" and e == 2:
            return LazyTensor(None, "sq", self.a, self.b)
        raise RuntimeError("LazyTensor shim: only (x_i - y_j) ** 2 is supported")

    def sum(self, dim=-1):
        if self.kind == "sq" and dim in (-1, 2):
            return LazyTensor(None, "sqdist", self.a, self.b)
        raise RuntimeError("LazyTensor shim: only ((x_i - y_j) ** 2).sum(-1) is supported")

    def argKmin(self, K, dim=1):
        if self.kind != "sqdist" or K != 1 or dim != 1:
            raise RuntimeError("LazyTensor shim: only sqdist.argKmin(1, dim=1) is supported")
        q, k = self.a.data, self.b.data
        me._require_cuda(q, "LazyTensor operands")
        if q.shape[1] != 4 or k.shape[1] != 4:
            raise RuntimeError("LazyTensor shim: (N,4) [b,x,y,z] coordinates expected")
        h = _lib.get_handle(q.device)
        qi = torch.round(q).to(torch.int32).contiguous()
        ki = torch.round(k).to(torch.int32).contiguous()
        idx = torch.empty(qi.shape[0], dtype=torch.int32, device=q.device)
        # the caller already scaled the batch column by 2*max (minkunet.py:408-410): plain squared L2
        h.nn_match(qi, None, qi.shape[0], ki, None, ki.shape[0], 1, idx)
        return idx.long()[:, None]
