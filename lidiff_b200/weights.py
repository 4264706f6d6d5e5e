"""Seeded random parameters with the reference's names and shapes for benchmarks and
smoke runs — the published checkpoints (`diff_net.ckpt`, `refine_net.ckpt`) are unreachable offline.
Kernels follow ME's `reset_parameters` (U(-s,s), s = 1/sqrt(fan)); BN affine/running stats are
randomised around identity so eval-mode BN is a non-trivial per-channel affine.  `calibrate_bn` sets
the running statistics from one forward pass so activations stay O(1) through all 49 layers."""
from __future__ import annotations

import torch

from . import minkunet as mk

_KINDS = {"enc": (mk.MinkGlobalEnc, {}), "diff": (mk.MinkUNetDiff, {}), "refine": (mk.MinkUNet, {"out_channels": 18})}


def random_state_dict(kind: str, seed: int = 0) -> dict:
    cls, kw = _KINDS[kind]
    g = torch.Generator().manual_seed(seed)
    with torch.random.fork_rng(devices=[]):          # CPU generator only (the default would initialise every visible GPU)
        torch.manual_seed(seed)
        net = cls(in_channels=3, **kw)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            c = m.num_features
            m.weight.data = torch.rand(c, generator=g) + 0.5
            m.bias.data = torch.randn(c, generator=g) * 0.1
            m.running_mean.data = torch.randn(c, generator=g) * 0.05
            m.running_var.data = torch.rand(c, generator=g) * 0.5 + 0.75
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


@torch.no_grad()
def calibrate_bn(pipe, scan: torch.Tensor, sigma: float = 0.5, seed: int = 0):
    """One training-mode forward of each network on `scan` (1,N,3) with BN momentum 1 => running stats =
    batch stats of that pass; then eval mode again.  Operator path (ME surface)."""
    g = torch.Generator(device=pipe.device).manual_seed(seed)
    bns = [m for m in pipe.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    for m in bns:
        m.momentum = 1.0
    pipe.train()
    scan = scan.to(pipe.device)
    noisy = scan + sigma * torch.randn(scan.shape, device=pipe.device, generator=g, dtype=scan.dtype)
    x_t, x_c = pipe.points_to_tensor(noisy), pipe.points_to_tensor(scan)
    t = torch.tensor([500], device=pipe.device)
    pipe.model(x_t, x_t.sparse(), pipe.partial_enc(x_c), t)
    pipe.model_refine(pipe.points_to_tensor(scan + 0.05 * torch.randn(scan.shape, device=pipe.device, generator=g, dtype=scan.dtype)))
    pipe.eval()
    for m in bns:
        m.momentum = 0.1
    pipe._engine = None
