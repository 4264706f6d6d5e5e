"""Oracle: the sampling loop of the reference, CPU restatement.

TEST INFRASTRUCTURE — see `oracle/__init__.py`.

Follows /root/reference/lidiff/tools/diff_completion_pipeline.py
  points_to_tensor :68-84   reset_partial_pcd :86-90   postprocess_scan :107-115
  complete_scan :117-132    forward :140-146           classfree_forward :148-153
  completion_loop :155-169
and the batched float32 twin /root/reference/lidiff/models/models.py:98-103,132-178.
Start noise and the per-step SDE noise are explicit inputs (the reference draws them unseeded).
"""
from __future__ import annotations

import numpy as np
import torch

from . import me_cpu as me
from .dpm import DPMSolverSDE2M
from .nets import Net


class DiffCompletionOracle:
    def __init__(self, sd_enc, sd_diff, sd_refine=None, *, denoising_steps=50, cond_weight=6.0,
                 resolution=0.05, max_range=50.0, div_mode="mul", round_batch_col=True,
                 dtype=torch.float32, t_steps=1000, beta_start=3.5e-5, beta_end=0.007):
        self.enc = Net(sd_enc, dtype)
        self.diff = Net(sd_diff, dtype)
        self.refine = Net(sd_refine, dtype) if sd_refine is not None else None
        self.w_uncond = cond_weight
        self.resolution = resolution
        self.max_range = max_range
        self.div_mode = div_mode
        self.round_batch_col = round_batch_col        # pipeline:72 rounds all 4 columns; models.py:166 cols 1:
        self.dpm = DPMSolverSDE2M(t_steps, beta_start, beta_end)
        self.dpm.set_timesteps(denoising_steps)
        self.trace = {}

    # pipeline:68-84
    def points_to_tensor(self, points: torch.Tensor) -> me.TensorField:
        x_feats = me.batched_coordinates(list(points[:]), dtype=torch.float32)
        x_coord = x_feats.clone()
        if self.round_batch_col:
            x_coord = me.quantize(x_coord, self.resolution, self.div_mode)
        else:
            x_coord[:, 1:] = me.quantize(x_feats[:, 1:], self.resolution, self.div_mode)
        return me.TensorField(features=x_feats[:, 1:], coordinates=x_coord)

    # pipeline:140-146
    def forward(self, x_full, x_full_sparse, x_part, t):
        part_feat = self.enc.global_enc(x_part)
        out = self.diff.unet_diff(x_full, x_full_sparse, part_feat, t)
        return out.reshape(t.shape[0], -1, 3)

    # pipeline:148-153
    def classfree_forward(self, x_t, x_cond, x_uncond, t):
        x_t_sparse = x_t.sparse()
        e_c = self.forward(x_t, x_t_sparse, x_cond, t)
        self.trace["eps_cond"] = e_c
        self.trace["cond_levels"] = dict(self.diff.trace)
        e_u = self.forward(x_t, x_t_sparse, x_uncond, t)
        self.trace["eps_uncond"] = e_u
        self.trace["geom"] = x_t_sparse.geom
        return e_u + self.w_uncond * (e_c - e_u)

    # pipeline:155-169
    def completion_loop(self, x_init, x_t, x_cond, x_uncond, step_noise, n_steps=None, fresh=True):
        """x_init (B,N,3) f64 (pipeline) or f32 (models.py twin); step_noise (T,B,N,3) f32.
        fresh=False keeps the multistep state of the previous trajectory, which is what the reference does for every scan
        after the first (its main loop never calls set_timesteps again)."""
        if fresh:
            self.dpm.set_timesteps(self.dpm.num_inference_steps)
        T = len(self.dpm.timesteps) if n_steps is None else n_steps
        hist = []
        for i in range(T):
            t = self.dpm.timesteps[i][None]
            noise_t = self.classfree_forward(x_t, x_cond, x_uncond, t)
            input_noise = x_t.F.reshape(t.shape[0], -1, 3) - x_init
            x_next = x_init + self.dpm.step(noise_t, t[0], input_noise, step_noise[i])
            hist.append(dict(eps=noise_t, x_next=x_next))
            x_t = self.points_to_tensor(x_next)
            x_cond = self.points_to_tensor(x_cond.F.reshape(t.shape[0], -1, 3))
            x_uncond = self.points_to_tensor(torch.zeros_like(x_cond.F.reshape(t.shape[0], -1, 3)))
        self.trace["hist"] = hist
        return x_t.F.numpy()

    # pipeline:107-115
    def postprocess_scan(self, completed: np.ndarray, input_scan: torch.Tensor) -> np.ndarray:
        dist = np.sqrt(np.sum(completed ** 2, -1))
        post = completed[dist < self.max_range]
        max_z = input_scan[..., 2].max().item()
        min_z = (input_scan[..., 2].mean() - 2 * input_scan[..., 2].std()).item()
        return post[(post[:, 2] < max_z) & (post[:, 2] > min_z)]

    # pipeline:117-132 (after preprocess_scan); `scan` (1,N,3) f64, noises explicit
    def complete_scan(self, scan: torch.Tensor, start_noise: torch.Tensor, step_noise: torch.Tensor, fresh=True):
        x_feats = scan + start_noise
        x_full = self.points_to_tensor(x_feats)
        x_cond = self.points_to_tensor(scan)
        x_uncond = self.points_to_tensor(torch.zeros_like(scan))
        completed = self.completion_loop(scan, x_full, x_cond, x_uncond, step_noise, fresh=fresh)
        post = self.postprocess_scan(completed, scan)
        refine_in = self.points_to_tensor(torch.from_numpy(post)[None, :, :])
        offset = self.refine.unet_refine(refine_in).reshape(-1, 6, 3)
        refined = post[:, None, :] + offset.numpy()
        return refined.reshape(-1, 3), post


# ------------------------------------------------------------------------------------------------
# preprocess_scan (pipeline:92-105): range filter + farthest point sampling (open3d semantics:
# start at index 0, pick argmax of the running min squared distance, first index on ties).
# ------------------------------------------------------------------------------------------------
def farthest_point_sample(points: np.ndarray, n: int) -> np.ndarray:
    pts = np.asarray(points, dtype=np.float64)
    N = pts.shape[0]
    sel = np.empty(n, dtype=np.int64)
    dist = np.full(N, np.inf)
    cur = 0
    for i in range(n):
        sel[i] = cur
        d = ((pts - pts[cur]) ** 2).sum(1)
        np.minimum(dist, d, out=dist)
        cur = int(np.argmax(dist))
    # open3d returns SelectByIndex(selected), which emits points in ORIGINAL index order [open3d-mem]
    return np.sort(sel)


def preprocess_scan(scan: np.ndarray, num_points=180000, max_range=50.0) -> torch.Tensor:
    dist = np.sqrt(np.sum(scan ** 2, -1))
    scan = scan[(dist < max_range) & (dist > 3.5)][:, :3]
    sel = farthest_point_sample(scan, int(num_points / 10))
    s = torch.tensor(scan[sel])
    return s.repeat(10, 1)[None, :, :]


def calibrated_state_dicts(scan: torch.Tensor, seed: int = 0, sigma: float = 0.5, resolution: float = 0.05):
    """Seeded random weights whose BN running stats are set from one oracle forward on `scan`
    (B,N,3), jittered so BN is not an identity and activations stay O(1) through all 49 layers
    (the published checkpoints are unreachable)."""
    from .nets import random_state_dict
    g = torch.Generator().manual_seed(seed + 17)
    sds = dict(enc=random_state_dict("enc", seed), diff=random_state_dict("diff", seed + 1),
               refine=random_state_dict("refine", seed + 2, out_channels=18))
    o = DiffCompletionOracle(sds["enc"], sds["diff"], sds["refine"], resolution=resolution)
    for n in (o.enc, o.diff, o.refine):
        n.calibrate, n.rng = True, g
    noisy = scan + sigma * torch.randn(scan.shape, generator=g, dtype=scan.dtype)
    x_t, x_c = o.points_to_tensor(noisy), o.points_to_tensor(scan)
    o.forward(x_t, x_t.sparse(), x_c, torch.tensor([500] * scan.shape[0]))
    o.refine.unet_refine(o.points_to_tensor(scan + 0.05 * torch.randn(scan.shape, generator=g, dtype=scan.dtype)))
    return sds
