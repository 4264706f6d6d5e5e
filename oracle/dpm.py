"""Oracle: diffusers-0.18 `DPMSolverMultistepScheduler` as the reference configures it
(tools/diff_completion_pipeline.py:38-46: linear betas, `sde-dpmsolver++`, solver_order=2; library
defaults epsilon prediction, midpoint, lower_order_final, no thresholding).  SURVEY.md App. A.8.

TEST INFRASTRUCTURE — see `oracle/__init__.py`.  diffusers is not installable here; the update
formulas are those of the DPM-Solver++ paper, the closed-form tables are pinned by
tests/golden/known_answers.json.

The fresh Gaussian noise the SDE solver draws each step is passed in explicitly (`noise=`) so that
the CUDA path and this oracle consume identical values.
"""
from __future__ import annotations

import numpy as np
import torch


class DPMSolverSDE2M:
    def __init__(self, num_train_timesteps=1000, beta_start=3.5e-5, beta_end=0.007):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.solver_order = 2
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, n: int):
        ts = np.linspace(0, self.num_train_timesteps - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        _, uniq = np.unique(ts, return_index=True)
        ts = ts[np.sort(uniq)]
        self.timesteps = torch.from_numpy(ts)
        self.num_inference_steps = len(ts)
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0

    # -- epsilon -> x0 -----------------------------------------------------------------------
    def convert_model_output(self, eps, t, sample):
        alpha_t, sigma_t = self.alpha_t[t], self.sigma_t[t]
        return (sample - sigma_t * eps) / alpha_t

    def coefficients(self, step_index: int):
        """fp32 scalars of the update at `step_index` (shared with the CUDA path's host code)."""
        t = self.timesteps[step_index]
        last = step_index == len(self.timesteps) - 1
        t_prev = torch.tensor(0) if last else self.timesteps[step_index + 1]
        lam_t, lam_s = self.lambda_t[t_prev], self.lambda_t[t]
        alpha_tp = self.alpha_t[t_prev]
        sigma_tp, sigma_s = self.sigma_t[t_prev], self.sigma_t[t]
        h = lam_t - lam_s
        c = dict(
            c_sample=sigma_tp / sigma_s * torch.exp(-h),
            c_x0=alpha_tp * (1 - torch.exp(-2.0 * h)),
            c_noise=sigma_tp * torch.sqrt(1.0 - torch.exp(-2.0 * h)),
            sigma_s=sigma_s, alpha_s=self.alpha_t[t], h=h,
        )
        # s1 = timesteps[step_index - 1]; at step_index 0 the index wraps to timesteps[-1] exactly as in diffusers 0.18 when
        # step() is called for a second trajectory without set_timesteps() (the reference never resets between scans,
        # tools/diff_completion_pipeline.py:213-222 -> :155-169)
        s1 = self.timesteps[step_index - 1]
        h0 = lam_s - self.lambda_t[s1]
        c["r0"] = h0 / h
        return c

    def step(self, eps: torch.Tensor, timestep, sample: torch.Tensor, noise: torch.Tensor):
        t = int(timestep)
        hits = (self.timesteps == t).nonzero()
        i = len(self.timesteps) - 1 if len(hits) == 0 else int(hits[0])
        n = len(self.timesteps)
        lower_order_final = (i == n - 1) and n < 15
        x0 = self.convert_model_output(eps, t, sample)
        self.model_outputs[0] = self.model_outputs[1]
        self.model_outputs[1] = x0
        c = self.coefficients(i)
        # diffusers draws the SDE noise in the model output's dtype (fp32): the noise term is an fp32 product, promoted on the add
        if self.lower_order_nums < 1 or lower_order_final:
            prev = c["c_sample"] * sample + c["c_x0"] * x0 + c["c_noise"] * noise
        else:
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            D0, D1 = m0, (1.0 / c["r0"]) * (m0 - m1)
            prev = (c["c_sample"] * sample + c["c_x0"] * D0 + 0.5 * c["c_x0"] * D1 + c["c_noise"] * noise)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        return prev
