"""`diffusers.DPMSolverMultistepScheduler` surface as the reference configures it
(/root/reference/lidiff/tools/diff_completion_pipeline.py:38-47,58-66,163): linear betas,
`sde-dpmsolver++`, solver_order 2, epsilon prediction, midpoint, lower_order_final.

`step()` is the operator-level path (torch elementwise ops, same dtype promotion as diffusers);
`coefficients()` feeds the fused CUDA tail kernel `lb2_guidance_dpm_step` used by the engine.
The SDE noise can be injected (`noise=`) so runs are reproducible; by default it is drawn with
torch.randn like diffusers does.
"""
from __future__ import annotations

import numpy as np
import torch


class DPMSolverMultistepScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 solver_order=2, prediction_type="epsilon", algorithm_type="dpmsolver++", solver_type="midpoint",
                 lower_order_final=True, thresholding=False, **unused):
        if beta_schedule != "linear" or algorithm_type != "sde-dpmsolver++" or solver_order != 2 \
                or prediction_type != "epsilon" or solver_type != "midpoint" or thresholding:
            raise NotImplementedError("only the configuration the LiDiff pipeline uses is implemented: linear betas, "
                                      "sde-dpmsolver++, order 2, epsilon prediction, midpoint")
        self.num_train_timesteps = num_train_timesteps
        self.solver_order = solver_order
        self.lower_order_final = lower_order_final
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5   # re-assigned at pipeline:66, unused here
        self.init_noise_sigma = 1.0
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, num_inference_steps, device=None):
        ts = np.linspace(0, self.num_train_timesteps - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        _, first = np.unique(ts, return_index=True)
        ts = ts[np.sort(first)]
        self.timesteps = torch.from_numpy(ts)
        if device is not None:
            self.timesteps = self.timesteps.to(device)
        self.num_inference_steps = len(ts)
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0

    # ------------------------------------------------------------------------------------------------
    def _cpu_tables(self):
        return (self.lambda_t.detach().cpu(), self.alpha_t.detach().cpu(), self.sigma_t.detach().cpu(),
                self.timesteps.detach().cpu())

    def coefficients(self, step_index: int) -> dict:
        """fp32 scalars of the update at `step_index`, evaluated with torch fp32 ops in the same
        order as diffusers' sde-dpmsolver++ branches."""
        lam, alpha, sigma, ts = self._cpu_tables()
        t = ts[step_index]
        last = step_index == len(ts) - 1
        t_prev = torch.tensor(0) if last else ts[step_index + 1]
        h = lam[t_prev] - lam[t]
        c = dict(c_sample=sigma[t_prev] / sigma[t] * torch.exp(-h),
                 c_x0=alpha[t_prev] * (1 - torch.exp(-2.0 * h)),
                 c_noise=sigma[t_prev] * torch.sqrt(1.0 - torch.exp(-2.0 * h)),
                 sigma_s=sigma[t], alpha_s=alpha[t])
        # second-order term: s1 = timesteps[step_index - 1].  At step_index 0 the Python index wraps to timesteps[-1], which is
        # what diffusers 0.18 evaluates when step() is called for a second trajectory without a new set_timesteps() (the
        # reference's main loop does exactly that for every scan after the first, pipeline:213-222): the previous scan's last
        # x0 prediction enters the first update with weight 0.5*c_x0/r0.
        h0 = lam[t] - lam[ts[step_index - 1]]
        c["inv_r0"] = 1.0 / (h0 / h)
        return {k: float(v) for k, v in c.items()}

    def step_index_of(self, timestep) -> int:
        t = int(timestep)
        hits = (self.timesteps.detach().cpu() == t).nonzero()
        return len(self.timesteps) - 1 if len(hits) == 0 else int(hits[0])

    def use_second_order(self, step_index: int) -> bool:
        n = len(self.timesteps)
        lower_order_final = (step_index == n - 1) and self.lower_order_final and n < 15
        return not (self.lower_order_nums < 1 or lower_order_final)

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, noise=None):
        i = self.step_index_of(timestep)
        c = self.coefficients(i)
        dev = sample.device
        f32 = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
        x0 = (sample - f32(c["sigma_s"]) * model_output) / f32(c["alpha_s"])
        second = self.use_second_order(i)
        self.model_outputs[0] = self.model_outputs[1]
        self.model_outputs[1] = x0
        if noise is None:       # diffusers draws the SDE noise in the model output's dtype (fp32 here), not in x0's (fp64)
            noise = torch.randn(model_output.shape, generator=generator, device=dev, dtype=model_output.dtype)
        prev = f32(c["c_sample"]) * sample + f32(c["c_x0"]) * x0
        if second:
            d1 = f32(c["inv_r0"]) * (x0 - self.model_outputs[0])
            prev = prev + 0.5 * f32(c["c_x0"]) * d1
        prev = prev + f32(c["c_noise"]) * noise
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        return {"prev_sample": prev} if return_dict else (prev,)
