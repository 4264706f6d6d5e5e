"""Generates tests/golden/known_answers.json: closed-form known-answer values for the hot path
 computed with plain torch/numpy float32 arithmetic, independent of oracle/ and of
lidiff_b200/.  The reference ships no golden vectors (it has no tests) and its dependencies cannot be
installed here, so these are the only values that pin the oracle.

    python tests/golden/make_known_answers.py
"""
import json
import os

import numpy as np
import torch


def main():
    T = 1000
    betas = torch.linspace(3.5e-5, 0.007, T, dtype=torch.float32)          # config.yaml:30-33
    ac = torch.cumprod(1.0 - betas, 0)
    alpha, sigma = torch.sqrt(ac), torch.sqrt(1 - ac)
    lam = torch.log(alpha) - torch.log(sigma)

    def timesteps(n):
        ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        _, idx = np.unique(ts, return_index=True)
        return ts[np.sort(idx)].tolist()

    def first_order(t, tp):
        h = lam[tp] - lam[t]
        return dict(h=float(h), c_sample=float(sigma[tp] / sigma[t] * torch.exp(-h)),
                    c_x0=float(alpha[tp] * (1 - torch.exp(-2.0 * h))),
                    c_noise=float(sigma[tp] * torch.sqrt(1.0 - torch.exp(-2.0 * h))))

    half = 48
    freq = torch.from_numpy(np.exp(np.arange(0, half) * -(np.log(10000) / (half - 1)))).float()
    arg = torch.tensor([999])[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], 1)[0]

    out = {
        "timesteps_50": timesteps(50), "timesteps_1000_len": len(timesteps(1000)),
        "timesteps_1000_head": timesteps(1000)[:3], "timesteps_1000_tail": timesteps(1000)[-3:],
        "timesteps_1": timesteps(1),
        "tables": {str(t): dict(alphas_cumprod=float(ac[t]), alpha_t=float(alpha[t]), sigma_t=float(sigma[t]),
                                lambda_t=float(lam[t])) for t in (999, 979, 20, 0)},
        "step_999_979": first_order(999, 979), "step_20_0": first_order(20, 0),
        "temb_999_0_4": emb[:4].tolist(), "temb_999_48_52": emb[48:52].tolist(),
        "round_half_even": torch.round(torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5])).tolist(),
        "param_counts": {"diff": 24174515, "enc": 8497952, "refine18": 21722926},
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
