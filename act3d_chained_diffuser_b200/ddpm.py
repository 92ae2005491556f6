"""DDPM posterior-step tables for the trajectory sampler (product code).

The reference constructs two ``diffusers`` DDPMSchedulers (diffusion_model.py:51-60):
positions use the "scaled_linear" betas, rotations "squaredcos_cap_v2", both with
prediction_type="sample" and diffusers' defaults (beta_start 1e-4, beta_end 0.02, clip_sample=True,
variance_type="fixed_small", "leading" timestep spacing).  ``diffusers`` is not a dependency of
this package: the closed-form coefficients of
    x_{t-1} = c_x0(t) * clip(x0_hat, -1, 1) + c_xt(t) * x_t + sigma(t) * eps      (t > 0)
are tabulated here in fp32 exactly as the scheduler evaluates them, and consumed by the
cd_post kernel.
"""
import math

import torch


def betas_for(schedule, n, beta_start=1e-4, beta_end=0.02):
    if schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    if schedule == "squaredcos_cap_v2":
        bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), 0.999) for i in range(n)], dtype=torch.float32)
    raise ValueError(schedule)


class PosteriorTable:
    """alphas_cumprod, inference timesteps and (c_x0, c_xt, sigma) per timestep."""

    def __init__(self, schedule, num_train_timesteps, num_inference_steps=None):
        self.num_train_timesteps = num_train_timesteps
        self.betas = betas_for(schedule, num_train_timesteps)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.set_timesteps(num_inference_steps or num_train_timesteps)

    def set_timesteps(self, n):
        if getattr(self, "num_inference_steps", None) == n and getattr(self, "coef", None) is not None:
            return          # the table is a pure function of (schedule, T, n): ~600 scalar tensor ops, 5 ms of host time per call
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = [int(round(i * ratio)) for i in range(n)][::-1]
        one = torch.tensor(1.0)
        coef = torch.zeros(self.num_train_timesteps, 3)
        for t in self.timesteps:
            prev = t - ratio
            a_t = self.alphas_cumprod[t]
            a_p = self.alphas_cumprod[prev] if prev >= 0 else one
            cur_alpha = a_t / a_p
            cur_beta = 1 - cur_alpha
            coef[t, 0] = (a_p ** 0.5 * cur_beta) / (1 - a_t)
            coef[t, 1] = cur_alpha ** 0.5 * (1 - a_p) / (1 - a_t)
            var = torch.clamp((1 - a_p) / (1 - a_t) * cur_beta, min=1e-20)
            coef[t, 2] = var ** 0.5 if t > 0 else 0.0
        self.coef = coef

    def add_noise(self, x0, noise, t):
        ac = self.alphas_cumprod.to(x0.device)
        a = (ac[t] ** 0.5).view(-1, *([1] * (x0.dim() - 1)))
        b = ((1 - ac[t]) ** 0.5).view(-1, *([1] * (x0.dim() - 1)))
        return a * x0 + b * noise
