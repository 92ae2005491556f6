"""Oracle: DDPM scheduler restatement (test infrastructure only).

PARITY UNPINNED: the reference takes ``DDPMScheduler`` from the third-party
package ``diffusers`` (unpinned, README.md:29; call sites
model/trajectory_optimization/diffusion_model.py:1,51-60,87-88,99,111-116,291,296-303)
which is not installed here and is not part of /root/reference.  This file restates
the published algorithm of ``diffusers.schedulers.scheduling_ddpm.DDPMScheduler``
(Ho et al. 2020, eq. 7 posterior; diffusers defaults beta_start=1e-4, beta_end=0.02,
variance_type="fixed_small", clip_sample=True with range 1.0, timestep_spacing
"leading", steps_offset 0) for the two configurations the reference constructs.
It is anchored by closed-form known answers in tests/test_ddpm.py, which also checks it against
``diffusers`` itself whenever that package is importable (skipped in this image).

It doubles as the import stub that lets the unmodified reference be imported in the
build container (oracle/ref_import.py).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


def make_betas(schedule: str, n: int, beta_start=1e-4, beta_end=0.02) -> torch.Tensor:
    if schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    if schedule == "squaredcos_cap_v2":
        def abar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        vals = [min(1 - abar((i + 1) / n) / abar(i / n), 0.999) for i in range(n)]
        return torch.tensor(vals, dtype=torch.float32)
    if schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    raise NotImplementedError(schedule)


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_schedule="linear", prediction_type="epsilon",
                 beta_start=1e-4, beta_end=0.02, clip_sample=True, clip_sample_range=1.0):
        assert prediction_type == "sample", "the reference only uses prediction_type='sample'"
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_schedule=beta_schedule,
                                      prediction_type=prediction_type, clip_sample=clip_sample,
                                      clip_sample_range=clip_sample_range)
        self.betas = make_betas(beta_schedule, num_train_timesteps, beta_start, beta_end)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, num_inference_steps):
        n = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        ratio = n // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def step_coefficients(self, t: int):
        """(c_x0, c_xt, sigma) of  x_{t-1} = c_x0 * clip(x0) + c_xt * x_t + sigma * eps."""
        steps = self.num_inference_steps or self.config.num_train_timesteps
        prev_t = t - self.config.num_train_timesteps // steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - a_t
        beta_prod_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        c_x0 = (a_prev ** 0.5 * cur_beta) / beta_prod_t
        c_xt = cur_alpha ** 0.5 * beta_prod_prev / beta_prod_t
        var = torch.clamp(beta_prod_prev / beta_prod_t * cur_beta, min=1e-20)
        sigma = var ** 0.5 if t > 0 else torch.tensor(0.0)
        return c_x0, c_xt, sigma

    def step(self, model_output, timestep, sample, generator=None):
        t = int(timestep)
        c_x0, c_xt, sigma = self.step_coefficients(t)
        x0 = model_output
        if self.config.clip_sample:
            r = self.config.clip_sample_range
            x0 = x0.clamp(-r, r)
        prev = c_x0.to(sample.dtype) * x0 + c_xt.to(sample.dtype) * sample
        if t > 0:
            noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                dtype=model_output.dtype)
            prev = prev + sigma.to(sample.dtype) * noise
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)

    def add_noise(self, original, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original.device, dtype=original.dtype)
        timesteps = timesteps.to(original.device)
        a = ac[timesteps] ** 0.5
        b = (1 - ac[timesteps]) ** 0.5
        while a.dim() < original.dim():
            a = a.unsqueeze(-1)
            b = b.unsqueeze(-1)
        return a * original + b * noise
