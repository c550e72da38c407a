"""ORACLE (test infrastructure) -- DDIM scheduler, CFG combine, add_noise and FreeInit low-pass mixing as the reference
uses them (configs/inference/inference.yaml:36-42; animatediff/pipelines/pipeline.py:987-1031;
custom/threestudio-animate3d/guidance/animatemv_guidance.py:428-487).  diffusers 0.28.0 `DDIMScheduler` and
`FreeInitMixin` are not installed here: restated from SURVEY.md Appendix B.10 / B.11 ("parity unpinned")."""
from __future__ import annotations

import math

import torch


class DDIMOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)   # "linear"
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)       # set_alpha_to_one=True
        self.T = num_train_timesteps
        self.steps_offset = steps_offset

    def set_timesteps(self, n):
        self.n = n
        ratio = self.T // n
        self.timesteps = (torch.arange(0, n) * ratio).round().flip(0).long() + self.steps_offset   # "leading"
        return self.timesteps

    def step(self, eps, t, x):
        prev_t = t - self.T // self.n
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        x0 = (x - (1 - a_t).sqrt() * eps) / a_t.sqrt()
        return a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps, x0

    def add_noise(self, x, noise, t):
        a = self.alphas_cumprod[t]
        while a.ndim < x.ndim:
            a = a[..., None]
        return a.sqrt() * x + (1 - a).sqrt() * noise


def cfg_pipeline(noise_pred, g):
    """pipeline.py:1023-1025 -- (uncond, cond) order."""
    u, c = noise_pred.chunk(2)
    return u + g * (c - u)


def cfg_guidance(noise_pred, g):
    """animatemv_guidance.py:452-459 -- (cond, uncond) order, text + g*(text - uncond)."""
    c, u = noise_pred.chunk(2)
    return c + g * (c - u)


def denoise_step(latents, noise_pred2, first_frame, g, sched: DDIMOracle, t):
    """One iteration of pipeline.py:1006-1031 after the UNet call."""
    eps = cfg_pipeline(noise_pred2, g)
    prev, _ = sched.step(eps, int(t), latents)
    return torch.cat([first_frame, prev[:, :, 1:]], dim=2)


def butterworth_lpf(shape, order=4, d_s=0.25, d_t=0.25):
    """FreeInit butterworth low-pass filter over (F,H,W) (Appendix B.11)."""
    T, H, W = shape[-3], shape[-2], shape[-1]
    t = torch.arange(T)[:, None, None].float()
    h = torch.arange(H)[None, :, None].float()
    w = torch.arange(W)[None, None, :].float()
    d2 = ((d_s / d_t) * (2 * t / T - 1)) ** 2 + (2 * h / H - 1) ** 2 + (2 * w / W - 1) ** 2
    return (1.0 / (1.0 + (d2 / d_s ** 2) ** order)).expand(shape)


def freeinit_mix(z_T, z_rand, lpf):
    dims = (-3, -2, -1)
    zf = torch.fft.fftshift(torch.fft.fftn(z_T, dim=dims), dim=dims)
    rf = torch.fft.fftshift(torch.fft.fftn(z_rand, dim=dims), dim=dims)
    mixed = zf * lpf + rf * (1 - lpf)
    return torch.fft.ifftn(torch.fft.ifftshift(mixed, dim=dims), dim=dims).real
