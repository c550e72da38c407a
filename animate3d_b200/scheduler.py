"""DDIM scheduler constants of the released MV-VDM (configs/inference/inference.yaml:36-42 of the reference: linear betas
0.00085..0.012, 1000 train steps, leading spacing, steps_offset 1, eta 0).  The update itself runs in the
a3d_ddim_cfg_step kernel fused with classifier-free guidance and the frame-0 re-injection of pipeline.py:1023-1031."""
from __future__ import annotations

import numpy as np


class DDIMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear",
                 steps_offset=1, clip_sample=False, set_alpha_to_one=True, **unused):
        if beta_schedule != "linear" or clip_sample:
            raise NotImplementedError("only the released configuration (linear betas, no clipping) is supported")
        betas = np.linspace(beta_start, beta_end, num_train_timesteps, dtype=np.float32)
        self.alphas_cumprod = np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32)
        self.final_alpha_cumprod = np.float32(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.order = 1
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        self.timesteps = (np.arange(0, num_inference_steps) * ratio).round()[::-1].astype(np.int64) + self.steps_offset
        return self.timesteps

    def alphas_for(self, t: int):
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_p

    def scale_model_input(self, sample, t=None):
        return sample
