"""ORACLE (test infrastructure): the reference's DDIMScheduler (trt_inference/utilities.py:370-529) and
initialize_timesteps (stable_diffusion_pipeline.py:348-355) restated; validated against the reference class executed
in place (tests/test_oracle_reference.py) and against tests/golden/ddim_*.json."""
from __future__ import annotations

import numpy as np
import torch


class DDIM:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        # utilities.py:382-394 with the constructor arguments of stable_diffusion_pipeline.py:109-116
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.train_alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.train_alphas_cumprod[0]  # set_alpha_to_one=False
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None

    def set_timesteps(self, steps: int):
        """utilities.py:432-439 followed by configure() (utilities.py:408-417): per-step alpha table."""
        steps = int(steps)
        self.num_inference_steps = steps
        ratio = self.num_train_timesteps // steps
        ts = (np.arange(0, steps) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts)
        self.alphas_cumprod = self.train_alphas_cumprod[self.timesteps]

    def initialize_timesteps(self, steps: int, strength: float = 1.0):
        """stable_diffusion_pipeline.py:348-355: with steps_offset = 1 and strength = 1 this yields t_start = 1, i.e. the
        loop consumes timesteps[1:] (steps - 1 UNet evaluations)."""
        self.set_timesteps(steps)
        offset = self.steps_offset
        init_timestep = min(int(steps * strength) + offset, steps)
        t_start = max(steps - init_timestep + offset, 0)
        return self.timesteps[t_start:], t_start

    def alpha_pair(self, idx: int):
        a_t = self.alphas_cumprod[idx]
        a_prev = self.alphas_cumprod[idx + 1] if idx + 1 < self.num_inference_steps else self.final_alpha_cumprod
        return a_t, a_prev

    def step(self, model_output, sample, idx):
        """utilities.py:441-522 with eta = 0, epsilon prediction, clip_sample False."""
        a_t, a_prev = self.alpha_pair(idx)
        beta_t = 1 - a_t
        x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5
        direction = (1 - a_prev) ** 0.5 * model_output
        return a_prev ** 0.5 * x0 + direction
