"""Host-side DDIM tables for the stamp path (product code; numpy/torch CPU only — the per-step arithmetic runs in the
fused guidance+DDIM kernel). Mirrors the reference's DDIMScheduler as configured by the server
(trt_inference/utilities.py:370-439, stable_diffusion_pipeline.py:109-116,348-355):
num_train_timesteps 1000, scaled-linear betas 0.00085..0.012, steps_offset 1, set_alpha_to_one False, eta 0."""
from __future__ import annotations

import numpy as np
import torch


class DDIMScheduler:
    def __init__(self, device="cuda", num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, clip_sample=False,
                 set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon"):
        assert prediction_type == "epsilon" and not clip_sample and not set_alpha_to_one
        self.device = device
        self.num_train_timesteps = num_train_timesteps
        self.beta_start, self.beta_end = beta_start, beta_end  # kept: update_infer_settings rebuilds the table from them
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self._rebuild()

    def _rebuild(self):
        betas = torch.linspace(self.beta_start ** 0.5, self.beta_end ** 0.5, self.num_train_timesteps,
                               dtype=torch.float32) ** 2
        self.train_alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.train_alphas_cumprod[0]

    def set_timesteps(self, num_inference_steps: int):
        n = int(num_inference_steps)  # wire settings arrive as numpy uint8 (server_io.py:105-108)
        if n < 1:
            raise ValueError("denoising steps must be >= 1")
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = torch.from_numpy((np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
                                          + self.steps_offset)

    def configure(self):
        self.alphas_cumprod = self.train_alphas_cumprod[self.timesteps]

    def scale_model_input(self, sample, *a, **k):
        return sample

    def evaluation_schedule(self, t_start: int):
        """(timesteps, alpha_t, alpha_prev) of the evaluations the loop executes from index t_start on: alpha_prev is the
        next table entry, or final_alpha_cumprod after the last one (utilities.py:467-470)."""
        n = self.num_inference_steps
        ts, at, ap = [], [], []
        for idx in range(t_start, n):
            ts.append(float(self.timesteps[idx]))
            at.append(float(self.alphas_cumprod[idx]))
            ap.append(float(self.alphas_cumprod[idx + 1]) if idx + 1 < n else float(self.final_alpha_cumprod))
        return ts, at, ap
