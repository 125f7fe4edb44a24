"""Drop-in for trt_inference/stable_diffusion_pipeline.py: same constructor / loadEngines / loadResources surface, but the
"engines" are the sm_100a kernels of libdtp_sm100.so (no ONNX export, no TensorRT build, no plan cache)."""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import weights as W
from .engine import Engine, arena_estimate
from .scheduler import DDIMScheduler


class StableDiffusionPipeline:
    def __init__(self, version="1.5", inpaint=False, stages=("vae_encoder", "unet", "vae"), max_batch_size=16,
                 denoising_steps=50, texture_guidance_steps=50, scheduler="DDIM", guidance_scale=7.5, device="cuda",
                 output_dir=".", hf_token=None, verbose=False, nvtx_profile=False, model_config=None,
                 state_dicts=None, weight_seed=20240726):
        # stable_diffusion_pipeline.py:83: classifier-free guidance needs a scale above 1
        assert guidance_scale > 1.0, "Guidance scale must be greater than 1.0 for classifier-free guidance"
        if scheduler != "DDIM":
            raise ValueError("only the DDIM scheduler is on the stamp path (trt_model.py:36)")
        if version != "1.5" or not inpaint:
            raise ValueError("only the SD-1.5 inpainting pipeline is implemented")
        self.denoising_steps = denoising_steps
        self.texture_guidance_steps = texture_guidance_steps
        self.guidance_scale = guidance_scale
        self.texture_guidance_scale = 0.0
        self.max_batch_size = max_batch_size
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.verbose = verbose
        self.nvtx_profile = nvtx_profile
        self.stages = list(stages)
        self.inpaint = inpaint
        self.model_config = model_config or W.sd15_config()
        self._state_dicts = state_dicts
        self._weight_seed = weight_seed
        self.scheduler = DDIMScheduler(device=self.device, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012)
        self.generator = None
        self.engine: Optional[Engine] = None
        self._schedule_key = None

    # ------------------------------------------------------------------ stable_diffusion_pipeline.py:189-334
    def loadEngines(self, engine_dir=None, onnx_dir=None, onnx_opset=16, opt_batch_size=1, opt_image_height=256,
                    opt_image_width=256, text_maxlen=14, lora_path=None, timing_cache=None, **_ignored):
        """Creates the native engine and uploads the (LoRA-merged) weights. engine_dir / onnx_dir / onnx_opset /
        timing_cache are accepted for signature compatibility and ignored: nothing is exported, built or cached."""
        dev_index = self.device.index or 0
        arena = arena_estimate(self.model_config, max(1, opt_batch_size), max(opt_image_height, opt_image_width))
        arena = int(os.environ.get("DTP_ARENA_BYTES", arena))
        self.engine = Engine(self.model_config, dev_index, arena_bytes=arena)
        if getattr(self, "_prepacked", None) is not None:
            self.engine.load_prepacked(self._prepacked)
            self._prepacked = None
            return
        sds = self._state_dicts or load_state_dicts(self.model_config, lora_path, self._weight_seed)
        self.engine.load_state_dicts(*sds)
        self._state_dicts = None

    # ------------------------------------------------------------------ stable_diffusion_pipeline.py:138-156
    def loadResources(self, image_height, image_width, batch_size, seed):
        self.set_seed(seed)
        self.scheduler.set_timesteps(self.denoising_steps)
        self.scheduler.configure()

    def set_seed(self, seed):
        self.generator = torch.Generator(device=self.device).manual_seed(seed) if seed else None
        self.noise_generator = torch.Generator(device=self.device).manual_seed((seed or 0) + 1)

    # ------------------------------------------------------------------ stable_diffusion_pipeline.py:340-355
    def initialize_latents(self, batch_size, unet_channels, latent_height, latent_width):
        latents = torch.randn((batch_size, unet_channels, latent_height, latent_width), device=self.device,
                              dtype=torch.float32, generator=self.generator)
        return latents * self.scheduler.init_noise_sigma

    def initialize_timesteps(self, timesteps, strength):
        timesteps = int(timesteps)
        self.scheduler.set_timesteps(timesteps)
        self.scheduler.configure()
        offset = self.scheduler.steps_offset
        init_timestep = min(int(timesteps * strength) + offset, timesteps)
        t_start = max(timesteps - init_timestep + offset, 0)
        return self.scheduler.timesteps[t_start:], t_start

    def teardown(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None


def load_state_dicts(cfg, lora_path=None, seed=20240726):
    """Real checkpoints when present at the reference's paths (trt_model.py:48,58; models.py:796-813), else the seeded
    synthetic inventory (this image has neither network nor weight files)."""
    unet_sd, vae_sd, enc_sd = W.synth_model(cfg, seed)
    if cfg.name != "sd15-inpaint":
        return unet_sd, vae_sd, enc_sd
    hf = os.environ.get("DTP_HF_DIR", "./HF_cache/stable-diffusion-inpainting")
    from .checkpoints import load_real_checkpoints
    # a checkpoint file that exists must load completely (DTP_STRICT_WEIGHTS=0 downgrades that to a printed report)
    strict = os.environ.get("DTP_STRICT_WEIGHTS", "1") != "0"
    reports = load_real_checkpoints(unet_sd, vae_sd, enc_sd, hf, lora_path,
                                    os.environ.get("DTP_IMAGE_ENCODER", "/workspace/checkpoints/image_encoder.pth"),
                                    strict=strict)
    for r in reports:
        print("[dtp] " + r.summary())
    return unet_sd, vae_sd, enc_sd
