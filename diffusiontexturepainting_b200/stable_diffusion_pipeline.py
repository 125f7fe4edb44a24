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


def synthetic_allowed() -> bool:
    return os.environ.get("DTP_SYNTHETIC_WEIGHTS", "0") == "1"


def load_state_dicts(cfg, lora_path=None, seed=20240726, synthetic=None):
    """The four checkpoint components the reference loads — SD-1.5-inpaint UNet and VAE (models.py:796-813), the LoRA file
    and image_encoder.pth (trt_model.py:48,58) — from the reference's paths. The reference crashes when one is absent; so
    does this: a component that is missing raises CheckpointError unless seeded synthetic weights were asked for
    explicitly (`synthetic=True` or DTP_SYNTHETIC_WEIGHTS=1: bench / tests / this image, which has neither network nor
    weight files). In synthetic mode every substituted component is printed, and a real UNet never receives random LoRA
    factors (the `up` factors of a missing LoRA file are zeroed, so the merge is a no-op)."""
    from .checkpoints import CheckpointError, load_real_checkpoints
    if synthetic is None:
        synthetic = synthetic_allowed()
    unet_sd, vae_sd, enc_sd = W.synth_model(cfg, seed)  # the inventory (keys and shapes), filled with seeded values
    if cfg.name != "sd15-inpaint":
        return unet_sd, vae_sd, enc_sd  # the narrow test configuration exists only with synthetic weights
    hf = os.environ.get("DTP_HF_DIR", "./HF_cache/stable-diffusion-inpainting")
    enc_path = os.environ.get("DTP_IMAGE_ENCODER", "/workspace/checkpoints/image_encoder.pth")
    # a checkpoint file that exists must load completely (DTP_STRICT_WEIGHTS=0 downgrades that to a printed report)
    strict = os.environ.get("DTP_STRICT_WEIGHTS", "1") != "0"
    reports = load_real_checkpoints(unet_sd, vae_sd, enc_sd, hf, lora_path, enc_path, strict=strict)
    for r in reports:
        print("[dtp] " + r.summary())
    have = {r.what.split(" <- ")[0] for r in reports}
    missing = [c for c in ("unet", "vae", "lora", "image encoder") if c not in have]
    if missing and not synthetic:
        where = {"unet": os.path.join(hf, "unet"), "vae": os.path.join(hf, "vae"), "lora": str(lora_path),
                 "image encoder": enc_path}
        raise CheckpointError(
            "checkpoint component(s) not found: " + ", ".join(f"{c} ({where[c]})" for c in missing) +
            ". Set DTP_SYNTHETIC_WEIGHTS=1 to run on seeded random weights instead (benchmarks / tests only).")
    if missing:
        print("[dtp] WARNING: SYNTHETIC (seeded random) weights in use for: " + ", ".join(missing))
        if "lora" in missing and "unet" in have:
            for k in unet_sd:
                if k.endswith("_lora.up.weight"):
                    unet_sd[k] = torch.zeros_like(unet_sd[k])
            print("[dtp] LoRA file absent next to a real UNet: LoRA factors zeroed (merge is a no-op)")
    return unet_sd, vae_sd, enc_sd
