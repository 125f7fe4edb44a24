"""ORACLE (test infrastructure): the reference's stamp path end to end in plain PyTorch.

Follows TRTConditionalInpainter.set_brush / generate_raw (trt_inference/trt_model.py:79-121), add_extra_context
(handler.py:25-33; kornia dilation restated as a padded max-pool), InpaintPipeline.infer
(inpaint_pipeline.py:52-153), denoise_latent / encode_image / decode_latent (stable_diffusion_pipeline.py:407-484) and
ConditionalInpainterBase.generate (model_base.py:51-58)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import image_encoder as ie
from . import unet as un
from . import vae as va
from .ddim import DDIM


def add_extra_context(source_image, masked_image, mask, pad=150):
    """handler.py:25-33. kornia.morphology.dilation(mask, ones(pad,pad)) with the default geodesic border (-1e4) and
    origin pad//2 == max over the window [i - pad//2, i + pad - pad//2 - 1]."""
    pad = max(int(pad), 1)
    lo, hi = pad // 2, pad - pad // 2 - 1
    dil = F.max_pool2d(F.pad(mask, (lo, hi, lo, hi), value=-1e4), kernel_size=pad, stride=1)
    hint_mask = 1 - dil
    hint_image = source_image * hint_mask
    return masked_image + hint_image, torch.clamp(mask + hint_mask, min=0, max=1)


def canvas_preprocess(canvas, brush_image, pad):
    """trt_model.py:103-109 -> (masked_images, masks, context_masked_image, context_mask) with 1 = generate."""
    images = canvas[:, :3, ...] * 2 - 1.0
    masks = canvas[:, 3:, ...]
    masked_images = images * masks
    ctx_img, ctx_mask = add_extra_context(brush_image * 2 - 1, masked_images, masks, pad=pad)
    return masked_images, 1 - masks, ctx_img, 1 - ctx_mask


class OraclePipeline:
    """State dicts must be LoRA-merged (weights.merge_lora) or carry the un-merged processor keys; dtype/device follow
    the tensors of the state dicts."""

    def __init__(self, cfg, unet_sd, vae_sd, enc_sd, resolution):
        self.cfg = cfg
        self.unet_sd, self.vae_sd, self.enc_sd = unet_sd, vae_sd, enc_sd
        self.res = resolution
        self.sched = DDIM()
        self.image = None
        self.conditioning = None
        p = next(iter(unet_sd.values()))
        self.device, self.dtype = p.device, p.dtype

    # trt_model.py:79-88
    def set_brush(self, image):
        self.image = ie.crop_resize_square(image, self.res).unsqueeze(0).to(self.device)
        emb, uncond = ie.encode_image(self.enc_sd, self.cfg.enc, self.image.to(self.dtype))
        self.conditioning = (emb, uncond)
        return self.conditioning

    def vae_encode(self, x, noise=None):
        # stable_diffusion_pipeline.py:464-474
        return 0.18215 * va.encode_sample(self.vae_sd, self.cfg.vae, x.to(self.dtype), noise).float()

    def unet(self, sample, t, ctx):
        return un.unet_forward(self.unet_sd, self.cfg.unet, sample.to(self.dtype), t, ctx.to(self.dtype)).float()

    # inpaint_pipeline.py:52-153 (+ stable_diffusion_pipeline.py:407-462)
    def infer(self, prompt, negative_prompt, input_image, mask_image, context_masked_image, context_mask, steps, cfg_w,
              tg_w, tg_steps, init_latents, vae_noise=(None, None), strict=False, trace=None):
        B = input_image.shape[0]
        h = input_image.shape[-1] // 8
        latents = init_latents.float() * self.sched.init_noise_sigma
        mask = F.interpolate(mask_image, size=(h, h))
        cmask = F.interpolate(context_mask, size=(h, h))
        mask = torch.cat([mask, mask, cmask])
        if strict:
            self.sched.set_timesteps(steps)
            timesteps, t_start = self.sched.timesteps, 0
        else:
            timesteps, t_start = self.sched.initialize_timesteps(steps, 1.0)
        ml = self.vae_encode(input_image, vae_noise[0])
        cml = self.vae_encode(context_masked_image, vae_noise[1])
        masked_latents = torch.cat([ml, ml, cml])
        emb = torch.cat([negative_prompt.expand(B, -1, -1), prompt.expand(B, -1, -1), prompt.expand(B, -1, -1)])
        tg = float(tg_w)
        for step_index, t in enumerate(timesteps):
            if step_index > int(tg_steps) - 1:
                tg = 0.0
            x = torch.cat([latents] * 3)
            x = torch.cat([x, mask, masked_latents], dim=1)
            eps = self.unet(x, t, emb)
            eu, ec, et = eps.chunk(3)
            eps = eu + float(cfg_w) * (ec - eu) + tg * (et - ec)
            latents = self.sched.step(eps, latents, t_start + step_index)
            if trace is not None:
                trace.append(latents.clone())
        latents = latents / 0.18215
        images = va.decode(self.vae_sd, self.cfg.vae, latents.to(self.dtype)).float()
        return (images / 2 + 0.5).clamp(0, 1)

    # trt_model.py:90-121
    def generate_raw(self, canvas, init_latents, vae_noise=(None, None), strict=False, trace=None, **settings):
        canvas = canvas.to(self.device).float()
        mi, m, ci, cm = canvas_preprocess(canvas, self.image.float(), int(settings["context_pad"]))
        emb, uncond = self.conditioning
        return self.infer(emb, uncond, mi, m, ci, cm, int(settings["steps"]), float(settings["cfg_weight"]),
                          float(settings["tg_weight"]), int(settings["tg_steps"]), init_latents, vae_noise, strict,
                          trace)

    # model_base.py:51-58
    def generate(self, canvas, init_latents, **kw):
        result = self.generate_raw(canvas, init_latents, **kw)
        canvas = canvas.to(self.device).float()
        alpha = canvas[:, 3:, ...]
        return canvas[:, :3, ...] * alpha + result[:, :3, ...] * (1 - alpha)
